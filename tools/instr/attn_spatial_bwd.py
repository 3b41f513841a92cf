p='hma_b200/csrc/attn_spatial_bwd.cu'
s=open(p).read()
def rep(a, b, count=1):
    global s
    assert s.count(a) >= 1, a[:70]
    s = s.replace(a, b, count)
rep('''  pdl_wait();
  if (warp == kComputeWarps) {
    // start the operand loads first''','''  pdl_wait();
  if (threadIdx.x == 0) HMA_TL(0, 0);
  if (warp == kComputeWarps) {
    // start the operand loads first''')
rep('''      mbar_wait(smem_u32(&bar_load), 0);
      tc_fence_after();
      constexpr uint32_t kHi64''','''      mbar_wait(smem_u32(&bar_load), 0);
      tc_fence_after();
      if (role == 0) HMA_TL(11, 0);
      constexpr uint32_t kHi64''')
rep('''          umma_commit(smem_u32(&bar_sdp));
          if (++qt == ntile) { qt = 0; ++kt; }''','''          umma_commit(smem_u32(&bar_sdp));
          HMA_TL(1, it);
          if (++qt == ntile) { qt = 0; ++kt; }''')
rep('''          umma_commit(smem_u32(&bar_mma[bsel]));
          if (qt == ntile - 1) umma_commit(smem_u32(&bar_kv));
          if (++qt == ntile) { qt = 0; ++kt; }''','''          umma_commit(smem_u32(&bar_mma[bsel]));
          if (qt == ntile - 1) umma_commit(smem_u32(&bar_kv));
          HMA_TL(3, it);
          if (++qt == ntile) { qt = 0; ++kt; }''')
rep('''          umma_commit(smem_u32(&bar_mma[bsel]));
          if (it == NI - 1) umma_commit(smem_u32(&bar_final));
          if (++qt == ntile) { qt = 0; ++kt; }''','''          umma_commit(smem_u32(&bar_mma[bsel]));
          if (it == NI - 1) umma_commit(smem_u32(&bar_final));
          HMA_TL(4, it);
          if (++qt == ntile) { qt = 0; ++kt; }''')
rep('''        if (it >= 2) mbar_wait(smem_u32(&bar_mma[bsel]), (uint32_t)(((it - 2) >> 1) & 1));
        mbar_wait(smem_u32(&bar_sdp), (uint32_t)(it & 1));
        tc_fence_after();''','''        if (threadIdx.x == 0) HMA_TL(5, it);
        if (it >= 2) mbar_wait(smem_u32(&bar_mma[bsel]), (uint32_t)(((it - 2) >> 1) & 1));
        if (threadIdx.x == 0) HMA_TL(2, it);
        mbar_wait(smem_u32(&bar_sdp), (uint32_t)(it & 1));
        tc_fence_after();
        if (threadIdx.x == 0) HMA_TL(6, it);''')
rep('''          tc_fence_before();
          mbar_arrive(smem_u32(&bar_tfree));''','''          tc_fence_before();
          mbar_arrive(smem_u32(&bar_tfree));
          if (threadIdx.x == 0) HMA_TL(7, it);''')
rep('''        fence_proxy_async();
        mbar_arrive(smem_u32(&bar_pds[bsel]));''','''        fence_proxy_async();
        mbar_arrive(smem_u32(&bar_pds[bsel]));
        if (threadIdx.x == 0) HMA_TL(8, it);''')
rep('''    mbar_wait(smem_u32(&bar_final), 0);
    tc_fence_after();''','''    mbar_wait(smem_u32(&bar_final), 0);
    tc_fence_after();
    if (threadIdx.x == 0) HMA_TL(9, 0);''')
rep('''  tc_fence_before();
  __syncthreads();
  if (warp == kComputeWarps) {
    tc_fence_after();
    tmem_dealloc''','''  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) HMA_TL(10, 0);
  if (warp == kComputeWarps) {
    tc_fence_after();
    tmem_dealloc''')
s=s.rstrip('\n')+'\n\nHMA_DEFINE_TIMELINE_READER(hma_timeline_attn_spatial_bwd)\n'
open(p,'w').write(s)
