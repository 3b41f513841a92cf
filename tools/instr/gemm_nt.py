"""Source patcher: timeline stamps for gemm_nt.cu (git checkout the file afterwards)."""
p='hma_b200/csrc/gemm_nt.cu'
s=open(p).read()
def rep(a, b):
    global s
    assert s.count(a) == 1, a[:60]
    s = s.replace(a, b)
rep('''  pdl_wait();               // everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
''','''  pdl_wait();               // everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  if (threadIdx.x == 0) HMA_TL(0, 0);
''')
rep('''        mbar_wait(smem_u32(&bar_tempty[as]), aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem''','''        HMA_TL(1, it);
        mbar_wait(smem_u32(&bar_tempty[as]), aph ^ 1u);
        tc_fence_after();
        HMA_TL(2, it);
        const uint32_t d_tmem''')
rep('''        umma_commit(smem_u32(&bar_tfull[as]));
      }''','''        umma_commit(smem_u32(&bar_tfull[as]));
        HMA_TL(3, it);
      }''')
rep('''      float4 pf[2][8];
      uint4 po[2][4];
''', '''      float4 pf[2][8];
      uint4 po[2][4];
      if (threadIdx.x == 128) HMA_TL(4, it);
''')
rep('''      mbar_wait(smem_u32(&bar_tfull[as]), aph);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < kChunks; ++j) {''','''      mbar_wait(smem_u32(&bar_tfull[as]), aph);
      tc_fence_after();
      if (threadIdx.x == 128) HMA_TL(5, it);
#pragma unroll
      for (int j = 0; j < kChunks; ++j) {''')
rep('''      tc_fence_before();
      mbar_arrive(smem_u32(&bar_tempty[as]));
      if constexpr (LN) {''','''      tc_fence_before();
      mbar_arrive(smem_u32(&bar_tempty[as]));
      if (threadIdx.x == 128) HMA_TL(6, it);
      if constexpr (LN) {''')
rep('''        pend_row0 = row0;
        pend_it = it;''', '''        pend_row0 = row0;
        pend_it = it;
        if (threadIdx.x == 128) HMA_TL(7, it);''')
rep('''      const int b = pend_it & 1;
      mbar_wait(smem_u32(&ln_xbar[b][ew]), (uint32_t)(pend_it >> 1) & 1u);''', '''      const int b = pend_it & 1;
      if (threadIdx.x == 128) HMA_TL(8, pend_it);
      mbar_wait(smem_u32(&ln_xbar[b][ew]), (uint32_t)(pend_it >> 1) & 1u);
      if (threadIdx.x == 128) HMA_TL(9, pend_it);''')
rep('''    };
    int it = 0;''', '''      if (threadIdx.x == 128) HMA_TL(11, pend_it);
    };
    int it = 0;''')
rep('''  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);''','''  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) HMA_TL(10, 0);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);''')
s=s.rstrip('\n')+'\n\nHMA_DEFINE_TIMELINE_READER(hma_timeline_gemm_nt)\n'
open(p,'w').write(s)
