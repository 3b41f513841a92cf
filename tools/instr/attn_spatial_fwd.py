p='hma_b200/csrc/attn_spatial_fwd.cu'
s=open(p).read()
def rep(a, b):
    global s
    assert a in s, a[:60]
    s = s.replace(a, b)
rep('''  pdl_wait();  // several waves of CTAs: dependents are released by CTA exit, not early
''','''  pdl_wait();  // several waves of CTAs: dependents are released by CTA exit, not early
  if (threadIdx.x == 0) HMA_TL(0, 0);
''')
rep('''      mbar_wait(bl, 0);
      tc_fence_after();
      // ------------------------------------------------ MMA issue''','''      mbar_wait(bl, 0);
      tc_fence_after();
      HMA_TL(11, 0);
      // ------------------------------------------------ MMA issue''')
rep('''        umma_commit(smem_u32(&bar_s));
        mbar_wait(smem_u32(&bar_p), pp); pp ^= 1u;
        tc_fence_after();
        if (t > 0) {''','''        umma_commit(smem_u32(&bar_s));
        HMA_TL(12, t);
        mbar_wait(smem_u32(&bar_p), pp); pp ^= 1u;
        tc_fence_after();
        HMA_TL(13, t);
        if (t > 0) {''')
rep('''          umma_commit(smem_u32(&bar_s));
          mbar_wait(smem_u32(&bar_p), pp); pp ^= 1u;
          tc_fence_after();
          for (int kk = 0; kk < nb / 16; ++kk)''','''          umma_commit(smem_u32(&bar_s));
          HMA_TL(14, t);
          mbar_wait(smem_u32(&bar_p), pp); pp ^= 1u;
          tc_fence_after();
          HMA_TL(15, t);
          for (int kk = 0; kk < nb / 16; ++kk)''')
rep('''      mbar_wait(smem_u32(&bar_s), ps); ps ^= 1u;
      tc_fence_after();
      softmax_half(na, ma, la, live);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_p));''','''      if (threadIdx.x == 0) HMA_TL(1, t);
      mbar_wait(smem_u32(&bar_s), ps); ps ^= 1u;
      tc_fence_after();
      if (threadIdx.x == 0) HMA_TL(2, t);
      softmax_half(na, ma, la, live);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_p));
      if (threadIdx.x == 0) HMA_TL(3, t);''')
rep('''        mbar_wait(smem_u32(&bar_s), ps); ps ^= 1u;
        tc_fence_after();
        softmax_half(nb, mb, lb, live);
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(smem_u32(&bar_p));
      }''','''        mbar_wait(smem_u32(&bar_s), ps); ps ^= 1u;
        tc_fence_after();
        if (threadIdx.x == 0) HMA_TL(4, t);
        softmax_half(nb, mb, lb, live);
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(smem_u32(&bar_p));
        if (threadIdx.x == 0) HMA_TL(5, t);
      }''')
rep('''      mbar_wait(smem_u32(&bar_o), po); po ^= 1u;
      tc_fence_after();
      uint32_t oa[32], ob[32];''','''      mbar_wait(smem_u32(&bar_o), po); po ^= 1u;
      tc_fence_after();
      if (threadIdx.x == 0) HMA_TL(6, t);
      uint32_t oa[32], ob[32];''')
rep('''          p.lse[((size_t)frame * p.heads + head) * n + qi] = m * p.scale_log2 + log2f(l);
      }
    }
  }
''','''          p.lse[((size_t)frame * p.heads + head) * n + qi] = m * p.scale_log2 + log2f(l);
      }
      if (threadIdx.x == 0) HMA_TL(7, t);
    }
  }
''')
rep('''  tc_fence_before();
  __syncthreads();
  if (warp == kSoftmaxWarps) {
    tc_fence_after();
    tmem_dealloc''','''  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) HMA_TL(10, 0);
  if (warp == kSoftmaxWarps) {
    tc_fence_after();
    tmem_dealloc''')
s=s.rstrip('\n')+'\n\nHMA_DEFINE_TIMELINE_READER(hma_timeline_attn_spatial_fwd)\n'
open(p,'w').write(s)
