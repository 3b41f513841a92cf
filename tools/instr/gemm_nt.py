p='hma_b200/csrc/gemm_nt.cu'
s=open(p).read()
s=s.replace('''  pdl_wait();               // everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
''','''  pdl_wait();               // everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  if (threadIdx.x == 0) HMA_TL(0, 0);
''')
s=s.replace('''        mbar_wait(smem_u32(&bar_tempty[as]), aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem''','''        HMA_TL(1, it);
        mbar_wait(smem_u32(&bar_tempty[as]), aph ^ 1u);
        tc_fence_after();
        HMA_TL(2, it);
        const uint32_t d_tmem''')
s=s.replace('''        umma_commit(smem_u32(&bar_tfull[as]));
      }''','''        umma_commit(smem_u32(&bar_tfull[as]));
        HMA_TL(3, it);
      }''')
s=s.replace('''      mbar_wait(smem_u32(&bar_tfull[as]), aph);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < kChunks; ++j) {''','''      if (threadIdx.x == 128) HMA_TL(4, it);
      mbar_wait(smem_u32(&bar_tfull[as]), aph);
      tc_fence_after();
      if (threadIdx.x == 128) HMA_TL(5, it);
#pragma unroll
      for (int j = 0; j < kChunks; ++j) {''')
s=s.replace('''      tc_fence_before();
      mbar_arrive(smem_u32(&bar_tempty[as]));
    }
  }
''','''      tc_fence_before();
      mbar_arrive(smem_u32(&bar_tempty[as]));
      if (threadIdx.x == 128) HMA_TL(6, it);
    }
  }
''')
s=s.replace('''  tc_fence_before();
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
