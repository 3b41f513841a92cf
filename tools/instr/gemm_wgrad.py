p='hma_b200/csrc/gemm_wgrad.cu'
s=open(p).read()
s=s.replace('''  pdl_wait();
  pdl_launch_dependents();
''','''  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x == 0) HMA_TL(0, 0);
''')
s=s.replace('''        if (++stage == kWStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {''','''        HMA_TL(1, kb);
        if (++stage == kWStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {''')
s=s.replace('''        mbar_wait(smem_u32(&bar_full[stage]), phase);
        tc_fence_after();
        const uint32_t g_addr''','''        mbar_wait(smem_u32(&bar_full[stage]), phase);
        tc_fence_after();
        HMA_TL(2, kb);
        const uint32_t g_addr''')
s=s.replace('''    mbar_wait(smem_u32(&bar_done), 0);
    tc_fence_after();
    const int m = m_blk''','''    mbar_wait(smem_u32(&bar_done), 0);
    tc_fence_after();
    if (threadIdx.x == 128) HMA_TL(3, 0);
    const int m = m_blk''')
s=s.replace('''  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BNW);''','''  if (threadIdx.x == 128) HMA_TL(4, 0);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) HMA_TL(5, 0);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BNW);''')
s=s.rstrip('\n')+'\n\nHMA_DEFINE_TIMELINE_READER(hma_timeline_gemm_wgrad)\n'
open(p,'w').write(s)
