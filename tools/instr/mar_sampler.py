"""Source patcher: clock64 timeline stamps for csrc/mar_sampler.cu (CTA 0). Apply, build with HMA_B200_TIMELINE=1, run
`python tools/timeline.py mar_sampler`, then `git checkout hma_b200/csrc/mar_sampler.cu`.
Events per GEMM stage (slot = stage index): 1/2 producer's first / last issue, 3/8/4 k-block 0 / 1 / 15 seen full by the MMA
thread, 5/6 epilogue sees the accumulator / has stored; per grid barrier (slot = epoch): 7 entered, 0 left."""
p = 'hma_b200/csrc/mar_sampler.cu'
s = open(p).read()


def rep(a, b):
    global s
    assert s.count(a) == 1, a[:70]
    s = s.replace(a, b)


rep('''  uint32_t tphase = 0;  // accumulator hand-over parity (one tile at a time)''',
    '''  uint32_t tphase = 0;  // accumulator hand-over parity (one tile at a time)
  int slot = 0;''')
rep('''        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(smem_u32(&sh.empty[ps.stage]), ps.phase ^ 1u);
          const uint32_t full = smem_u32(&sh.full[ps.stage]);''', '''        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(smem_u32(&sh.empty[ps.stage]), ps.phase ^ 1u);
          if (kb == 0) HMA_TL(1, ps.slot);
          if (kb == KB - 1) HMA_TL(2, ps.slot);
          const uint32_t full = smem_u32(&sh.full[ps.stage]);''')
rep('''        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(smem_u32(&sh.full[ps.stage]), ps.phase);
          tc_fence_after();''', '''        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(smem_u32(&sh.full[ps.stage]), ps.phase);
          tc_fence_after();
          if (kb == 0) HMA_TL(3, ps.slot);
          if (kb == 1) HMA_TL(8, ps.slot);
          if (kb == KB - 1) HMA_TL(4, ps.slot);''')
rep('''      mbar_wait(smem_u32(&sh.tfull), ps.tphase);
      tc_fence_after();
      uint32_t r[32];''', '''      mbar_wait(smem_u32(&sh.tfull), ps.tphase);
      tc_fence_after();
      if (threadIdx.x == 64) HMA_TL(5, ps.slot);
      uint32_t r[32];''')
rep('''        for (int q = 0; q < 4; ++q) dst[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      }
    }
  }
}''', '''        for (int q = 0; q < 4; ++q) dst[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      }
      if (threadIdx.x == 64) HMA_TL(6, ps.slot);
    }
  }
  ps.slot++;
}''')
rep('''__device__ __forceinline__ void grid_sync(unsigned* ctr, unsigned& epoch) {
  fence_proxy_async_all();''', '''__device__ __forceinline__ void grid_sync(unsigned* ctr, unsigned& epoch) {
  if (threadIdx.x == 0) HMA_TL(7, (int)epoch);
  fence_proxy_async_all();''')
rep('''    __threadfence();
  }
  __syncthreads();
  fence_proxy_async_all();
}''', '''    __threadfence();
  }
  __syncthreads();
  fence_proxy_async_all();
  if (threadIdx.x == 0) HMA_TL(0, (int)epoch);
}''')
s = s.rstrip('\n') + '\n\nHMA_DEFINE_TIMELINE_READER(hma_timeline_mar_sampler)\n'
open(p, 'w').write(s)
