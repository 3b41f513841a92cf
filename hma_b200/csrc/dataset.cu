// Device-resident token dataset (reference hma/data.py:159-294 RawTokenDataset): the memmapped
// token table (video.bin, u16/u32 [num_images, h*w]) and action table (actions/*.bin, f32 [num_images, adim]) live in
// HBM; a batch is a gather of `window` frames `stride` apart per start index, widened to i64 on the way — what
// RawTokenDataset.__getitem__ + torch.stack do on the host, one sample at a time, in the reference's dataloader workers.
// Pure HBM-bound index work: 4 B in, 8 B out per token, coalesced along the h*w axis.
#include "common.cuh"
#include "../../include/hma_b200.h"

namespace hma {

template <typename Tok>
__global__ void __launch_bounds__(256) gather_token_windows_kernel(const Tok* video, const long long* starts, int B, int window,
                                                                   int stride, int frame_elems, long long num_images,
                                                                   long long* out) {
  pdl_wait();
  const long long total = (long long)B * window * frame_elems;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % frame_elems);
    const long long bt = i / frame_elems;
    const int t = (int)(bt % window);
    const long long frame = starts[bt / window] + (long long)t * stride;
    out[i] = frame < num_images ? (long long)video[frame * frame_elems + e] : -1;  // -1 never happens for valid starts
  }
}

__global__ void __launch_bounds__(256) gather_rows_f32_kernel(const float* table, const long long* starts, int B,
                                                              long long row_elems, long long num_rows, long long rows_per_sample,
                                                              float* out) {
  pdl_wait();
  const long long per = rows_per_sample * row_elems;
  const long long total = (long long)B * per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / per, r = i % per;
    const long long src = starts[b] * row_elems + r;
    out[i] = src < num_rows * row_elems ? table[src] : 0.f;
  }
}

}  // namespace hma

extern "C" int hma_gather_token_windows(const void* video, int elem_bytes, long long num_images, const long long* starts, int B,
                                        int window, int stride, int frame_elems, long long* out, void* stream_) {
  using namespace hma;
  if (B == 0) return 0;
  HMA_REQUIRE(window > 0 && stride > 0 && frame_elems > 0 && num_images > 0, "gather_token_windows: bad shape");
  HMA_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "gather_token_windows: token dtype must be uint16 or uint32 (got %d bytes)",
              elem_bytes);
  const long long total = (long long)B * window * frame_elems;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)hma_host::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (elem_bytes == 4)
    HMA_CHECK_CUDA(hma_host::launch_pdl(gather_token_windows_kernel<uint32_t>, dim3((unsigned)blocks), dim3(256), 0, stream,
                                        static_cast<const uint32_t*>(video), starts, B, window, stride, frame_elems, num_images,
                                        out));
  else
    HMA_CHECK_CUDA(hma_host::launch_pdl(gather_token_windows_kernel<uint16_t>, dim3((unsigned)blocks), dim3(256), 0, stream,
                                        static_cast<const uint16_t*>(video), starts, B, window, stride, frame_elems, num_images,
                                        out));
  return 0;
}

extern "C" int hma_gather_rows_f32(const float* table, long long num_rows, long long row_elems, const long long* starts, int B,
                                   long long rows_per_sample, float* out, void* stream_) {
  using namespace hma;
  if (B == 0) return 0;
  HMA_REQUIRE(num_rows > 0 && row_elems > 0 && rows_per_sample > 0, "gather_rows_f32: bad shape");
  const long long total = (long long)B * rows_per_sample * row_elems;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)hma_host::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  HMA_CHECK_CUDA(hma_host::launch_pdl(gather_rows_f32_kernel, dim3((unsigned)blocks), dim3(256), 0,
                                      static_cast<cudaStream_t>(stream_), table, starts, B, row_elems, num_rows, rows_per_sample,
                                      out));
  return 0;
}
