// Host-side plumbing shared by every entry point: error string, TMA descriptor encoding
// (through the driver entry point, so the library does not link libcuda), SM count.
#include <cstdlib>
#include "common.cuh"
#include "../../include/hma_b200.h"

#include <cstdarg>
#include <cstdio>
#include <mutex>

namespace hma_host {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                              CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                              CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeFn>(p);
    }
  });
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  return make_tmap_bf16_2d_sw(map, base, inner, outer, row_stride_bytes, box_inner, box_outer, 128);
}

int make_tmap_bf16_2d_sw(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer,
                         uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
  EncodeFn enc = get_encode();
  HMA_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  HMA_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16-byte aligned");
  HMA_REQUIRE((row_stride_bytes & 15) == 0, "TMA row stride must be a multiple of 16 bytes");
  HMA_REQUIRE(swizzle_bytes == 128 || swizzle_bytes == 64, "unsupported swizzle %d", swizzle_bytes);
  HMA_REQUIRE((int)box_inner * 2 == swizzle_bytes, "inner box must span exactly one swizzle row");
  HMA_REQUIRE(box_outer >= 1 && box_outer <= 256, "TMA box rows out of range");
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstr[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  HMA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// 3-D bf16 map with 64-byte swizzle over a row-major [d2][d1][d0] view (d0 contiguous); box {32, b1, b2}.
int make_tmap_bf16_3d_sw64(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                           uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b1, uint32_t b2) {
  EncodeFn enc = get_encode();
  HMA_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  HMA_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (stride1_bytes & 15) == 0 && (stride2_bytes & 15) == 0,
              "TMA base/strides must be 16-byte aligned");
  HMA_REQUIRE(b1 >= 1 && b1 <= 256 && b2 >= 1 && b2 <= 256, "TMA box out of range");
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstr[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {32, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  HMA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed with CUresult %d", (int)r);
  return 0;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("HMA_B200_NO_PDL");
    v = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  return dev;
}

int sm_count() {
  static int n[64] = {};
  const int dev = current_device() & 63;
  if (n[dev] == 0) {
    if (cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n[dev] = 148;
  }
  return n[dev];
}

}  // namespace hma_host

extern "C" {
int hma_abi_version(void) { return HMA_B200_ABI_VERSION; }
const char* hma_last_error(void) { return hma_host::last_error(); }
int hma_device_check(void) {
  int dev = 0;
  HMA_CHECK_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  HMA_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  HMA_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  HMA_REQUIRE(major == 10, "hma_b200 kernels are sm_100a only; device is sm_%d%d", major, minor);
  return 0;
}
}
