// Shared host/device helpers for the segb200 library.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/segb200.h"

namespace segb {

typedef __nv_bfloat16 bf16;

// thread-local last error text (seg_last_error_string)
void set_error(const char* fmt, ...);

#define SEG_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      segb::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(_e), \
                      cudaGetErrorString(_e));                                            \
      return SEG_E_CUDA;                                                                  \
    }                                                                                     \
  } while (0)

#define SEG_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      segb::set_error(__VA_ARGS__);    \
      return (code);                   \
    }                                  \
  } while (0)

#define SEG_LAUNCH_CHECK() SEG_CHECK_CUDA(cudaGetLastError())

// Launch with the programmatic-stream-serialization attribute (PDL) when enabled
// (seg_set_option key 7): the kernel may start while its predecessor in the stream is
// still running; every kernel launched through here calls pdl_wait() before it touches
// global memory.
extern int g_pdl;
// the kernel most recently launched through launch_k / launch_kc on this thread
// (seg_last_kernel_name: which tile kernel / instantiation a call was planned onto)
extern thread_local const void* g_last_kernel_fn;
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kc(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                    cudaStream_t st, int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (g_pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  g_last_kernel_fn = reinterpret_cast<const void*>(kern);
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                   cudaStream_t st, Args&&... args) {
  return launch_kc(kern, grid, block, smem, st, 1, static_cast<Args&&>(args)...);
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
// seg_set_option key 17: SMs the persistent tile kernels size their grids for (0 = all).
// Data-parallel training leaves a few SMs to the all-reduce kernels: a persistent grid of one
// CTA per SM that finds some SMs held by NCCL for the length of a bucket runs a second wave.
extern int g_sm_limit;
static inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return (g_sm_limit > 0 && g_sm_limit < n) ? g_sm_limit : n;
}

__device__ __forceinline__ const bf16* view_at(const seg_view& v, int n, int y, int x) {
  return reinterpret_cast<const bf16*>(v.ptr) + n * v.sn + y * v.sh + x * v.sw;
}
__device__ __forceinline__ bf16* view_at_mut(const seg_view& v, int n, int y, int x) {
  return reinterpret_cast<bf16*>(v.ptr) + n * v.sn + y * v.sh + x * v.sw;
}

static inline bool view_dense(const seg_view& v) {
  return v.sw == v.c && v.sh == (int64_t)v.w * v.c && v.sn == (int64_t)v.h * v.w * v.c;
}

static inline seg_view null_view() {
  seg_view v;
  v.ptr = nullptr;
  v.n = v.h = v.w = v.c = 0;
  v.sn = v.sh = v.sw = 0;
  return v;
}

}  // namespace segb
