// C ABI of the self-test / micro-benchmark hooks (include/segb200_probes.h).  Linked only
// into libsegb200_probes.so, next to the same kernel objects as the product library: the
// product library libsegb200.so exports no probe.
#include "common.cuh"
#include "../../include/segb200_probes.h"

namespace segb {
int umma_probe(int mode, int M, int N, int K, const void* a, const void* b, float* d,
               cudaStream_t st);
int umma_probe_shift(int K, const void* a, const void* b, int shift, int use_bo, int split,
                     float* d, cudaStream_t st);
int probe_red_rate(int mode, int ctas, int elems, int regions, int op_bytes, float* dst,
                   long long* out, cudaStream_t st);
int umma_probe_rate(int kc, int bn, int b_mn, int wp, int shifted, int iters, int a_mn, int ctas,
                    long long* out, cudaStream_t st);
}  // namespace segb

using namespace segb;

extern "C" {

SEG_API int32_t seg_probe_umma(int32_t mode, int32_t m, int32_t n, int32_t k, const void* a,
                       const void* b, float* d, void* stream) {
  SEG_REQUIRE(a && b && d && m > 0 && n % 16 == 0 && k % 16 == 0, SEG_E_BAD_SHAPE,
              "probe_umma: bad argument");
  if (mode & 0x100)   // row-shift experiment: bits 16..23 = shift, bit 9 = base_offset field,
                      // bits 24..31 = split row of a two-box load (0 = single box)
    return umma_probe_shift(k, a, b, (mode >> 16) & 0xff, (mode >> 9) & 1, (mode >> 24) & 0xff, d,
                            (cudaStream_t)stream);
  return umma_probe(mode, m, n, k, a, b, d, (cudaStream_t)stream);
}

SEG_API int32_t seg_probe_red_rate(int32_t mode, int32_t ctas, int32_t elems, int32_t regions,
                           int32_t op_bytes, float* dst, int64_t* out, void* stream) {
  SEG_REQUIRE(dst && out, SEG_E_BAD_SHAPE, "probe_red_rate: null argument");
  return probe_red_rate(mode, ctas, elems, regions, op_bytes, dst,
                        reinterpret_cast<long long*>(out), (cudaStream_t)stream);
}

SEG_API int32_t seg_probe_mma_rate(int32_t kc, int32_t bn, int32_t b_mn, int32_t wp, int32_t shifted,
                           int32_t iters, int32_t a_mn, int32_t ctas, int64_t* out, void* stream) {
  SEG_REQUIRE(out, SEG_E_BAD_SHAPE, "probe_mma_rate: null out");
  return umma_probe_rate(kc, bn, b_mn, wp, shifted, iters, a_mn, ctas,
                         reinterpret_cast<long long*>(out), (cudaStream_t)stream);
}

}  // extern "C"
