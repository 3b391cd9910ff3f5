/* segb200_probes.h - self-test and micro-benchmark hooks of the B200 segmentation kernels.
 *
 * NOT part of the product ABI (include/segb200.h): these entry points live in
 * segmentation_b200/libsegb200_probes.so, which links the same kernel objects as the product
 * library plus csrc/probe.cu + csrc/probe_api.cu.  tests/ use seg_probe_umma to check the
 * tcgen05 shared-memory / instruction descriptors against plain matrix products; tools/ use
 * the rate probes behind the design decisions recorded in profiles/.
 */
#ifndef SEGB200_PROBES_H_
#define SEGB200_PROBES_H_

#include "segb200.h"

#ifdef __cplusplus
extern "C" {
#endif

SEG_API int32_t seg_probe_umma(int32_t mode, int32_t m, int32_t n, int32_t k, const void* a,
                       const void* b, float* d, void* stream);
/* MMA issue/retire rate of `iters` x 9 taps x kc/16 tcgen05.mma (M=128, N=bn) per CTA;
 * out[2*cta] = cycles to issue, out[2*cta+1] = cycles to retire (tools/probe_rate.py) */
SEG_API int32_t seg_probe_mma_rate(int32_t kc, int32_t bn, int32_t b_mn, int32_t wp, int32_t shifted,
                           int32_t iters, int32_t a_mn, int32_t ctas, int64_t* out, void* stream);

/* fp32 global-reduction rate: `ctas` CTAs each add `elems` floats from shared memory into
 * dst + (cta % regions) * elems with red.global.v4 (mode 0 coalesced, 1 row-per-thread) or
 * cp.reduce.async.bulk (mode 2: 256-byte rows, 3: op_bytes per operation); out[2*cta] /
 * out[2*cta+1] = cycles to issue / to complete (tools/probe_red.py) */
SEG_API int32_t seg_probe_red_rate(int32_t mode, int32_t ctas, int32_t elems, int32_t regions,
                           int32_t op_bytes, float* dst, int64_t* out, void* stream);


/* Exported by PROFILING builds of the product library only (SEGB200_KERNEL_PROF=1 python -m
 * segmentation_b200.build, tools/layer_prof.py): device buffer of 3*16*4 int64 that CTA 0 of
 * the halo / spatial-tile conv kernels fills with clock64() marks per role (producer / MMA
 * issuer / epilogue) and tile; null disables.  The default build compiles the marks out. */
int32_t seg_debug_prof_buffer(void* device_buf);

#ifdef __cplusplus
}
#endif
#endif /* SEGB200_PROBES_H_ */
