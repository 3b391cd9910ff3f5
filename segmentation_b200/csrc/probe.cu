// Descriptor experiments (test hooks, not on the hot path): does a K-major UMMA
// A-descriptor whose start address is shifted by `shift` rows inside a swizzled
// tile read rows [shift, shift+128)?  This decides whether one staged halo tile
// can serve all filter taps of a 3x3 convolution.
#include "umma_conv.cuh"

namespace segb {

int make_probe_tmap(CUtensorMap* tm, const void* ptr, int64_t cols, int64_t rows, int box_cols,
                    int box_rows, int swizzle_bytes);

template <int KC>
__global__ void __launch_bounds__(128, 1)
probe_shift_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmA2, int split, int shift,
                   int use_base_offset, float* d_out) {
  constexpr int ROWS = 160;            // staged A rows (>= 128 + max shift)
  constexpr int BN = 64;
  constexpr int SWZ = KC * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((ROWS * SWZ + 1023) / 1024) * 1024;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb + BN * SWZ);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<64>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], ROWS * SWZ + BN * SWZ);
    if (split > 0) {
      // two boxes: the second one lands at a smem address that is NOT aligned to the
      // swizzle repeat -> tests whether TMA's swizzle is a function of the smem address
      tma_load_2d(&tmA, &bars[0], sa, 0, 0);
      tma_load_2d(&tmA2, &bars[0], sa + split * SWZ, 0, split);
    } else {
      tma_load_2d(&tmA, &bars[0], sa, 0, 0);
    }
    tma_load_2d(&tmB, &bars[0], sb, 0, 0);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
#pragma unroll
    for (int kk = 0; kk < KC / 16; ++kk) {
      const uint32_t a_addr = smem_u32(sa) + shift * SWZ + kk * 32;
      uint64_t da = umma_smem_desc(a_addr, 0, 8 * SWZ, SWZ);
      if (use_base_offset) {
        // base offset = row phase of the start address inside the swizzle repeat
        const uint64_t bo = (a_addr / SWZ) & 7u;
        da |= bo << 49;
      }
      const uint64_t db = umma_smem_desc(smem_u32(sb) + kk * 32, 0, 8 * SWZ, SWZ);
      umma_f16(tmem_base, da, db, idesc, kk ? 1u : 0u);
    }
    umma_commit(&bars[1]);
  }
  __syncthreads();
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  uint32_t r[32];
  for (int cc = 0; cc < BN; cc += 32) {
    tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + cc, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j)
      d_out[(warp * 32 + lane) * BN + cc + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<64>(tmem_base);
  }
}

template <int KC>
static int run_probe_shift(const void* a, const void* b, int shift, int use_bo, int split,
                           float* d, cudaStream_t st) {
  CUtensorMap tmA, tmB, tmA2;
  int rc = make_probe_tmap(&tmA, a, KC, 160, KC, split > 0 ? split : 160, KC * 2);
  if (rc) return rc;
  rc = make_probe_tmap(&tmA2, a, KC, 160, KC, split > 0 ? 160 - split : 160, KC * 2);
  if (rc) return rc;
  rc = make_probe_tmap(&tmB, b, KC, 64, KC, 64, KC * 2);
  if (rc) return rc;
  const int smem = 160 * KC * 2 + 64 * KC * 2 + 4096;
  SEG_CHECK_CUDA(cudaFuncSetAttribute(probe_shift_kernel<KC>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_shift_kernel<KC><<<1, 128, smem, st>>>(tmA, tmB, tmA2, split, shift, use_bo, d);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

// a: [160][K] bf16, b: [64][K] bf16 (K-major), d: [128][64] fp32 = a[shift:shift+128] @ b^T
int umma_probe_shift(int K, const void* a, const void* b, int shift, int use_bo, int split,
                     float* d, cudaStream_t st) {
  SEG_REQUIRE(shift >= 0 && shift <= 32, SEG_E_BAD_SHAPE, "probe_shift: shift out of range");
  switch (K) {
    case 64: return run_probe_shift<64>(a, b, shift, use_bo, split, d, st);
    case 32: return run_probe_shift<32>(a, b, shift, use_bo, split, d, st);
    case 16: return run_probe_shift<16>(a, b, shift, use_bo, split, d, st);
  }
  set_error("probe_shift: K must be 16, 32 or 64");
  return SEG_E_UNSUPPORTED;
}

}  // namespace segb
