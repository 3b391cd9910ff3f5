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

// ---------------------------------------------------------------------------
// MMA issue/execute rate experiment: one thread issues `iters` x 9 taps x KC/16
// tcgen05.mma (M=128, N=bn) against zero-filled smem, with the A start address either
// fixed (aligned) or shifted by (r*wp+s) rows per tap as the halo-tile conv does.
// out[2*cta] = cycles until the last MMA was issued, out[2*cta+1] = until all retired.
// ---------------------------------------------------------------------------
namespace segb {

__global__ void __launch_bounds__(192, 1)
probe_rate_kernel(int kc, int bn, int b_mn, int wp, int shifted, int iters, int a_mn,
                  long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int swz = kc * 2;
  const int a_rows = 128 + 2 * wp + 2 + 8;
  const int a_bytes = ((a_rows * swz + 1023) / 1024) * 1024;
  const int b_bytes = bn * kc * 2;
  uint8_t* sa = smem;
  uint8_t* sb = smem + a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb + ((b_bytes + 1023) / 1024) * 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  volatile uint32_t* done_flag = reinterpret_cast<volatile uint32_t*>(bars + 3);
  uint8_t* sx = reinterpret_cast<uint8_t*>(bars) + 1024;     // 32 KB scratch for stress copies
  const int flags = a_mn >> 4;                                // stress flags (see probe_rate.py)
  a_mn &= 1;
  long long* scratch = out + 2 * gridDim.x + (size_t)blockIdx.x * 8192;   // 64 KB per CTA
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (a_bytes + b_bytes) / 4; i += blockDim.x) {
    uint32_t v = 0;
    if (flags & 1) {   // pseudo-random bf16 pairs in about [-2, 2]
      uint32_t h = (uint32_t)i * 2654435761u + 12345u;
      h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
      v = (h & 0x807F807Fu) | 0x3F003F00u;
    }
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    *done_flag = 0;
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    // warp-uniform branch, one elected lane issues (same shape as the production loops)
    const uint32_t idesc = umma_idesc_bf16(128, bn, a_mn, b_mn);
    const uint32_t hiA = umma_desc_hi(8 * swz, swz);
    const int atom_n = bn < 64 ? bn : 64;
    const int atom_bytes = atom_n * 2;
    const uint32_t hiB = b_mn ? umma_desc_hi(8 * atom_bytes, atom_bytes) : umma_desc_hi(8 * swz, swz);
    const uint32_t lboB = b_mn ? kc * atom_bytes : 0;
    const uint32_t kstepB = b_mn ? (16 * atom_bytes) >> 4 : 2;
    // a_mn: A is [K rows = pixels][M cols = channels], atom = swz bytes of M; LBO = one row
    // (tap-shifted atoms), K step = 16 rows
    const uint32_t lboA = a_mn ? swz : 0;
    const uint32_t kstepA = a_mn ? (16 * swz) >> 4 : 2;
    const uint32_t a0 = umma_desc_lo(smem_u32(sa), lboA);
    const uint32_t b0 = umma_desc_lo(smem_u32(sb), lboB);
    const uint32_t row16 = swz >> 4;
    const uint32_t sh = shifted ? 1u : 0u;
    const uint32_t d1 = sh * row16, dw = sh * (uint32_t)wp * row16;
    long long t0 = 0, t1 = 0, t2 = 0;
    if (elect_one()) {
      t0 = clock64();
      uint32_t acc = 0;
      if (flags & 32) {
        // rolled variant: the tap loop is NOT unrolled, so every tap re-writes the same
        // uniform registers (as a generic runtime-kh/kw loop does)
        for (int it = 0; it < iters; ++it) {
          uint32_t a_row = a0;
#pragma unroll 1
          for (int r = 0; r < 3; ++r, a_row += dw) {
#pragma unroll 1
            for (int s = 0; s < 3; ++s) {
              const uint32_t a_tap = a_row + s * d1;
#pragma unroll 1
              for (int kk = 0; kk < kc / 16; ++kk) {
                umma_f16(tmem_base, umma_desc_pack(hiA, a_tap + kk * kstepA),
                         umma_desc_pack(hiB, b0 + kk * kstepB), idesc, acc);
                acc = 1;
              }
            }
          }
        }
      } else if (flags & 64) {
        // tap loop rolled, 4 k-steps unrolled (the production hconv loop shape)
        for (int it = 0; it < iters; ++it) {
          uint32_t a_row = a0;
#pragma unroll 1
          for (int r = 0; r < 3; ++r, a_row += dw) {
#pragma unroll 1
            for (int s = 0; s < 3; ++s) {
              const uint32_t a_tap = a_row + s * d1;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                if (kk < kc / 16) {
                  umma_f16(tmem_base, umma_desc_pack(hiA, a_tap + kk * kstepA),
                           umma_desc_pack(hiB, b0 + kk * kstepB), idesc, acc);
                  acc = 1;
                }
              }
            }
          }
        }
      } else {
      for (int it = 0; it < iters; ++it) {
          uint32_t a_row = a0;
  #pragma unroll
          for (int r = 0; r < 3; ++r, a_row += dw) {
  #pragma unroll
            for (int s = 0; s < 3; ++s) {
              const uint32_t a_tap = a_row + s * d1;
              if (kc == 64) {
  #pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  umma_f16(tmem_base, umma_desc_pack(hiA, a_tap + kk * kstepA),
                           umma_desc_pack(hiB, b0 + kk * kstepB), idesc, acc);
                  acc = 1;
                }
              } else if (kc == 32) {
  #pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                  umma_f16(tmem_base, umma_desc_pack(hiA, a_tap + kk * kstepA),
                           umma_desc_pack(hiB, b0 + kk * kstepB), idesc, acc);
                  acc = 1;
                }
              } else {
                umma_f16(tmem_base, umma_desc_pack(hiA, a_tap), umma_desc_pack(hiB, b0), idesc, acc);
                acc = 1;
              }
            }
          }
        }
      }
      umma_commit(&bars[0]);
      t1 = clock64();
    }
    __syncwarp();
    mbar_wait(&bars[0], 0);
    t2 = clock64();
    t0 = __shfl_sync(0xffffffffu, t0, 0) | 0;
    if (t1 != 0) {
      out[2 * blockIdx.x] = t1 - t0;
      out[2 * blockIdx.x + 1] = t2 - t0;
      *done_flag = 1;
    }
  } else if (warp == 1) {
    if ((flags & 2) && lane == 0) {     // bulk global -> smem copies, 16 KB each, back to back
      uint32_t ph = 0;
      int slot = 0;
      while (!*done_flag) {
        mbar_expect_tx(&bars[1], 16384);
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            ::"r"(smem_u32(sx + slot * 16384)), "l"(scratch), "r"(16384), "r"(smem_u32(&bars[1]))
            : "memory");
        mbar_wait(&bars[1], ph);
        ph ^= 1u;
        slot ^= 1;
      }
    }
  } else {
    if (flags & 4) {                    // epilogue-like TMEM reads of another accumulator stage
      uint32_t r[32];
      uint32_t sink = 0;
      while (!*done_flag) {
        tmem_ld_32x32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256, r);
        tmem_ld_wait();
        sink += r[lane & 31];
      }
      if (sink == 0x12345678u) scratch[lane] = sink;
    }
    if (flags & 128) {
      // TMEM drain rate: every warp 2..5 (one per lane quadrant) performs 512 x32 loads
      // (4 KB each), two in flight, and reports its cycles in out[2*cta + 64 + warp]
      uint32_t r0[32], r1[32];
      uint32_t sink = 0;
      const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256;
      const long long c0 = clock64();
      tmem_ld_32x32(ta, r0);
      for (int k = 0; k < 256; ++k) {
        tmem_ld_wait();
        tmem_ld_32x32(ta + 32, r1);
        sink += r0[k & 31];
        tmem_ld_wait();
        tmem_ld_32x32(ta, r0);
        sink += r1[k & 31];
      }
      tmem_ld_wait();
      const long long c1 = clock64();
      if (lane == 0) out[2 * gridDim.x + 64 + warp] = c1 - c0;
      if (sink == 0x12345678u) scratch[lane] = sink;
    }
    if (flags & 8) {                    // row-strided 16-byte global stores (uncoalesced epilogue)
      uint4* dst = reinterpret_cast<uint4*>(scratch) + (warp - 2) * 2048;
      int k = 0;
      while (!*done_flag) {
        dst[lane * 8 + (k & 7)] = make_uint4(k, k, k, k);
        ++k;
      }
    }
    if (flags & 16) {                   // broadcast global loads (bias-style)
      float acc = 0.f;
      int k = 0;
      while (!*done_flag) {
        acc += __ldg(reinterpret_cast<const float*>(scratch) + (k & 63));
        ++k;
      }
      if (acc == 1.2345f) scratch[lane] = 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

int umma_probe_rate(int kc, int bn, int b_mn, int wp, int shifted, int iters, int a_mn, int ctas,
                    long long* out, cudaStream_t st) {
  SEG_REQUIRE((kc == 16 || kc == 32 || kc == 64) && bn >= 16 && bn <= 256 && bn % 16 == 0 &&
                  wp >= 1 && wp <= 256 && ctas >= 1,
              SEG_E_BAD_SHAPE, "probe_rate: bad argument");
  const int a_rows = 128 + 2 * wp + 2 + 8;
  const int smem = ((a_rows * kc * 2 + 1023) / 1024) * 1024 + ((bn * kc * 2 + 1023) / 1024) * 1024 +
                   2048 + 1024 + 32768;
  SEG_CHECK_CUDA(cudaFuncSetAttribute(probe_rate_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  probe_rate_kernel<<<ctas, 192, smem, st>>>(kc, bn, b_mn, wp, shifted, iters, a_mn, out);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

// ---------------------------------------------------------------------------
// fp32 global-reduction rate probe (tools/probe_red.py): every CTA adds `elems` floats
// (rows of 64) from shared memory into dst + (cta % regions) * elems.
//   mode 0: red.global.add.v4.f32, warp-coalesced (512 contiguous bytes per instruction)
//   mode 1: red.global.add.v4.f32, thread = row of 64 floats (the weight-gradient pattern)
//   mode 2: cp.reduce.async.bulk .add.f32, one 256-byte row per operation, thread = row
//   mode 3: cp.reduce.async.bulk .add.f32, `op_bytes` per operation, issued by warp 0
// out[2*cta] = cycles until the last operation was issued, out[2*cta+1] = until complete
// (bulk modes; the red modes cannot observe completion).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
probe_red_kernel(int mode, int elems, int regions, int op_bytes, float* dst, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  float* S = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  for (int i = threadIdx.x; i < elems; i += blockDim.x) S[i] = 1.0f;
  fence_proxy_async();
  __syncthreads();
  float* g = dst + (size_t)(blockIdx.x % regions) * elems;
  const int rows = elems / 64;
  const long long t0 = clock64();
  if (mode == 0) {
    for (int i = threadIdx.x * 4; i < elems; i += blockDim.x * 4)
      red_add_v4(g + i, S[i], S[i + 1], S[i + 2], S[i + 3]);
  } else if (mode == 1) {
    for (int r = threadIdx.x; r < rows; r += blockDim.x)
#pragma unroll
      for (int j = 0; j < 64; j += 4) {
        const int i = r * 64 + j;
        red_add_v4(g + i, S[i], S[i + 1], S[i + 2], S[i + 3]);
      }
  } else if (mode == 2) {
    for (int r = threadIdx.x; r < rows; r += blockDim.x)
      bulk_reduce_add_f32(g + r * 64, smem_u32(S + r * 64), 256);
    bulk_commit_group();
  } else {
    if (threadIdx.x < 32) {
      const int ops = elems * 4 / op_bytes;
      for (int o = threadIdx.x; o < ops; o += 32)
        bulk_reduce_add_f32(g + (size_t)o * (op_bytes / 4), smem_u32(S) + o * op_bytes, op_bytes);
      bulk_commit_group();
    }
  }
  const long long t1 = clock64();
  if (mode >= 2) bulk_wait_group<0>();
  __syncthreads();
  const long long t2 = clock64();
  if (threadIdx.x == 0) {
    out[2 * blockIdx.x] = t1 - t0;
    out[2 * blockIdx.x + 1] = t2 - t0;
  }
}

int probe_red_rate(int mode, int ctas, int elems, int regions, int op_bytes, float* dst,
                   long long* out, cudaStream_t st) {
  SEG_REQUIRE(ctas >= 1 && elems >= 64 && elems % 64 == 0 && elems * 4 <= 200 * 1024 &&
                  regions >= 1 && op_bytes >= 16 && op_bytes % 16 == 0 && (elems * 4) % op_bytes == 0,
              SEG_E_BAD_SHAPE, "probe_red_rate: bad argument");
  SEG_CHECK_CUDA(cudaFuncSetAttribute(probe_red_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  probe_red_kernel<<<ctas, 128, elems * 4 + 256, st>>>(mode, elems, regions, op_bytes, dst, out);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

}  // namespace segb

