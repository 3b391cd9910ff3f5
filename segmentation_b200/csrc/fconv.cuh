// First-layer tcgen05 convolution (3x3, stride 1, RGB input): forward and weight gradient.
//
// The first layer of every reference graph convolves 3 input channels (models/unet.py:111,
// models/fcn.py:110): arithmetic intensity 25 flop/B, i.e. HBM-bound, and with the input
// padded to 16 channels for the generic spatial-tile kernels it was also TMA-row bound
// (32-byte pixel rows) - 10 % of the U-Net step for 0.5 % of its FLOPs.  Here the input
// lives in HBM as 4 bf16 per pixel, (R, G, B, 1): 8 bytes.  Four builder warps gather the
// 3x3x4 patch of each of a tile's 128 consecutive output pixels with nine 8-byte cp.async
// copies per pixel (no register staging, so several tiles are in flight per CTA; L1/L2 serve
// the 9-fold reuse) straight into the canonical K-major SWIZZLE_128B operand layout in
// shared memory - one 128-byte row per pixel, k-slot = r*12 + s*4 + c (c = 3: weight rows
// zero), slots 36 and 37 a constant 1 whose weight rows hold the bias split into two bf16
// (hi + lo: the bias add costs the epilogue nothing) - so a tile is ONE 128 x BN x 48 MMA
// group and nothing padded is ever read from or written to HBM.  Output pixels are
// flattened over (n, y, x), so an output tile is 128 consecutive rows of the dense
// [pixels][BN] activation: TMEM -> relu + bf16 (one cvt per pair) -> one TMA store box per
// epilogue warp; eight epilogue warps, the two of a TMEM lane quadrant take tiles alternately.
//
// Weight gradient: the same patch tile read MN-major (pixels are the GEMM K axis) times the
// dZ tile fetched by one TMA load, accumulated in TMEM over all of the CTA's tiles:
// dW[k-slot][co] += sum_px patch[px][k-slot] * dZ[px][co].  The constant-1 slot 36
// accumulates sum_px dZ - the bias gradient - for free.  One red.add of 27 x BN (+ BN)
// floats per CTA at the end.
#pragma once
#include "umma_conv.cuh"

namespace segb {

struct FconvParams {
  const uint2* x4;          // [N][H][W] pixels of 4 bf16 (R, G, B, 1), dense
  int H, W, Ho, Wo;
  int pad_t, pad_l;
  int M_total;              // N * Ho * Wo (< 2^30)
  int tiles;                // ceil(M_total / 128)
  uint32_t div_wo_mul, div_wo_shr;      // m / Wo and m / (Ho*Wo) as __umulhi(m, mul) >> shr
  uint32_t div_hw_mul, div_hw_shr;
  const bf16* w;            // bf16 shadow [3][3][cin_pad][cout_pad]
  int cin_pad, cout_pad, cout;
  const float* bias;
  int flags;
  float* dw;                // fp32 master layout [3][3][3][cout]
  float* db;                // nullable
};

constexpr int kFcThreads = 448;          // MMA issuer, TMA/alloc warp, 8 epilogue, 4 builder warps
constexpr int kFcStages = 4;
constexpr int kFcABytes = 128 * 128;     // one patch tile: 128 pixels x 128-byte row
constexpr int kFcRows = 37;              // accumulator rows of the weight gradient that are used

template <int BN, bool WGRAD>
struct FconvCfg {
  static constexpr int rowB = BN * 2;
  static constexpr int kZBytes = 128 * rowB;            // one dZ tile (wgrad)
  static constexpr int kWBytes = BN * 128;              // weights, K-major 128-byte rows (fwd)
  static constexpr int kStgBytes = 32 * rowB;           // one epilogue warp's store box (fwd)
  static constexpr int kOffBars = kFcStages * kFcABytes +
                                  (WGRAD ? kFcStages * kZBytes : kWBytes + 16 * kStgBytes);
  static constexpr int kSmemBytes = kOffBars + 256 + 1024 /*base alignment*/;
  static constexpr int kTmemCols = WGRAD ? (BN < 32 ? 32 : BN) : 2 * BN;
};

template <int BN, bool WGRAD>
__global__ void __launch_bounds__(kFcThreads, 2)
fconv_kernel(const __grid_constant__ CUtensorMap tmIO, const FconvParams P) {
  using Cfg = FconvCfg<BN, WGRAD>;
  constexpr int S = kFcStages;
  constexpr int rowB = Cfg::rowB;
  static_assert(BN == 32 || BN == 64, "first-layer kernel: 32 or 64 output channels per tile");

  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* a_ring = smem;
  uint8_t* z_ring = smem + S * kFcABytes;               // wgrad
  uint8_t* w_smem = smem + S * kFcABytes;               // fwd
  uint8_t* stg = w_smem + Cfg::kWBytes;                 // fwd
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBars);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + S;
  uint64_t* z_full = bars + 2 * S;
  uint64_t* tfull = bars + 3 * S;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_my = ((int)blockIdx.x < P.tiles)
                       ? (P.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmIO);
    for (int i = 0; i < S; ++i) {
      mbar_init(&a_full[i], 128);
      mbar_init(&a_empty[i], 1);
      mbar_init(&z_full[i], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  // initialise the patch ring once: the builders only ever write the first 72 bytes of a row
  // (k-slots 0..35); slots 36 and 37 are the constant 1, slots 38..47 (read by the third
  // k-step) stay zero
  for (int i = threadIdx.x; i < S * kFcABytes / 16; i += kFcThreads)
    sts128(smem_u32(a_ring) + i * 16, make_uint4(0u, 0u, 0u, 0u));
  __syncthreads();
  for (int i = threadIdx.x; i < S * 128; i += kFcThreads) {
    const uint32_t row = smem_u32(a_ring) + (uint32_t)i * 128u;
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(row + ((4u ^ ((uint32_t)i & 7u)) << 4) + 8u),
                 "r"(0x3F803F80u) : "memory");
  }
  pdl_wait();                 // everything above overlaps the previous kernel's tail
  if (!WGRAD) {
    // weights -> K-major SWIZZLE_128B rows: row = output channel, k-slot = r*12 + s*4 + c
    for (int idx = threadIdx.x; idx < BN * 64; idx += kFcThreads) {
      const int co = idx >> 6, k = idx & 63;
      const int r = k / 12, rem = k - r * 12, s = rem >> 2, c = rem & 3;
      bf16 v = __float2bfloat16(0.f);
      if (k < 36 && c < 3 && co < P.cout_pad) {
        v = P.w[((int64_t)(r * 3 + s) * P.cin_pad + c) * P.cout_pad + co];
      } else if ((k == 36 || k == 37) && (P.flags & SEG_EPI_BIAS) && co < P.cout) {
        const float bv = __ldg(P.bias + co);
        const bf16 hi = __float2bfloat16(bv);
        v = k == 36 ? hi : __float2bfloat16(bv - __bfloat162float(hi));
      }
      *reinterpret_cast<bf16*>(w_smem + co * 128 + (((k >> 3) ^ (co & 7)) << 4) + (k & 7) * 2) = v;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================ MMA issuer ============================
    if (elect_one()) {
      if (!WGRAD) {
        constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BN, 0, 0);
        constexpr uint32_t hi = umma_desc_hi(8 * 128, 128);
        const uint32_t b0 = umma_desc_lo(smem_u32(w_smem), 0);
        for (int i = 0; i < n_my; ++i) {
          const int s = i % S, as = i & 1;
          mbar_wait(&tempty[as], ((uint32_t)(i >> 1) & 1u) ^ 1u);
          mbar_wait(&a_full[s], (uint32_t)(i / S) & 1u);
          fence_proxy_async();           // the builders' cp.async writes -> tensor-core reads
          tc_fence_after();
          const uint32_t a0 = umma_desc_lo(smem_u32(a_ring + s * kFcABytes), 0);
#pragma unroll
          for (int kk = 0; kk < 3; ++kk)
            umma_f16(tmem_base + as * BN, umma_desc_pack(hi, a0 + kk * 2),
                     umma_desc_pack(hi, b0 + kk * 2), idesc, kk != 0 ? 1u : 0u);
          umma_commit(&a_empty[s]);
          umma_commit(&tfull[as]);
        }
      } else {
        // both operands MN-major; M = 128 = two 64-slot atoms, the second one (LBO = one
        // stage further: valid shared memory, rows 64..127 of the accumulator) is ignored
        constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BN, 1, 1);
        constexpr uint32_t hiA = umma_desc_hi(8 * 128, 128);
        constexpr uint32_t hiB = umma_desc_hi(8 * rowB, rowB);
        for (int i = 0; i < n_my; ++i) {
          const int s = i % S;
          const uint32_t ph = (uint32_t)(i / S) & 1u;
          mbar_wait(&a_full[s], ph);
          mbar_wait(&z_full[s], ph);
          fence_proxy_async();
          tc_fence_after();
          const uint32_t a_addr = smem_u32(a_ring + s * kFcABytes);
          const uint32_t z_addr = smem_u32(z_ring + s * Cfg::kZBytes);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            umma_f16(tmem_base, umma_desc_pack(hiA, umma_desc_lo(a_addr + j * 2048, kFcABytes)),
                     umma_desc_pack(hiB, umma_desc_lo(z_addr + j * 16 * rowB, Cfg::kZBytes)), idesc,
                     (i | j) != 0 ? 1u : 0u);
          umma_commit(&a_empty[s]);
        }
        umma_commit(&tfull[0]);
      }
    }
  } else if (warp == 1) {
    // ======================= dZ tile producer (wgrad) =======================
    if (WGRAD && elect_one()) {
      for (int i = 0; i < n_my; ++i) {
        const int s = i % S;
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        mbar_wait(&a_empty[s], ((uint32_t)(i / S) & 1u) ^ 1u);
        mbar_expect_tx(&z_full[s], Cfg::kZBytes);
        tma_load_2d(&tmIO, &z_full[s], z_ring + s * Cfg::kZBytes, 0, tile * 128);
      }
    }
  } else if (warp < 10) {
    // ============================== epilogue ==============================
    const int quad = warp & 3;                       // TMEM lanes [32*quad, +32)
    const int half = (warp - 2) >> 2;                // which warp of the quadrant's pair
    if (!WGRAD) {
      // warp (quad, half) takes the tiles i with (i & 1) == half: accumulator stage `half`
      constexpr int NCH = BN / 32;
      uint8_t* my_stg = stg + (warp - 2) * 2 * Cfg::kStgBytes;
      const uint32_t tsrc = tmem_base + ((uint32_t)(quad * 32) << 16) + half * BN;
      int j = 0;
      for (int i = half; i < n_my; i += 2, ++j) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        uint8_t* box = my_stg + (j & 1) * Cfg::kStgBytes;
        if (lane == 0) bulk_wait_group_read<1>();    // the store that last read this box
        __syncwarp();
        mbar_wait(&tfull[half], (uint32_t)j & 1u);
        tc_fence_after();
        const uint32_t row = smem_u32(box) + (uint32_t)lane * rowB;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tsrc + c * 32, r);
          tmem_ld_wait();
          if (c == NCH - 1) {                        // the stage is drained: next tile may start
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[half]);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            if (P.flags & SEG_EPI_RELU) {
              o.x = pack_bf16x2_relu(__uint_as_float(r[q * 8 + 0]), __uint_as_float(r[q * 8 + 1]));
              o.y = pack_bf16x2_relu(__uint_as_float(r[q * 8 + 2]), __uint_as_float(r[q * 8 + 3]));
              o.z = pack_bf16x2_relu(__uint_as_float(r[q * 8 + 4]), __uint_as_float(r[q * 8 + 5]));
              o.w = pack_bf16x2_relu(__uint_as_float(r[q * 8 + 6]), __uint_as_float(r[q * 8 + 7]));
            } else {
              o.x = pack_bf16x2(__uint_as_float(r[q * 8 + 0]), __uint_as_float(r[q * 8 + 1]));
              o.y = pack_bf16x2(__uint_as_float(r[q * 8 + 2]), __uint_as_float(r[q * 8 + 3]));
              o.z = pack_bf16x2(__uint_as_float(r[q * 8 + 4]), __uint_as_float(r[q * 8 + 5]));
              o.w = pack_bf16x2(__uint_as_float(r[q * 8 + 6]), __uint_as_float(r[q * 8 + 7]));
            }
            const uint32_t chunk = (uint32_t)(c * 4 + q);
            const uint32_t pos = BN == 32 ? (chunk ^ ((uint32_t)(lane >> 1) & 3u))
                                          : (chunk ^ ((uint32_t)lane & 7u));
            sts128(row + (pos << 4), o);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmIO, box, 0, tile * 128 + quad * 32);
          bulk_commit_group();
        }
      }
      if (lane == 0) bulk_wait_group<0>();
    } else if (n_my > 0 && half == 0) {
      // partial sums of this CTA: accumulator rows 0..36 (k-slots + the constant-1 slot) ->
      // shared memory [37][BN] fp32 (the patch ring is idle once every MMA has completed);
      // the cluster-wide reduction below adds them to dW / db
      mbar_wait(&tfull[0], 0);
      tc_fence_after();
      const int L = quad * 32 + lane;                // accumulator row = k-slot
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + c * 32, r);
        tmem_ld_wait();
        if (L < kFcRows) {
          const uint32_t dst = smem_u32(a_ring) + (uint32_t)((L * BN + c * 32) * 4);
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            sts128(dst + j * 4, make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]));
        }
      }
    }
  } else {
    // ========================== patch-tile builders ==========================
    // four warps, thread = one output pixel of the tile: nine 8-byte cp.async copies (zero
    // filled outside the image) straight to the pixel's swizzled operand row, then one
    // deferred mbarrier arrival that fires when they have landed.  Nothing waits for data
    // here, so the builders run up to kFcStages tiles ahead of the tensor core.
    const int t = (warp - 10) * 32 + lane;
    const int HoWo = P.Ho * P.Wo;
    const uint32_t x7 = (uint32_t)t & 7u;
    for (int i = 0; i < n_my; ++i) {
      const int m = ((int)blockIdx.x + i * (int)gridDim.x) * 128 + t;
      const bool live = m < P.M_total;
      const uint32_t mm = live ? (uint32_t)m : 0u;
      const int n = (int)(__umulhi(mm, P.div_hw_mul) >> P.div_hw_shr);
      const uint32_t rem = mm - (uint32_t)(n * HoWo);
      const int oy = (int)(__umulhi(rem, P.div_wo_mul) >> P.div_wo_shr);
      const int ox = (int)rem - oy * P.Wo;
      const int iy0 = oy - P.pad_t, ix0 = ox - P.pad_l;
      const int s = i % S;
      mbar_wait(&a_empty[s], ((uint32_t)(i / S) & 1u) ^ 1u);
      const uint32_t row = smem_u32(a_ring + s * kFcABytes) + (uint32_t)t * 128u;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int iy = iy0 + r;
        const bool yok = live && (unsigned)iy < (unsigned)P.H;
        const int iyc = yok ? iy : 0;
#pragma unroll
        for (int sx = 0; sx < 3; ++sx) {
          const int ix = ix0 + sx;
          const bool ok = yok && (unsigned)ix < (unsigned)P.W;
          const uint2* src = P.x4 + ((int64_t)(n * P.H + iyc) * P.W + (ok ? ix : 0));
          const int q = r * 3 + sx;                  // 8-byte piece q of the row
          cp_async_8(row + ((((uint32_t)q >> 1) ^ x7) << 4) + (uint32_t)(q & 1) * 8u, src,
                     ok ? 8u : 0u);
        }
      }
      cp_async_arrive_noinc(&a_full[s]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (WGRAD) {
    // Every CTA adds its [37][BN] partial sum to dW / db: consecutive threads take
    // consecutive elements, so a warp's red.add covers whole 128-byte lines (one row of the
    // accumulator per thread - 28 lines per instruction - cost 20 us here; a DSMEM
    // pre-reduction over clusters of 8 CTAs was slower still: cluster scheduling, 47 vs 29 us)
    const uint32_t base = smem_u32(a_ring);
    for (uint32_t e = threadIdx.x; e < (uint32_t)(kFcRows * BN); e += kFcThreads) {
      const float v = lds32f(base + e * 4u);
      const int L = (int)(e / BN), co = (int)(e % BN);
      if (co < P.cout) {
        if (L < 36) {
          const int rr = L / 12, rem = L - rr * 12, ss = rem >> 2, ch = rem & 3;
          if (ch < 3) atomicAdd(P.dw + ((int64_t)((rr * 3 + ss) * 3 + ch)) * P.cout + co, v);
        } else if (P.db != nullptr) {                // the constant-1 slot: sum of dZ
          atomicAdd(P.db + co, v);
        }
      }
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace segb
