// Spatial-tile tcgen05 weight gradient for 3x3 stride-1 convolutions.
//
//   dW[tap][ci][co] += sum over pixels  X[pixel + tap][ci] * dZ[pixel][co]
//
// One CTA owns one (AW-channel chunk of ci) x (BN-channel slice of co) block of dW for
// ALL nine taps and a subset of the pixel tiles; its accumulators stay in TMEM for the
// whole kernel and are added to dW once at the end (16-byte vector reductions).
//
// A pixel tile is 8 rows x 16 columns.  Per tile the producer issues ONE TMA load of the
// X halo box [10][18][AW] and one of the dZ box [8][16][BN]; out-of-bounds pixels are
// zero-filled, so padding and ragged edges contribute nothing.  Pixels are the GEMM K
// axis, so both operands are MN-major: a k-step is one tile row (16 pixels = two 8-row
// groups, SBO = 8 pixel rows).  The M = 128 rows of one MMA are 128/AW "atoms" of AW
// channels; consecutive atoms are the SAME staged box shifted by a whole number of
// pixels (the descriptor's leading-dimension byte offset), i.e. different filter taps:
//   AW = 64: taps (2m, 2m+1) per MMA m = 0..4  (LBO = 1 pixel, or PW-2 across a row end)
//   AW = 32 / 16: taps (r,0..2) + unused atoms per MMA r = 0..2  (LBO = 1 pixel)
// so X is fetched from L2 once per tile instead of once per tap.
//
// BiasAddGrad: the four otherwise idle epilogue warps sum the staged dZ tiles over
// pixels (CTAs of the first ci chunk only).
//
// Epilogue.  Every CTA holds a partial sum over its pixel tiles, to be added into dW.
// red.global.v4 with one accumulator row per thread (32 scattered 16-byte requests per
// warp instruction) ran at 36-40k cycles for 148 CTAs x 36864 floats - twice the main
// loop; whole 128-byte lines are 2.7x faster (tools/probe_red.py).  So each thread
// stages its row in the (by then idle) pipeline shared memory and hands BN-float row
// segments to TMA reduce-add (cp.reduce.async.bulk .add.f32).
#pragma once
#include "umma_conv.cuh"

namespace segb {

struct TwgradParams {
  int tiles_x, tiles_y, batch;
  int chunks1, chunks2;        // AW-channel chunks from source 1 / source 2
  int n_slices;                // padded Cout / BN
  int ctas_per_combo;
  int pad_t, pad_l;
  int BC, SC;                  // logical dims of dW [9][BC][SC]
  float* dw;
  float* db;                   // nullable
  int stages, stage_bytes, x_bytes, off_bars;
  int staged_ok;               // the pipeline stages can hold the staged accumulators
};

constexpr int kTwTH = 8, kTwTW = 16, kTwPW = kTwTW + 2, kTwPH = kTwTH + 2;
constexpr int kTwMaxStages = 8;

// TRED (seg_set_option key 15, default on; 1.056 -> 1.042 ms / U-Net step): the partial sums
// leave the CTA as TMA TENSOR reduce-adds - one [AW ci] x [32 co] fp32 box per (tap, 32-column block), 4 per MMA
// and 18-20 per CTA, issued by one thread per group of four epilogue warps - instead of one
// 256-byte bulk reduce-add per accumulator row and MMA (640 per CTA, whose per-lane issue
// was measured at ~9 k of the epilogue's ~25 k cycles).  tmDW: dW as a 3-D tensor
// {SC, BC, 9} with 128-byte-swizzled boxes {32, AW, 1} (clipped at BC and SC).
template <int AW, int BN, bool TRED = false>
__global__ void __launch_bounds__(kConvThreads, 1)
twgrad_kernel(const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmX2,
              const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmDW,
              const TwgradParams P) {
  constexpr int rowA = AW * 2;                          // bytes per staged X pixel
  constexpr int kAtomN = BN < 64 ? BN : 64;
  constexpr int rowB = kAtomN * 2;                      // bytes per staged dZ pixel (per atom)
  constexpr int kNB = BN / kAtomN;
  constexpr int kZAtomBytes = kTwTH * kTwTW * rowB;     // one dZ atom: 128 pixels
  constexpr uint32_t kXBoxBytes = kTwPW * kTwPH * rowA;
  constexpr uint32_t kTxBytes = kXBoxBytes + kNB * kZAtomBytes;
  constexpr int kNMma = AW == 64 ? 5 : 3;
  constexpr int kAtoms = kBlockM / AW;                  // atoms per MMA
  constexpr int kCols = kNMma * BN;
  constexpr int kTmemCols = kCols <= 32 ? 32 : kCols <= 64 ? 64 : kCols <= 128 ? 128
                            : kCols <= 256 ? 256 : 512;
  static_assert(kCols <= 512, "accumulators exceed TMEM");

  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P.off_bars);
  uint64_t* full = bars;
  uint64_t* empty = bars + kTwMaxStages;
  uint64_t* tfull = bars + 2 * kTwMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kTwMaxStages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int combo = blockIdx.x / P.ctas_per_combo;
  const int sub = blockIdx.x - combo * P.ctas_per_combo;
  const int chunk = combo / P.n_slices;
  const int n0 = (combo - chunk * P.n_slices) * BN;
  const int total_tiles = P.batch * P.tiles_y * P.tiles_x;
  const int my_tiles = sub < total_tiles ? (total_tiles - sub + P.ctas_per_combo - 1) / P.ctas_per_combo : 0;
  const bool do_db = P.db != nullptr && chunk == 0;
  constexpr int kPitch = kCols + 4;                    // floats per staged accumulator row
  const bool vec_ok = (P.SC & 3) == 0 && (reinterpret_cast<uintptr_t>(P.dw) & 15) == 0;
  const bool bulk_ok = vec_ok && P.staged_ok && n0 + BN <= P.SC;   // whole row segments

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX1);
    tma_prefetch_desc(&tmX2);
    tma_prefetch_desc(&tmZ);
    for (int i = 0; i < P.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], do_db ? 5 : 1);
    }
    mbar_init(&tfull[0], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                 // everything above overlaps the previous kernel's tail

  if (my_tiles > 0) {
    if (warp == 0) {
      // =========================== TMA producer ===========================
      if (elect_one()) {
        const bool second = chunk >= P.chunks1;
        const CUtensorMap* tmX = second ? &tmX2 : &tmX1;
        const int c0 = (second ? chunk - P.chunks1 : chunk) * AW;
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = sub; tile < total_tiles; tile += P.ctas_per_combo) {
          const int tx = tile % P.tiles_x;
          int t2 = tile / P.tiles_x;
          const int ty = t2 % P.tiles_y;
          const int img = t2 / P.tiles_y;
          const int x0 = tx * kTwTW, y0 = ty * kTwTH;
          mbar_wait(&empty[stage], phase ^ 1u);
          uint8_t* sx = smem + stage * P.stage_bytes;
          uint8_t* sz = sx + P.x_bytes;
          mbar_expect_tx(&full[stage], kTxBytes);
          tma_load_4d(tmX, &full[stage], sx, c0, x0 - P.pad_l, y0 - P.pad_t, img);
#pragma unroll
          for (int b = 0; b < kNB; ++b)
            tma_load_4d(&tmZ, &full[stage], sz + b * kZAtomBytes, n0 + b * kAtomN, x0, y0, img);
          if (++stage == P.stages) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (warp == 1) {
      // ============================ MMA issuer ============================
      if (elect_one()) {
        constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BN, 1, 1);
        constexpr uint32_t hiA = umma_desc_hi(8 * rowA, rowA);
        constexpr uint32_t hiB = umma_desc_hi(8 * rowB, rowB);
        int stage = 0;
        uint32_t phase = 0;
        uint32_t acc = 0;
        for (int it = 0; it < my_tiles; ++it) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sx = smem_u32(smem + stage * P.stage_bytes);
          const uint32_t sz = sx + P.x_bytes;
          const uint32_t b_lo = umma_desc_lo(sz, kZAtomBytes);
#pragma unroll
          for (int y = 0; y < kTwTH; ++y) {
#pragma unroll
            for (int m = 0; m < kNMma; ++m) {
              // first tap of this MMA and the pixel distance to the next atom's tap
              constexpr int kDummy = 0;
              (void)kDummy;
              const int t0 = AW == 64 ? 2 * m : 3 * m;
              const int r0 = t0 / 3, s0 = t0 % 3;
              const int dpx = (AW == 64 && m == 1) ? kTwPW - 2 : 1;
              const uint32_t a_addr = sx + (uint32_t)(((y + r0) * kTwPW + s0) * rowA);
              const uint32_t a_lo = umma_desc_lo(a_addr, (uint32_t)(dpx * rowA));
              umma_f16(tmem_base + m * BN, umma_desc_pack(hiA, a_lo),
                       umma_desc_pack(hiB, b_lo + (uint32_t)((y * kTwTW * rowB) >> 4)), idesc,
                       y == 0 ? acc : 1u);
            }
          }
          acc = 1;
          umma_commit(&empty[stage]);
          if (++stage == P.stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull[0]);
      }
    } else {
      // ====================== bias gradient + epilogue ======================
      // 8 warps, two per TMEM lane quadrant; the first four also sum the bias gradient
      // during the main loop, all eight share the accumulator drain (alternate MMAs)
      const int quad = warp & 3;
      const int half = (warp - 2) >> 2;
      const int et = quad * 32 + lane;                 // 0..127 within each group of four warps
      if (do_db && half == 0) {
        // thread -> one 16-byte chunk (8 channels) of the dZ rows r with r % kRG == rg
        constexpr int kChunks = BN / 8;                // 16-byte chunks per pixel (all atoms)
        constexpr int kRG = 128 / kChunks;             // row groups
        const int ch = et % kChunks, rg = et / kChunks;
        const int atom = ch / (kAtomN / 8), q = ch % (kAtomN / 8);
        float s[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] = 0.f;
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < my_tiles; ++it) {
          mbar_wait(&full[stage], phase);
          const uint32_t sz = smem_u32(smem + stage * P.stage_bytes + P.x_bytes + atom * kZAtomBytes);
#pragma unroll 4
          for (int p = rg; p < 128; p += kRG) {
            const uint32_t sw = rowB == 128 ? (uint32_t)(p & 7) : rowB == 64 ? (uint32_t)((p >> 1) & 3)
                                                                             : (uint32_t)((p >> 2) & 1);
            const uint4 u = lds128(sz + p * rowB + (((uint32_t)q ^ sw) << 4));
            s[0] += bf16_lo(u.x); s[1] += bf16_hi(u.x);
            s[2] += bf16_lo(u.y); s[3] += bf16_hi(u.y);
            s[4] += bf16_lo(u.z); s[5] += bf16_hi(u.z);
            s[6] += bf16_lo(u.w); s[7] += bf16_hi(u.w);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[stage]);
          if (++stage == P.stages) { stage = 0; phase ^= 1u; }
        }
        // lanes of a warp sharing the same chunk: lane stride kChunks
#pragma unroll
        for (int o = 16; o >= kChunks && o > 0; o >>= 1)
#pragma unroll
          for (int e = 0; e < 8; ++e) s[e] += __shfl_xor_sync(0xffffffffu, s[e], o);
        if (lane < kChunks || kChunks > 32) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int co = n0 + ch * 8 + e;
            if (co < P.SC) atomicAdd(P.db + co, s[e]);
          }
        }
      }
      // accumulator row L = (atom, channel): which tap / ci does it hold?
      const int L = et;
      const int a = L / AW, cl = L % AW;
      const int ci = chunk * AW + cl;
      mbar_wait(&tfull[0], 0);
      tc_fence_after();
      // every MMA has completed, so all stages were consumed; the other epilogue warps may
      // still be summing the last dZ tiles out of them
      if (bulk_ok) named_bar_sync(2, 256);
      if (TRED) {
        // box layout: box (m, atom, 32-column block) = AW rows of 128 bytes, 16-byte chunk j
        // of row cl stored at chunk j ^ (cl & 7) (the tensor map's 128-byte swizzle)
        constexpr int kCB = BN / 32;
        constexpr uint32_t kBoxBytes = AW * 128;
#pragma unroll 1
        for (int m = half; m < kNMma; m += 2) {
#pragma unroll 1
          for (int cc = 0; cc < BN; cc += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + m * BN + cc, r);
            tmem_ld_wait();
            const uint32_t row = smem_u32(smem) +
                                 (uint32_t)((m * kAtoms + a) * kCB + (cc >> 5)) * kBoxBytes +
                                 (uint32_t)cl * 128u;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts128(row + (uint32_t)((j ^ (cl & 7)) << 4),
                     make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]));
          }
          fence_proxy_async();
          named_bar_sync(3 + half, 128);           // the four warps that staged MMA m
          if (et == 0) {
#pragma unroll 1
            for (int a2 = 0; a2 < kAtoms; ++a2) {
              const int tap2 = AW == 64 ? 2 * m + a2 : (a2 < 3 ? 3 * m + a2 : 9);
              if (tap2 >= 9) continue;
#pragma unroll 1
              for (int cb = 0; cb < kCB; ++cb)
                if (n0 + cb * 32 < P.SC)
                  tma_reduce_add_3d(&tmDW,
                                    smem_u32(smem) + (uint32_t)((m * kAtoms + a2) * kCB + cb) * kBoxBytes,
                                    n0 + cb * 32, chunk * AW, tap2);
            }
          }
        }
        if (et == 0) {
          bulk_commit_group();
          bulk_wait_group<0>();                    // complete before the CTA exits
        }
      } else {
#pragma unroll 1
      for (int m = half; m < kNMma; m += 2) {
        const int tap = AW == 64 ? 2 * m + a : (a < 3 ? 3 * m + a : 9);
        const bool row_ok = tap < 9 && ci < P.BC;
        float* dst = P.dw + ((int64_t)tap * P.BC + ci) * P.SC;
#pragma unroll 1
        for (int cc = 0; cc < BN; cc += 32) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + m * BN + cc, r);
          tmem_ld_wait();
          if (bulk_ok) {
            // stage row L, columns [m*BN+cc, +32) of the partial sum: S[L][kPitch]
            const uint32_t sa = smem_u32(smem) + (uint32_t)((L * kPitch + m * BN + cc) * 4);
#pragma unroll
            for (int j = 0; j < 32 && j < BN; j += 4)
              sts128(sa + j * 4, make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]));
          } else if (row_ok) {
            if (vec_ok) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const int sc = n0 + cc + j;
                if (sc < P.SC)
                  red_add_v4(dst + sc, __uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                             __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int sc = n0 + cc + j;
                if (sc < P.SC) atomicAdd(dst + sc, __uint_as_float(r[j]));
              }
            }
          }
        }
        if (bulk_ok) {
          // one TMA reduce-add of this thread's BN-float row segment: L2 receives whole
          // 128-byte lines (row-per-thread red.v4 was measured 2.7x slower, probe_red.py)
          fence_proxy_async();
          if (row_ok)
            bulk_reduce_add_f32(dst + n0, smem_u32(smem) + (uint32_t)((L * kPitch + m * BN) * 4),
                                BN * 4);
        }
      }
      if (bulk_ok) {
        bulk_commit_group();
        bulk_wait_group<0>();                          // complete before the CTA exits
      }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace segb
