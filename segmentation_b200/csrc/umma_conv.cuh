// tcgen05 / TMEM / TMA implicit-GEMM kernels (device side) for sm_100a.
//
//   igemm_kernel : D[M=pixels, N] = A[pixels, K=(tap, channel)] * B[K, N]
//       A is gathered by TMA im2col loads straight from the NHWC activation tensor
//       (or a dense 2-D tiled load for 1x1), B is the bf16 weight shadow in its TF
//       layout, read K-major or MN-major depending on the op.  Used for conv fwd,
//       conv dgrad (flipped taps, full padding), k==s transposed conv fwd
//       (pixel-shuffle epilogue) and its dgrad.
//   wgrad_kernel : dW[(tap, ci), co] += A^T[(tap,ci), pixels] * B[pixels, co]
//       both operands MN-major (pixels are the GEMM K axis), split-K over pixels,
//       fp32 red.global.add into the master-layout gradient.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM
// allocator, warps 2..5 = epilogue (TMEM lane quadrant = warp_idx % 4).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace segb {

constexpr int kBlockM = 128;
constexpr int kIgemmThreads = 192;   // producer + MMA issuer + 4 epilogue warps (weight-gradient kernels)
// conv kernels with the register/LSU epilogue: 8 epilogue warps, two per TMEM lane quadrant,
// taking the 32-column chunks of a tile alternately.  One warp per scheduler issues only one
// instruction per ~5 cycles in this dependent code (3.1k cycles per chunk measured with 4
// warps), and for the small layers the epilogue of the last tile is not hidden by anything.
constexpr int kConvThreads = 320;

struct EpiDest {
  void* ptr;          // bf16 or fp32
  const bf16* mask;   // optional ReluGrad mask source (same geometry)
  int64_t sn, sh, sw; // strides of the destination view (elements)
  int64_t msn, msh, msw;
  int cols;           // valid columns in this destination
};

struct IgemmParams {
  // ---- A (activations)
  int a_tiled2d;             // 1: dense [M][C] 2-D tiled loads (1x1, stride 1, no pad)
  int M_total;               // GEMM rows = N*Ho*Wo
  int Ho, Wo;                // output-pixel grid (m -> n,p,q)
  int stride;                // base-pixel traversal stride
  int base_h, base_w;        // lower corner = -pad
  int kh, kw;
  int chunks1, chunks2;      // KC-chunks per tap taken from source 1 / source 2
  // ---- B (weights, 2-D [rows][cols] bf16)
  int tap_flip;              // use tap (taps-1-t) of B (conv dgrad)
  int b_rows_per_tap;
  // ---- N
  int N_total;               // multiple of BN
  // ---- epilogue
  EpiDest d0, d1;            // columns [0,split_n) -> d0, [split_n, N_total) -> d1
  int split_n;
  const float* bias;
  int flags;                 // SEG_EPI_*
  // pixel-shuffle (k==s transposed conv fwd): n -> (tap, co), m -> input pixel
  int ps_k;                  // 0 = linear rows; else kernel size (= stride)
  int ps_cout;               // channels per tap (padded)
};

template <int KC, int BN>
struct IgemmCfg {
  static constexpr int kSwzA = KC * 2;                       // bytes per A row
  static constexpr int kABytes = kBlockM * KC * 2;
  static constexpr int kBBytes = BN * KC * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // enough bytes in flight to cover the TMA latency even when a stage is only a few KB
  static constexpr int kStagesRaw = (200 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 24 ? 24 : kStagesRaw;
  static constexpr int kAtomN = BN < 64 ? BN : 64;           // MN-major B atom width
  static constexpr int kTmemCols = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128
                                   : 2 * BN <= 256 ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 512 /*barriers*/;
};

template <int KC, int BN, bool B_MN>
__global__ void __launch_bounds__(kConvThreads, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
             const __grid_constant__ CUtensorMap tmB, const IgemmParams P) {
  using Cfg = IgemmCfg<KC, BN>;
  constexpr int S = Cfg::kStages;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* tfull = bars + 2 * S;
  uint64_t* tempty = bars + 2 * S + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int taps = P.kh * P.kw;
  const int chunks = P.chunks1 + P.chunks2;
  const int num_kb = taps * chunks;
  const int m_tiles = (P.M_total + kBlockM - 1) / kBlockM;
  const int n_tiles = P.N_total / BN;
  const int total_tiles = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                 // everything above overlaps the previous kernel's tail

  if (warp == 0) {
    // =========================== TMA producer ===========================
    // warp-uniform control flow; one elected lane issues the copies
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * kBlockM;
        const int n0 = (tile % n_tiles) * BN;
        const int q0 = m0 % P.Wo;
        const int p0 = (m0 / P.Wo) % P.Ho;
        const int img0 = m0 / (P.Wo * P.Ho);
        const int cw = q0 * P.stride + P.base_w;
        const int ch = p0 * P.stride + P.base_h;
        int t = 0, j = 0, r = 0, s = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          if (elect_one()) {
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            uint8_t* sb = sa + Cfg::kABytes;
            mbar_expect_tx(&full[stage], Cfg::kStageBytes);
            const bool second = j >= P.chunks1;
            const CUtensorMap* tm = second ? &tmA2 : &tmA1;
            const int c0 = (second ? j - P.chunks1 : j) * KC;
            if (P.a_tiled2d)
              tma_load_2d(tm, &full[stage], sa, c0, m0);
            else
              tma_load_im2col_4d(tm, &full[stage], sa, c0, cw, ch, img0, (uint16_t)s, (uint16_t)r);
            const int bt = P.tap_flip ? taps - 1 - t : t;
            if (B_MN) {
              const int row = bt * P.b_rows_per_tap + j * KC;
#pragma unroll
              for (int a = 0; a < BN / Cfg::kAtomN; ++a)
                tma_load_2d(&tmB, &full[stage], sb + a * (KC * Cfg::kAtomN * 2),
                            n0 + a * Cfg::kAtomN, row);
            } else {
              tma_load_2d(&tmB, &full[stage], sb, j * KC, bt * P.b_rows_per_tap + n0);
            }
          }
          __syncwarp();
          if (++j == chunks) { j = 0; ++t; if (++s == P.kw) { s = 0; ++r; } }
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BN, 0, B_MN ? 1 : 0);
      constexpr uint32_t hiA = umma_desc_hi(8 * Cfg::kSwzA, Cfg::kSwzA);
      constexpr int atom_bytes = Cfg::kAtomN * 2;
      constexpr uint32_t hiB = B_MN ? umma_desc_hi(8 * atom_bytes, atom_bytes)
                                    : umma_desc_hi(8 * Cfg::kSwzA, Cfg::kSwzA);
      constexpr uint32_t lboB = B_MN ? KC * atom_bytes : 0;
      constexpr uint32_t kstepB = B_MN ? (16 * atom_bytes) >> 4 : 2;
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[as], aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        uint32_t acc = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t a_lo = umma_desc_lo(sa, 0);
          const uint32_t b_lo = umma_desc_lo(sa + Cfg::kABytes, lboB);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < KC / 16; ++kk) {
              umma_f16(tmem_d, umma_desc_pack(hiA, a_lo + kk * 2),
                       umma_desc_pack(hiB, b_lo + kk * kstepB), idesc, acc);
              acc = 1;
            }
            umma_commit(&empty[stage]);
          }
          __syncwarp();
          acc = 1;
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) umma_commit(&tfull[as]);
        __syncwarp();
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ============================= epilogue =============================
    const int quad = warp & 3;            // TMEM lanes [32*quad, 32*quad+32)
    const int half = (warp - 2) >> 2;     // which warp of the quadrant's pair
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * kBlockM;
      const int n0 = (tile % n_tiles) * BN;
      const int m = m0 + quad * 32 + lane;
      const bool row_ok = m < P.M_total;
      const bool second = P.d1.ptr != nullptr && n0 >= P.split_n;
      const EpiDest& D = second ? P.d1 : P.d0;
      const int nl0 = second ? n0 - P.split_n : n0;   // column inside destination
      // row -> destination offset
      int64_t off = 0, moff = 0;
      int ps_col_shift = 0;
      if (row_ok) {
        const int q = m % P.Wo;
        const int p = (m / P.Wo) % P.Ho;
        const int img = m / (P.Wo * P.Ho);
        if (P.ps_k) {
          const int tap = n0 / P.ps_cout;
          const int a = tap / P.ps_k, b = tap - a * P.ps_k;
          off = img * D.sn + (int64_t)(p * P.ps_k + a) * D.sh + (int64_t)(q * P.ps_k + b) * D.sw;
          ps_col_shift = tap * P.ps_cout;
        } else {
          off = img * D.sn + p * D.sh + q * D.sw;
          moff = img * D.msn + p * D.msh + q * D.msw;
        }
      }
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int cc = half * (BN >= 32 ? 32 : 16); cc < BN; cc += 2 * (BN >= 32 ? 32 : 16)) {
        constexpr int W = BN >= 32 ? 32 : 16;
        uint32_t r[W];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN + cc;
        if (W == 32) tmem_ld_32x32(taddr, r); else tmem_ld_32x16(taddr, r);
        tmem_ld_wait();
        if (row_ok) {
          const int ncol = nl0 + cc - ps_col_shift;     // first column of this chunk in D
          float v[W];
#pragma unroll
          for (int j = 0; j < W; ++j) {
            v[j] = __uint_as_float(r[j]);
            // bias is indexed by the destination column (logical channel); padded
            // columns are never stored, so never read a bias for them either
            if ((P.flags & SEG_EPI_BIAS) && ncol + j < D.cols) v[j] += __ldg(P.bias + ncol + j);
            if (P.flags & SEG_EPI_RELU) v[j] = fmaxf(v[j], 0.f);
          }
          if ((P.flags & SEG_EPI_RELU_MASK) && D.mask) {
            const bf16* mp = D.mask + moff + ncol;
            if (ncol + W <= D.cols) {
#pragma unroll
              for (int j = 0; j < W; j += 8) {
                const uint4 u = *reinterpret_cast<const uint4*>(mp + j);
                const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (!(bf16_lo(w4[e]) > 0.f)) v[j + 2 * e] = 0.f;
                  if (!(bf16_hi(w4[e]) > 0.f)) v[j + 2 * e + 1] = 0.f;
                }
              }
            } else {
              for (int j = 0; j < W && ncol + j < D.cols; ++j)
                if (!(__bfloat162float(mp[j]) > 0.f)) v[j] = 0.f;
            }
          }
          if (P.flags & SEG_EPI_OUT_F32) {
            float* op = reinterpret_cast<float*>(D.ptr) + off + ncol;
            for (int j = 0; j < W && ncol + j < D.cols; ++j) op[j] = v[j];
          } else {
            bf16* op = reinterpret_cast<bf16*>(D.ptr) + off + ncol;
            if (ncol + W <= D.cols) {
#pragma unroll
              for (int j = 0; j < W; j += 8) {
                uint4 o;
                o.x = pack_bf16x2(v[j], v[j + 1]);
                o.y = pack_bf16x2(v[j + 2], v[j + 3]);
                o.z = pack_bf16x2(v[j + 4], v[j + 5]);
                o.w = pack_bf16x2(v[j + 6], v[j + 7]);
                *reinterpret_cast<uint4*>(op + j) = o;
              }
            } else {
              for (int j = 0; j < W && ncol + j < D.cols; ++j) op[j] = __float2bfloat16(v[j]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------
// wgrad: dW[(tap, bc), sc] += sum_pixels big[pixel@tap, bc] * small[pixel, sc]
// ---------------------------------------------------------------------------
struct WgradUmmaParams {
  int M_total;            // pixels of `small` (GEMM K)
  int Ho, Wo, stride, base_h, base_w, kh, kw;
  int chunks1, chunks2;   // AW-channel chunks per tap from big / big2
  int total_atoms;        // taps * (chunks1 + chunks2)
  int n_tiles;            // small channels / BN
  int kb_per_split;       // k-blocks (PK pixels each) per CTA
  int BC, SC;             // logical channel counts of dW [taps][BC][SC]
  float* dw;
  // Bias gradient on the tensor core: the A-atom slot with index == total_atoms is
  // filled with ones (never loaded), so its first accumulator row is
  // sum_pixels small[pixel, :] = BiasAddGrad when `small` is dz.  Null: disabled.
  float* db;
};

constexpr int kWgradPK = 64;   // pixels per pipeline stage

template <int AW, int BN>
struct WgradCfg {
  static constexpr int kNA = kBlockM / AW;                    // A atoms per MMA
  static constexpr int kAtomBytesA = kWgradPK * AW * 2;
  static constexpr int kABytes = kNA * kAtomBytesA;           // = 128 * PK * 2
  static constexpr int kAtomN = BN < 64 ? BN : 64;
  static constexpr int kNB = BN / kAtomN;
  static constexpr int kAtomBytesB = kWgradPK * kAtomN * 2;
  static constexpr int kBBytes = kNB * kAtomBytesB;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagesRaw = (200 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 16 ? 16 : kStagesRaw;
  static constexpr int kTmemCols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 512;
};

template <int AW, int BN>
__global__ void __launch_bounds__(kIgemmThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
             const __grid_constant__ CUtensorMap tmB, const WgradUmmaParams P) {
  using Cfg = WgradCfg<AW, BN>;
  constexpr int S = Cfg::kStages;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* tfull = bars + 2 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int chunks = P.chunks1 + P.chunks2;
  const int group = blockIdx.x / P.n_tiles;           // M-group of kNA atoms
  const int n0 = (blockIdx.x % P.n_tiles) * BN;
  const int total_kb = (P.M_total + kWgradPK - 1) / kWgradPK;
  const int kb_begin = blockIdx.y * P.kb_per_split;
  const int kb_end = min(total_kb, kb_begin + P.kb_per_split);
  const int num_kb = kb_end - kb_begin;                // may be <= 0 for the last split

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&tfull[0], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  // A-atom slots that hold no real (tap, channel-chunk) atom are filled with 1.0 once,
  // in every stage (all-ones is invariant under the smem swizzle).
  int real_atoms = P.total_atoms - group * Cfg::kNA;
  real_atoms = real_atoms < 0 ? 0 : (real_atoms > Cfg::kNA ? Cfg::kNA : real_atoms);
  if (real_atoms < Cfg::kNA) {
    const int words_per_atom = Cfg::kAtomBytesA / 4;
    for (int st = 0; st < S; ++st)
      for (int a = real_atoms; a < Cfg::kNA; ++a) {
        uint32_t* dst =
            reinterpret_cast<uint32_t*>(smem + st * Cfg::kStageBytes + a * Cfg::kAtomBytesA);
        for (int i = threadIdx.x; i < words_per_atom; i += blockDim.x) dst[i] = 0x3F803F80u;
      }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                 // everything above overlaps the previous kernel's tail

  if (num_kb > 0) {
    if (warp == 0) {
      // TMA producer: warp-uniform control flow, elected lane issues
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = real_atoms * Cfg::kAtomBytesA + Cfg::kBBytes;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        const int m0 = kb * kWgradPK;
        const int q0 = m0 % P.Wo;
        const int p0 = (m0 / P.Wo) % P.Ho;
        const int img0 = m0 / (P.Wo * P.Ho);
        const int cw = q0 * P.stride + P.base_w;
        const int ch = p0 * P.stride + P.base_h;
        mbar_wait(&empty[stage], phase ^ 1u);
        if (elect_one()) {
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_expect_tx(&full[stage], tx_bytes);
#pragma unroll
          for (int a = 0; a < Cfg::kNA; ++a) {
            const int atom = group * Cfg::kNA + a;
            if (atom >= P.total_atoms) break;          // ones-filled slots: no load
            const int t = atom / chunks;
            const int j = atom - t * chunks;
            const int r = t / P.kw, s = t - r * P.kw;
            const bool second = j >= P.chunks1;
            const CUtensorMap* tm = second ? &tmA2 : &tmA1;
            const int c0 = (second ? j - P.chunks1 : j) * AW;
            tma_load_im2col_4d(tm, &full[stage], sa + a * Cfg::kAtomBytesA, c0, cw, ch, img0,
                               (uint16_t)s, (uint16_t)r);
          }
#pragma unroll
          for (int b = 0; b < Cfg::kNB; ++b)
            tma_load_2d(&tmB, &full[stage], sb + b * Cfg::kAtomBytesB, n0 + b * Cfg::kAtomN, m0);
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BN, 1, 1);
      constexpr int rowA = AW * 2, rowB = Cfg::kAtomN * 2;
      constexpr uint32_t hiA = umma_desc_hi(8 * rowA, rowA);
      constexpr uint32_t hiB = umma_desc_hi(8 * rowB, rowB);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t acc = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
        const uint32_t a_lo = umma_desc_lo(sa, Cfg::kAtomBytesA);
        const uint32_t b_lo = umma_desc_lo(sa + Cfg::kABytes, Cfg::kAtomBytesB);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kWgradPK / 16; ++kk) {
            umma_f16(tmem_base, umma_desc_pack(hiA, a_lo + kk * ((16 * rowA) >> 4)),
                     umma_desc_pack(hiB, b_lo + kk * ((16 * rowB) >> 4)), idesc, acc);
            acc = 1;
          }
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        acc = 1;
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
      if (elect_one()) umma_commit(&tfull[0]);
      __syncwarp();
    } else {
      const int quad = warp & 3;
      const int L = quad * 32 + lane;                  // accumulator row = (atom, channel)
      const int atom = group * Cfg::kNA + L / AW;
      const int chn = L % AW;
      const bool atom_ok = atom < P.total_atoms;
      const int t = atom_ok ? atom / chunks : 0;
      const int bc = atom_ok ? (atom - t * chunks) * AW + chn : 0;
      bool row_ok = atom_ok && bc < P.BC;
      float* dst = P.dw + ((int64_t)t * P.BC + bc) * P.SC;
      if (P.db != nullptr && atom == P.total_atoms && chn == 0) {   // ones-atom: bias grad
        row_ok = true;
        dst = P.db;
      }
      mbar_wait(&tfull[0], 0);
      tc_fence_after();
      // Whole BN-float row segments go to L2 as one TMA reduce-add each (row-per-thread
      // red.v4 is 2.7x slower, tools/probe_red.py): the row is staged in the pipeline
      // shared memory, which is idle once every MMA has completed.
      constexpr int kPitch = BN + 4;
      const bool vec_ok = (P.SC & 3) == 0 && (reinterpret_cast<uintptr_t>(P.dw) & 15) == 0;
      const bool bulk_ok = vec_ok && BN >= 32 && n0 + BN <= P.SC &&
                           128 * kPitch * 4 <= S * Cfg::kStageBytes;
      const bool bias_row = P.db != nullptr && dst == P.db;
      const uint32_t s_row = smem_u32(smem) + (uint32_t)(L * kPitch * 4);
#pragma unroll 1
      for (int cc = 0; cc < BN; cc += (BN >= 32 ? 32 : 16)) {
        constexpr int W = BN >= 32 ? 32 : 16;
        uint32_t r[W];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + cc;
        if (W == 32) tmem_ld_32x32(taddr, r); else tmem_ld_32x16(taddr, r);
        tmem_ld_wait();
        if (bulk_ok && !bias_row) {
#pragma unroll
          for (int j = 0; j < W; j += 4)
            sts128(s_row + (uint32_t)(cc + j) * 4, make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]));
        } else if (row_ok) {
          if (vec_ok && !bias_row) {
            // 16-byte vector reductions (dW rows are SC floats, SC % 4 == 0 keeps alignment)
#pragma unroll
            for (int j = 0; j < W; j += 4) {
              const int sc = n0 + cc + j;
              if (sc < P.SC)
                red_add_v4(dst + sc, __uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                           __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            }
          } else {
#pragma unroll
            for (int j = 0; j < W; ++j) {
              const int sc = n0 + cc + j;
              if (sc < P.SC) atomicAdd(dst + sc, __uint_as_float(r[j]));
            }
          }
        }
      }
      if (bulk_ok && !bias_row) {
        fence_proxy_async();
        if (row_ok) bulk_reduce_add_f32(dst + n0, s_row, BN * 4);
        bulk_commit_group();
        bulk_wait_group<0>();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace segb
