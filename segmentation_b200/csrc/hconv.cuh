// Halo-tile tcgen05 convolution (stride 1): ONE staged input tile per channel chunk
// serves every filter tap.
//
// Position space: the (virtually padded) input is a grid of Hp x Wp positions per
// image; position m = (img*Hp + yp)*Wp + xp.  The output pixel (img, yp, xp) — valid
// iff yp < Ho, xp < Wo — is  sum_{r,s,c} in[m + r*Wp + s, c] * w[r,s,c,:]:  for a tile of
// 128 consecutive positions every tap is the SAME staged rows shifted by r*Wp+s rows,
// i.e. the same K-major UMMA descriptor with its start address advanced by
// (r*Wp+s)*row_bytes (the smem swizzle is a function of the absolute address, see
// tools/probe_shift.py).  A is therefore fetched from L2 once per (tile, chunk)
// instead of once per tap; junk positions (xp >= Wo or yp >= Ho) are computed and
// dropped (efficiency Ho*Wo/(Hp*Wp)).
//
// Staging modes:  flat — dense unpadded input: 2-D TMA boxes over [N*H*W][C];
//                 rows — any NHWC view / zero padding: one 4-D TMA box per padded
//                        input row (out-of-bounds coordinates are zero-filled).
// B (weights) streams through its own smem ring, one (chunk, tap) tile per stage, or
// stays resident for the whole kernel when all tiles of the CTA's N-slice fit.
#pragma once
#include "umma_conv.cuh"

namespace segb {

struct HconvParams {
  int Wp;                    // position-space row pitch (>= row_px; smem rows per padded row)
  int row_px;                // pixels per padded input row actually loaded (rows mode box width)
  int Hp, batch, Ho, Wo;
  int P_total;               // batch*Hp*Wp
  int kh, kw;
  int flat, box_rows, nboxes;
  int pad_t, pad_l;
  int a_stage_bytes;
  int chunks1, chunks2;
  int tap_flip, b_rows_per_tap;
  int use_tap_rows;          // weight-matrix row of tap t (loop order r*kw+s) comes from tap_rows[t]
  int tap_rows[25];          //   instead of (t or taps-1-t) * b_rows_per_tap: sub-kernels of a
                             //   strided transposed conv pick every stride-th tap of the k x k bank
  int N_total;
  int SA, SB;
  int b_resident;
  EpiDest d0, d1;
  int split_n;
  const float* bias;
  int flags;
  // optional per-column affine map applied after bias / ReLU (an inference batch-norm folded
  // to scale and shift, seg_batchnorm_fold): N_total floats each, or null
  const float* post_scale;
  const float* post_shift;
  long long* prof;           // optional in-kernel timeline of CTA 0 (test hook), else null
};

constexpr int kHconvMaxSA = 8;
constexpr int kProfTiles = 16;     // tiles recorded per role; 4 events each
// The timeline hook is compiled in only with -DSEGB200_KERNEL_PROF=1 (tools/layer_prof.py):
// its checks sit inside the single-thread MMA issue loop, which is on the critical path of
// the small deep layers.
#ifndef SEGB200_KERNEL_PROF
#define SEGB200_KERNEL_PROF 0
#endif
__device__ __forceinline__ void prof_mark(long long* prof, int role, int tile_i, int ev) {
#if SEGB200_KERNEL_PROF
  if (prof != nullptr && blockIdx.x == 0 && tile_i < kProfTiles)
    prof[(role * kProfTiles + tile_i) * 4 + ev] = clock64();
#endif
}
constexpr int kHconvMaxSB = 40;

// TPS: filter taps per weight stage (1, or 3 = one filter row of a k x 3 kernel per stage:
// a third of the full/empty barrier waits, commits and elections in the single-thread issue
// loop, whose ~80 instructions per tap - not the tensor pipe - set the pace of the small
// deep layers, profiles/r01_ncu_kernels.md §5).  Streamed weights only.
template <int KC, int BN, bool B_MN, int TPS = 1>
__global__ void __launch_bounds__(kConvThreads, 1)
hconv_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
             const __grid_constant__ CUtensorMap tmB, const HconvParams P) {
  constexpr int SWZ = KC * 2;
  constexpr int kBBytes = BN * KC * 2;
  constexpr int kStageB = TPS * kBBytes;          // one weight stage
  static_assert(TPS == 1 || TPS == 3, "weight stages hold one tap or one filter row");
  constexpr int kAtomN = BN < 64 ? BN : 64;
  constexpr int kTmemCols = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128
                            : 2 * BN <= 256 ? 256 : 512;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_b = smem + P.SA * P.a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + P.SB * kStageB);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + kHconvMaxSA;
  uint64_t* b_full = bars + 2 * kHconvMaxSA;
  uint64_t* b_empty = b_full + kHconvMaxSB;
  uint64_t* tfull = b_empty + kHconvMaxSB;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int taps = P.kh * P.kw;
  const int chunks = P.chunks1 + P.chunks2;
  const int m_tiles = (P.P_total + kBlockM - 1) / kBlockM;
  const int n_tiles = P.N_total / BN;
  const int halo = (P.kh - 1) * P.Wp + P.kw - 1;
  const int total_tiles = m_tiles * n_tiles;
  const int first_unit = (int)blockIdx.x;
  const int unit_step = (int)gridDim.x;
  auto tile_m0 = [&](int unit) { return (unit / n_tiles) * kBlockM; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < P.SA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < P.SB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                 // everything above overlaps the previous kernel's tail

  if (warp == 0) {
    // =========================== TMA producer ===========================
    // whole warp runs the (uniform) control flow; one elected lane issues the copies
    {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      bool first_tile = true;
      int ti = 0;
      for (int tile = first_unit; tile < total_tiles; tile += unit_step, ++ti) {
        const int m0 = tile_m0(tile);
        const int n0 = (tile % n_tiles) * BN;
        if (lane == 0) prof_mark(P.prof, 0, ti, 0);
        const int g0 = m0 / P.Wp;                          // first padded row (global)
        const int g1 = (m0 + kBlockM - 1 + halo) / P.Wp;   // last padded row needed
        for (int j = 0; j < chunks; ++j) {
          const bool second = j >= P.chunks1;
          const CUtensorMap* tm = second ? &tmA2 : &tmA1;
          const int c0 = (second ? j - P.chunks1 : j) * KC;
          mbar_wait(&a_empty[sa], pa ^ 1u);
          if (j == 0 && lane == 0) prof_mark(P.prof, 0, ti, 1);
          uint8_t* dst = smem + sa * P.a_stage_bytes;
          if (elect_one()) {
            if (P.flat) {
              mbar_expect_tx(&a_full[sa], (uint32_t)(P.nboxes * P.box_rows) * SWZ);
              for (int b = 0; b < P.nboxes; ++b)
                tma_load_2d(tm, &a_full[sa], dst + b * P.box_rows * SWZ, c0, m0 + b * P.box_rows);
            } else {
              mbar_expect_tx(&a_full[sa], (uint32_t)((g1 - g0 + 1) * P.row_px) * SWZ);
              int img = g0 / P.Hp;
              int yp = g0 - img * P.Hp;
              for (int g = g0; g <= g1; ++g) {
                tma_load_4d(tm, &a_full[sa], dst + (g - g0) * P.Wp * SWZ, c0, -P.pad_l,
                            yp - P.pad_t, img);
                if (++yp == P.Hp) { yp = 0; ++img; }
              }
            }
          }
          __syncwarp();
          if (++sa == P.SA) { sa = 0; pa ^= 1u; }
          if (!P.b_resident || first_tile) {
            int bt = P.tap_flip ? taps - 1 : 0;
            for (int t = 0; t < taps; t += TPS) {
              mbar_wait(&b_empty[sb], pb ^ 1u);
              if (elect_one()) {
                mbar_expect_tx(&b_full[sb], kStageB);
#pragma unroll
                for (int u = 0; u < TPS; ++u) {
                uint8_t* sbp = smem_b + sb * kStageB + u * kBBytes;
                const int bt_u = bt + (P.tap_flip ? -u : u);
                const int tap_row = P.use_tap_rows ? P.tap_rows[t + u] : bt_u * P.b_rows_per_tap;
                if (B_MN) {
                  const int row = tap_row + j * KC;
#pragma unroll
                  for (int a = 0; a < BN / kAtomN; ++a)
                    tma_load_2d(&tmB, &b_full[sb], sbp + a * (KC * kAtomN * 2), n0 + a * kAtomN,
                                row);
                } else {
                  tma_load_2d(&tmB, &b_full[sb], sbp, j * KC, tap_row + n0);
                }
                }
              }
              __syncwarp();
              bt += P.tap_flip ? -TPS : TPS;
              if (++sb == P.SB) { sb = 0; pb ^= 1u; }
            }
          }
        }
        first_tile = false;
        if (lane == 0) prof_mark(P.prof, 0, ti, 2);
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    // warp-uniform control flow; descriptors advance by 32-bit adds on the low word;
    // one elected lane issues tcgen05.mma / tcgen05.commit
    {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BN, 0, B_MN ? 1 : 0);
      constexpr uint32_t hiA = umma_desc_hi(8 * SWZ, SWZ);
      constexpr int atom_bytes = kAtomN * 2;
      constexpr uint32_t hiB = B_MN ? umma_desc_hi(8 * atom_bytes, atom_bytes)
                                    : umma_desc_hi(8 * SWZ, SWZ);
      constexpr uint32_t lboB = B_MN ? KC * atom_bytes : 0;
      constexpr uint32_t kstepB = B_MN ? (16 * atom_bytes) >> 4 : 2;   // per UMMA_K, in 16-B units
      const uint32_t row16 = SWZ >> 4;                                  // one A row, 16-B units
      int sa = 0, sb = 0, as = 0;
      uint32_t pa = 0, pb = 0, aphase = 0;
      bool first_tile = true;
      int ti = 0;
      for (int tile = first_unit; tile < total_tiles; tile += unit_step, ++ti) {
        const int m0 = tile_m0(tile);
        const int a_off = P.flat ? 0 : m0 - (m0 / P.Wp) * P.Wp;
        if (lane == 0) prof_mark(P.prof, 1, ti, 0);
        mbar_wait(&tempty[as], aphase ^ 1u);
        tc_fence_after();
        if (lane == 0) prof_mark(P.prof, 1, ti, 1);
        const uint32_t tmem_d = tmem_base + as * BN;
        if (P.b_resident) sb = 0;
        uint32_t acc = 0;
        for (int j = 0; j < chunks; ++j) {
          mbar_wait(&a_full[sa], pa);
          tc_fence_after();
          if (j == 0 && lane == 0) prof_mark(P.prof, 1, ti, 2);
          const uint32_t a_lo0 =
              umma_desc_lo(smem_u32(smem + sa * P.a_stage_bytes), 0) + (uint32_t)a_off * row16;
          uint32_t a_row = a_lo0;                       // tap (r, 0)
          if (TPS == 3) {
            // one weight stage per filter row: one wait, one election, twelve MMAs, one commit
            for (int r = 0; r < P.kh; ++r, a_row += (uint32_t)P.Wp * row16) {
              mbar_wait(&b_full[sb], pb);
              tc_fence_after();
              const uint32_t b_row0 = umma_desc_lo(smem_u32(smem_b + sb * kStageB), lboB);
              if (elect_one()) {
#pragma unroll
                for (int s3 = 0; s3 < 3; ++s3) {
#pragma unroll
                  for (int kk = 0; kk < KC / 16; ++kk) {
                    umma_f16(tmem_d, umma_desc_pack(hiA, a_row + s3 * row16 + kk * 2),
                             umma_desc_pack(hiB, b_row0 + s3 * (kBBytes >> 4) + kk * kstepB), idesc,
                             acc);
                    acc = 1;
                  }
                }
                umma_commit(&b_empty[sb]);
              }
              __syncwarp();
              acc = 1;
              if (++sb == P.SB) { sb = 0; pb ^= 1u; }
            }
          } else {
          for (int r = 0; r < P.kh; ++r, a_row += (uint32_t)P.Wp * row16) {
            uint32_t a_tap = a_row;
            for (int s = 0; s < P.kw; ++s, a_tap += row16) {
              if (!P.b_resident || first_tile) {
                mbar_wait(&b_full[sb], pb);
                tc_fence_after();
              }
              const uint32_t b_lo0 = umma_desc_lo(smem_u32(smem_b + sb * kBBytes), lboB);
              if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < KC / 16; ++kk) {
                  umma_f16(tmem_d, umma_desc_pack(hiA, a_tap + kk * 2),
                           umma_desc_pack(hiB, b_lo0 + kk * kstepB), idesc, acc);
                  acc = 1;
                }
                if (!P.b_resident) umma_commit(&b_empty[sb]);
#if SEGB200_KERNEL_PROF
                // per-tap issue timestamps of tiles 2 and 3 (role 3 of the timeline hook)
                if (P.prof != nullptr && blockIdx.x == 0 && (ti == 2 || ti == 3) && j == 0)
                  P.prof[3 * kProfTiles * 4 + (ti - 2) * 16 + r * P.kw + s] = clock64();
#endif
              }
              __syncwarp();
              acc = 1;
              if (++sb == P.SB) { sb = 0; pb ^= 1u; }
            }
          }
          }
          if (elect_one()) umma_commit(&a_empty[sa]);
          __syncwarp();
          if (++sa == P.SA) { sa = 0; pa ^= 1u; }
        }
        if (elect_one()) umma_commit(&tfull[as]);
        __syncwarp();
        if (lane == 0) prof_mark(P.prof, 1, ti, 3);
        if (++as == 2) { as = 0; aphase ^= 1u; }
        first_tile = false;
      }
    }
  } else {
    // ============================= epilogue =============================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;              // two epilogue warps per TMEM lane quadrant
    int as = 0;
    uint32_t aphase = 0;
    const int img_stride = P.Hp * P.Wp;
    int ti = 0;
    long long* eprof = (warp == 2 && lane == 0) ? P.prof : nullptr;
    for (int tile = first_unit; tile < total_tiles; tile += unit_step, ++ti) {
      const int m0 = tile_m0(tile);
      const int n0 = (tile % n_tiles) * BN;
      const int m = m0 + quad * 32 + lane;
      const int img = m / img_stride;
      const int rem = m - img * img_stride;
      const int yp = rem / P.Wp;
      const int xp = rem - yp * P.Wp;
      const bool row_ok = img < P.batch && yp < P.Ho && xp < P.Wo;
      const bool second = P.d1.ptr != nullptr && n0 >= P.split_n;
      const EpiDest& D = second ? P.d1 : P.d0;
      const int nl0 = second ? n0 - P.split_n : n0;
      const int64_t off = row_ok ? img * D.sn + yp * D.sh + xp * D.sw : 0;
      const int64_t moff = row_ok ? img * D.msn + yp * D.msh + xp * D.msw : 0;
      prof_mark(eprof, 2, ti, 0);
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      prof_mark(eprof, 2, ti, 1);
#pragma unroll 1
      for (int cc = half * (BN >= 32 ? 32 : 16); cc < BN; cc += 2 * (BN >= 32 ? 32 : 16)) {
        constexpr int W = BN >= 32 ? 32 : 16;
        uint32_t r[W];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN + cc;
        if (W == 32) tmem_ld_32x32(taddr, r); else tmem_ld_32x16(taddr, r);
        tmem_ld_wait();
        if (row_ok) {
          const int ncol = nl0 + cc;
          float v[W];
#pragma unroll
          for (int j = 0; j < W; ++j) {
            v[j] = __uint_as_float(r[j]);
            if ((P.flags & SEG_EPI_BIAS) && ncol + j < D.cols) v[j] += __ldg(P.bias + ncol + j);
            if (P.flags & SEG_EPI_RELU) v[j] = fmaxf(v[j], 0.f);
          }
          if (P.post_scale != nullptr) {
#pragma unroll
            for (int j = 0; j < W; ++j)
              v[j] = fmaf(v[j], __ldg(P.post_scale + ncol + j), __ldg(P.post_shift + ncol + j));
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
      prof_mark(eprof, 2, ti, 2);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      prof_mark(eprof, 2, ti, 3);
      if (++as == 2) { as = 0; aphase ^= 1u; }
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
