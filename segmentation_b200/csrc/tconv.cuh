// Spatial-tile tcgen05 convolution (3x3, stride 1): conv fwd and conv dgrad.
//
// One CTA tile = 16 output rows x (8*MT) output columns of one image.  The producer
// fetches the input halo box [18][8*MT+2][KC] with ONE 4-D TMA load per channel chunk
// (out-of-bounds coordinates are zero-filled: SAME padding, full padding of dgrad and
// ragged image edges cost nothing).  In shared memory the box is a list of pixel rows
// of KC channels; the A operand of filter tap (r, s) for column block mb is the same
// box read through a K-major UMMA descriptor whose start address is advanced by
// ((r*PW + s + 8*mb) pixel rows and whose 8-row-group stride (SBO) is one box row
// (PW pixels): group g = output row y, rows inside the group = 8 consecutive output
// columns.  (The smem swizzle is a function of the absolute address, so shifted /
// strided starts read what TMA wrote — tools/probe_shift.py.)  All 9 taps, all MT
// column blocks and all KC/16 k-steps of a chunk are issued as straight-line
// tcgen05.mma by one elected lane (a rolled tap loop costs ~150 cycles per tap in
// issue latency — tools/probe_rate.py, tools/layer_prof.py).
//
// Epilogue: TMEM -> registers -> bias/ReLU (or ReLU-grad mask) -> bf16 -> swizzled
// smem staging -> TMA tensor store, one [4 rows][8 cols][<=64 ch] box per epilogue warp;
// the store clips at the tensor bounds, so ragged tiles and padded channels need no
// predicates.  The ReLU-grad mask tile is fetched by TMA into the same layout.  (Plain
// LSU stores — direct or smem-staged and coalesced — were measured to slow the
// concurrent UMMA stream by ~50 %; TMA stores do not.)
#pragma once
#include "umma_conv.cuh"
#include "hconv.cuh"

namespace segb {

struct TconvParams {
  int tiles_x, tiles_y, batch;
  int n_tiles;                 // N_total / BN
  int chunks1, chunks2;        // KC-channel chunks taken from source 1 / source 2
  int pad_t, pad_l;            // input coordinate of output (0,0), tap (0,0) is (-pad_t, -pad_l)
  int tap_flip, b_rows_per_tap;
  int SA, SB, b_resident;
  int a_stage_bytes;
  int off_b, off_stage, off_mask, off_bias, off_bars;   // byte offsets in the aligned smem
  int split_n;                 // columns >= split_n go to destination 1 (0: single destination)
  int nstg;                    // staging boxes per epilogue warp (2..4)
  const float* bias;
  int bias_cols;
  int n_total;
  int flags;
  long long* prof;             // optional in-kernel timeline of CTA 0 (test hook), else null
};

constexpr int kTconvMaxSA = 8;
constexpr int kTconvMaxSB = 40;
constexpr int kTconvTH = 16;
// Epilogue warps per TMEM lane quadrant.  The drain of one 32-column chunk (TMEM load, bias by
// shuffle, ReLU / mask, bf16, staging box, TMA store) is ~450 dependent instructions that one
// warp issues at ~0.2 IPC (ncu: 41 % of its samples in fixed-latency waits, 12 % instruction
// fetch).  In steady state it hides behind the next tile's MMAs (in-kernel timeline,
// profiles/r02_ncu_tconv.md), but it is exposed after a CTA's last tile and before the next
// kernel can start: four warps per quadrant (16 epilogue warps, 4 per scheduler) instead of
// two interleave the chains and shorten that tail - U-Net step 0.934 -> 0.926 ms.
constexpr int kTconvEW = 4;
constexpr int kTconvThreads = 64 + 128 * kTconvEW;   // producer, MMA issuer, 4 * kTconvEW epilogue warps

template <int KC, int BN, bool B_MN, int MT>
__global__ void __launch_bounds__(kTconvThreads, 1)
tconv_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
             const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD0,
             const __grid_constant__ CUtensorMap tmD1, const __grid_constant__ CUtensorMap tmM0,
             const __grid_constant__ CUtensorMap tmM1, const TconvParams P) {
  constexpr int SWZ = KC * 2;                       // bytes per staged pixel row
  constexpr int TW = 8 * MT;
  constexpr int PW = TW + 2, PH = kTconvTH + 2;     // halo box (pixels)
  constexpr uint32_t kABoxBytes = PW * PH * SWZ;
  constexpr int kBBytes = BN * KC * 2;
  constexpr int kAtomN = BN < 64 ? BN : 64;
  constexpr int BNH = BN < 64 ? BN : 64;            // channels per store box
  constexpr int NH = BN / BNH;
  constexpr int kHalfBytes = 32 * BNH * 2;          // one warp's [4][8][BNH] store box
  constexpr int kStgBytes = NH * kHalfBytes;
  constexpr int kAccCols = MT * BN;
  constexpr int kTmemCols = 2 * kAccCols <= 32 ? 32 : 2 * kAccCols <= 64 ? 64
                            : 2 * kAccCols <= 128 ? 128 : 2 * kAccCols <= 256 ? 256 : 512;
  static_assert(2 * kAccCols <= 512, "accumulator stages exceed TMEM");
  static_assert(BN % 32 == 0, "BN must be a multiple of 32");

  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_b = smem + P.off_b;
  float* s_bias = reinterpret_cast<float*>(smem + P.off_bias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P.off_bars);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kTconvMaxSA;
  uint64_t* b_full = a_empty + kTconvMaxSA;
  uint64_t* b_empty = b_full + kTconvMaxSB;
  uint64_t* tfull = b_empty + kTconvMaxSB;
  uint64_t* tempty = tfull + 2;
  uint64_t* mask_bar = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mask_bar + 4 * kTconvEW);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int chunks = P.chunks1 + P.chunks2;
  const int m_tiles = P.batch * P.tiles_y * P.tiles_x;
  const int total_tiles = m_tiles * P.n_tiles;
  const bool has_mask = (P.flags & SEG_EPI_RELU_MASK) != 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD0);
    tma_prefetch_desc(&tmD1);
    for (int i = 0; i < 4 * kTconvEW; ++i) mbar_init(&mask_bar[i], 1);
    for (int i = 0; i < P.SA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < P.SB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4 * (MT * (BN / 32) < kTconvEW ? MT * (BN / 32) : kTconvEW)); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  pdl_wait();                 // everything above overlaps the previous kernel's tail
  // bias vector -> smem (zero beyond the valid columns)
  for (int i = threadIdx.x; i < P.n_total; i += blockDim.x)
    s_bias[i] = ((P.flags & SEG_EPI_BIAS) && i < P.bias_cols) ? __ldg(P.bias + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (elect_one()) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      bool first_tile = true;
      int ti = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
        prof_mark(P.prof, 0, ti, 0);
        const int n0 = (tile % P.n_tiles) * BN;
        int mt = tile / P.n_tiles;
        const int tx = mt % P.tiles_x;
        mt /= P.tiles_x;
        const int ty = mt % P.tiles_y;
        const int img = mt / P.tiles_y;
        const int xin = tx * TW - P.pad_l;
        const int yin = ty * kTconvTH - P.pad_t;
        for (int j = 0; j < chunks; ++j) {
          const bool second = j >= P.chunks1;
          const CUtensorMap* tm = second ? &tmA2 : &tmA1;
          const int c0 = (second ? j - P.chunks1 : j) * KC;
          mbar_wait(&a_empty[sa], pa ^ 1u);
          if (j == 0) prof_mark(P.prof, 0, ti, 1);
          mbar_expect_tx(&a_full[sa], kABoxBytes);
          tma_load_4d(tm, &a_full[sa], smem + sa * P.a_stage_bytes, c0, xin, yin, img);
          if (++sa == P.SA) { sa = 0; pa ^= 1u; }
          if (!P.b_resident || first_tile) {
            int bt = P.tap_flip ? 8 : 0;
            for (int t = 0; t < 9; ++t) {
              mbar_wait(&b_empty[sb], pb ^ 1u);
              uint8_t* sbp = smem_b + sb * kBBytes;
              mbar_expect_tx(&b_full[sb], kBBytes);
              if (B_MN) {
                const int row = bt * P.b_rows_per_tap + j * KC;
#pragma unroll
                for (int a = 0; a < BN / kAtomN; ++a)
                  tma_load_2d(&tmB, &b_full[sb], sbp + a * (KC * kAtomN * 2), n0 + a * kAtomN, row);
              } else {
                tma_load_2d(&tmB, &b_full[sb], sbp, j * KC, bt * P.b_rows_per_tap + n0);
              }
              bt += P.tap_flip ? -1 : 1;
              if (++sb == P.SB) { sb = 0; pb ^= 1u; }
            }
          }
        }
        first_tile = false;
        prof_mark(P.prof, 0, ti, 2);
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BN, 0, B_MN ? 1 : 0);
      constexpr uint32_t hiA = umma_desc_hi(PW * SWZ, SWZ);     // SBO = one box row
      constexpr int atom_bytes = kAtomN * 2;
      constexpr uint32_t hiB = B_MN ? umma_desc_hi(8 * atom_bytes, atom_bytes)
                                    : umma_desc_hi(8 * SWZ, SWZ);
      constexpr uint32_t lboB = B_MN ? KC * atom_bytes : 0;
      constexpr uint32_t kstepB = B_MN ? (16 * atom_bytes) >> 4 : 2;
      int sa = 0, sb = 0, as = 0;
      uint32_t pa = 0, pb = 0, aphase = 0;
      bool first_tile = true;
      // The barriers of the NEXT step (next chunk's A stage; at a tile boundary also the
      // next accumulator stage) are waited for in the middle of the current step, while
      // the tensor pipe still has queued work: a wait costs ~100-200 cycles even when the
      // barrier is already complete, and the MMA queue is only a few instructions deep.
      // With a streamed B ring the producer reaches the next A stage only after all nine
      // B tiles of this chunk, which need this chunk's b_empty commits: pre-wait after the
      // last tap there (after tap 4 would deadlock), mid-step when B is resident.
      bool pre_a = false, pre_t = false;
      const int pre_tap = P.b_resident ? 4 : 8;
      int ti = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
        prof_mark(P.prof, 1, ti, 0);
        if (!pre_t) {
          mbar_wait(&tempty[as], aphase ^ 1u);
          tc_fence_after();
        }
        prof_mark(P.prof, 1, ti, 1);
        const uint32_t tmem_d = tmem_base + as * kAccCols;
        const bool more_tiles = tile + (int)gridDim.x < total_tiles;
        if (P.b_resident) sb = 0;
        for (int j = 0; j < chunks; ++j) {
          if (!pre_a) {
            mbar_wait(&a_full[sa], pa);
            tc_fence_after();
          }
          pre_a = pre_t = false;
          if (j == 0) prof_mark(P.prof, 1, ti, 2);
          const uint32_t a0 = umma_desc_lo(smem_u32(smem + sa * P.a_stage_bytes), 0);
          int sa_n = sa + 1;
          uint32_t pa_n = pa;
          if (sa_n == P.SA) { sa_n = 0; pa_n ^= 1u; }
#pragma unroll
          for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              if (!P.b_resident || first_tile) {
                mbar_wait(&b_full[sb], pb);
                tc_fence_after();
              }
              const uint32_t b0 = umma_desc_lo(smem_u32(smem_b + sb * kBBytes), lboB);
#pragma unroll
              for (int mb = 0; mb < MT; ++mb) {
#pragma unroll
                for (int kk = 0; kk < KC / 16; ++kk) {
                  const uint32_t a_off = (uint32_t)(((r * PW + s + 8 * mb) * SWZ + kk * 32) >> 4);
                  umma_f16(tmem_d + mb * BN, umma_desc_pack(hiA, a0 + a_off),
                           umma_desc_pack(hiB, b0 + kk * kstepB), idesc,
                           (j | r | s | kk) != 0 ? 1u : 0u);
                }
              }
              if (!P.b_resident) umma_commit(&b_empty[sb]);
              if (++sb == P.SB) { sb = 0; pb ^= 1u; }
              if (r * 3 + s == pre_tap) {
                if (j + 1 < chunks) {
                  mbar_wait(&a_full[sa_n], pa_n);
                  pre_a = true;
                } else if (more_tiles) {
                  mbar_wait(&tempty[as ^ 1], (as == 1 ? aphase ^ 1u : aphase) ^ 1u);
                  mbar_wait(&a_full[sa_n], pa_n);
                  pre_a = pre_t = true;
                }
                tc_fence_after();
              }
            }
          }
          umma_commit(&a_empty[sa]);
          sa = sa_n;
          pa = pa_n;
        }
        umma_commit(&tfull[as]);
        prof_mark(P.prof, 1, ti, 3);
        if (++as == 2) { as = 0; aphase ^= 1u; }
        first_tile = false;
      }
    }
  } else {
    // ============================= epilogue =============================
    // Everything that leaves or enters the SM here goes through the async proxy (TMA):
    // LSU global stores issued next to a running UMMA stream were measured to slow the
    // kernel, TMA traffic does not.  Eight warps, two per TMEM lane quadrant, take the
    // 32-column chunks of a tile alternately (a single warp per scheduler issues only one
    // instruction per ~5 cycles in this dependent code, the TMEM drain itself needs ~290
    // cycles per chunk).  Each thread owns one output pixel (TMEM lane): bias (one column
    // per lane, broadcast by shuffle), ReLU or ReLU-grad mask, bf16, then its four 16-byte
    // chunks go to a swizzled [4][8][32] staging box that one lane hands to a TMA tensor
    // store; kStg boxes per warp rotate so that a store's smem-read completion (~1000
    // cycles behind its commit) is never waited for, and the TMEM load of the warp's next
    // chunk is in flight while the current one is processed.
    const int quad = warp & 3;                       // TMEM lanes [32*quad, 32*quad+32)
    const int sub = (warp - 2) >> 2;                 // which warp of the quadrant's group
    constexpr int NCH = BN / 32;                     // 32-column chunks per column block
    constexpr int NLD = MT * NCH;                    // chunks (TMEM loads) per tile
    constexpr int kBoxBytes = 32 * 32 * 2;           // [4 rows][8 cols][32 ch] bf16
    constexpr int kMyMax = (NLD + kTconvEW - 1) / kTconvEW;   // chunks per warp per tile
    if (sub < NLD) {
      uint8_t* stg = smem + P.off_stage + (warp - 2) * (P.nstg * kBoxBytes);
      const uint32_t msk_u32 = smem_u32(smem + P.off_mask + (warp - 2) * (kMyMax * kBoxBytes));
      // swizzled (64-byte rows) offset of this lane's pixel, 16-byte chunk q: ^ ((lane>>1)&3)
      const uint32_t row_off = (uint32_t)lane * 64;
      const uint32_t xor_sel = (uint32_t)((lane >> 1) & 3);
      int as = 0, sbuf = 0;
      uint32_t aphase = 0, mphase = 0;
      int ti = 0;
      long long* eprof = (warp == 2 && lane == 0) ? P.prof : nullptr;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
        prof_mark(eprof, 2, ti, 0);
        const int n0 = (tile % P.n_tiles) * BN;
        int mt = tile / P.n_tiles;
        const int tx = mt % P.tiles_x;
        mt /= P.tiles_x;
        const int ty = mt % P.tiles_y;
        const int img = mt / P.tiles_y;
        const int x0 = tx * TW;
        const int yw = ty * kTconvTH + 4 * quad;     // first output row of this warp's boxes
        const bool second = P.split_n > 0 && n0 >= P.split_n;
        const CUtensorMap* tmD = second ? &tmD1 : &tmD0;
        const CUtensorMap* tmM = second ? &tmM1 : &tmM0;
        const int nl0 = second ? n0 - P.split_n : n0;   // column inside the destination
        if (has_mask && lane == 0) {
          int cnt = 0;
#pragma unroll
          for (int k = 0; k < kMyMax; ++k)
            if (sub + k * kTconvEW < NLD) ++cnt;
          mbar_expect_tx(&mask_bar[warp - 2], cnt * kBoxBytes);
#pragma unroll
          for (int k = 0; k < kMyMax; ++k) {
            const int i = sub + k * kTconvEW;
            if (i < NLD)
              tma_load_4d(tmM, &mask_bar[warp - 2], smem + P.off_mask +
                              ((warp - 2) * kMyMax + k) * kBoxBytes,
                          nl0 + 32 * (i % NCH), x0 + 8 * (i / NCH), yw, img);
          }
        }
        // bias of this tile's columns: lane l holds column 32*c + l of chunk c
        float bl[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) bl[c] = lds32f(smem_u32(s_bias + n0 + 32 * c + lane));
        mbar_wait(&tfull[as], aphase);
        tc_fence_after();
        prof_mark(eprof, 2, ti, 1);
        const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + as * kAccCols;
        uint32_t rr[kMyMax > 1 ? 2 : 1][32];
        tmem_ld_32x32(tbase + (sub / NCH) * BN + (sub % NCH) * 32, rr[0]);
        if (has_mask) {
          mbar_wait(&mask_bar[warp - 2], mphase);
          mphase ^= 1u;
        }
#pragma unroll
        for (int k = 0; k < kMyMax; ++k) {
          const int i = sub + k * kTconvEW;          // this warp's chunk (runtime: sub)
          if (i < NLD) {
            const int mb = i / NCH, c = i % NCH;
            // rotate to the next staging box; the store that last read it was committed
            // nstg boxes ago: allow the nstg-1 most recent groups to be still reading
            uint8_t* sbp = stg + sbuf * kBoxBytes;
            if (lane == 0) {
              if (P.nstg >= 4) bulk_wait_group_read<3>();
              else if (P.nstg == 3) bulk_wait_group_read<2>();
              else if (P.nstg == 2) bulk_wait_group_read<1>();
              else bulk_wait_group_read<0>();
            }
            __syncwarp();
            tmem_ld_wait();
            if (k + 1 < kMyMax && i + kTconvEW < NLD)
              tmem_ld_32x32(tbase + ((i + kTconvEW) / NCH) * BN + ((i + kTconvEW) % NCH) * 32,
                            rr[(k + 1) & (kMyMax > 1 ? 1 : 0)]);
            const uint32_t* r = rr[k & (kMyMax > 1 ? 1 : 0)];
            float bias_l = bl[0];
#pragma unroll
            for (int cc = 1; cc < NCH; ++cc)
              if (cc == c) bias_l = bl[cc];
            const uint32_t sb32 = smem_u32(sbp) + row_off;
            const uint32_t mk32 = msk_u32 + k * kBoxBytes + row_off;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                v[e] = __uint_as_float(r[q * 8 + e]) + __shfl_sync(0xffffffffu, bias_l, q * 8 + e);
                if (P.flags & SEG_EPI_RELU) v[e] = fmaxf(v[e], 0.f);
              }
              const uint32_t off = ((uint32_t)q ^ xor_sel) << 4;
              if (has_mask) {
                const uint4 u = lds128(mk32 + off);
                const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (!(bf16_lo(w4[e]) > 0.f)) v[2 * e] = 0.f;
                  if (!(bf16_hi(w4[e]) > 0.f)) v[2 * e + 1] = 0.f;
                }
              }
              uint4 o;
              o.x = pack_bf16x2(v[0], v[1]);
              o.y = pack_bf16x2(v[2], v[3]);
              o.z = pack_bf16x2(v[4], v[5]);
              o.w = pack_bf16x2(v[6], v[7]);
              sts128(sb32 + off, o);
            }
            if (i + kTconvEW >= NLD) {               // this warp's share of the stage is read
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty[as]);
              prof_mark(eprof, 2, ti, 2);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(tmD, sbp, nl0 + 32 * c, x0 + 8 * mb, yw, img);
              bulk_commit_group();
            }
            if (++sbuf == P.nstg) sbuf = 0;
          }
        }
        prof_mark(eprof, 2, ti, 3);
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
      if (lane == 0) bulk_wait_group<0>();
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
