// First-layer tcgen05 convolution (3x3, stride 1, RGB input): forward and weight gradient,
// each optionally fused with the 2x2/2 max-pool that follows the layer.
//
// The first layer of every reference graph convolves 3 input channels (models/unet.py:111,
// models/fcn.py:110): arithmetic intensity 25 flop/B, i.e. HBM-bound, and with the input
// padded to 16 channels for the generic spatial-tile kernels it was also TMA-row bound
// (32-byte pixel rows) - 10 % of the U-Net step for 0.5 % of its FLOPs.  Here the input
// lives in HBM as 4 bf16 per pixel, (R, G, B, 1): 8 bytes.  Four builder warps gather the
// 3x3x4 patch of each of a tile's 128 output pixels with nine 8-byte cp.async copies per
// pixel (no register staging, so several tiles are in flight per CTA; L1/L2 serve the
// 9-fold reuse) straight into the canonical K-major SWIZZLE_128B operand layout in shared
// memory - one 128-byte row per pixel, k-slot = r*12 + s*4 + c (c = 3: weight rows zero),
// slots 36 and 37 a constant 1 whose weight rows hold the bias split into two bf16 (hi + lo:
// the bias add costs the epilogue nothing) - so a tile is ONE 128 x BN x 48 MMA group and
// nothing padded is ever read from or written to HBM.
//
// Plain forward: a tile is 128 consecutive output pixels (flattened over n, y, x), i.e. 128
// consecutive rows of the dense [pixels][BN] activation: TMEM -> relu + bf16 (one cvt per
// pair) -> one TMA store box per epilogue warp; eight epilogue warps, the two of a TMEM lane
// quadrant take tiles alternately.
//
// Forward + max-pool (POOL; models/unet.py:111-120, models/fcn.py:110-117): a tile is two
// output rows x 64 columns, pixel (2Y+dy, 64tx + 2xp + dx) in accumulator row
// 4xp + 2dy + dx, so a 2x2 pool window is four adjacent TMEM lanes: the warp transposes
// through a small shared-memory staging area (lane k of a window reads 16-byte chunk k of
// its four pixels), scans the four candidates on packed bf16 pairs for the maximum and the
// slot of the first maximum (row-major scan order, as seg_maxpool_fwd) and writes chunk k of
// the pooled pixel and its 8 slot bytes (a butterfly-shuffle version that kept every lane's
// whole pixel cost ~500 instructions per warp and tile - 41 us; this one ~100).  The
// full-resolution
// activation is written only inside a caller-given window (U-Net: the crop conv1_2 reads;
// FCN: nothing) - 66 MB of stores and the pool kernel's 66 MB of loads disappear.
//
// Weight gradient: the same patch tile read MN-major (pixels are the GEMM K axis) times the
// dZ tile, accumulated in TMEM over all of the CTA's tiles:
// dW[k-slot][co] += sum_px patch[px][k-slot] * dZ[px][co]; the constant-1 slot 36
// accumulates sum_px dZ - the bias gradient - for free.  One coalesced red.add of 28 x BN
// floats per CTA at the end.  The dZ tile is one TMA load, or (POOL) is BUILT in shared
// memory from the pooled gradient, the argmax slots and the ReLU mask (the input at the
// argmax IS the pooled value) - the max-pool backward fused into the operand producer, so
// the full-resolution gradient never exists in HBM.  (A gradient arriving from a second
// consumer of the activation - U-Net's conv1_2 window - is a separate, linear term: the
// caller runs the plain weight gradient on that window.)
#pragma once
#include "umma_conv.cuh"

namespace segb {

struct FconvParams {
  const uint2* x4;          // pixels of 4 bf16 (R, G, B, 1): pixel (n, y, x) at x4[n*x_sn + y*x_sh + x]
  int64_t x_sn, x_sh;       //   (a crop of the staged input is a view: pointer offset + strides)
  int H, W, Ho, Wo;
  int pad_t, pad_l;
  int M_total;              // N * Ho * Wo (< 2^30)
  int tiles;
  // exact division by multiply-high + shift (values < 2^31); mul == 0: divisor 1.
  //   plain: a = Wo, b = Ho*Wo (pixel index -> n, oy, ox)
  //   POOL:  a = tiles_x, b = Hp*tiles_x (tile index -> n, Y, tx)
  uint32_t div_a, div_a_mul, div_a_shr;
  uint32_t div_b, div_b_mul, div_b_shr;
  const bf16* w;            // bf16 shadow [3][3][cin_pad][cout_pad]
  int cin_pad, cout_pad, cout;
  const float* bias;
  int flags;
  float* dw;                // fp32 master layout [3][3][3][cout]
  float* db;                // nullable
  // ---- POOL variants (BN = 32)
  int Hp, Wp;               // pooled grid = Ho/2 x Wo/2
  bf16* y;                  // forward: full-resolution activation, written inside the window
  int64_t y_sn, y_sh, y_sw;
  int win_y0, win_x0, win_y1, win_x1;   // the window, output coordinates [y0,y1) x [x0,x1)
  bf16* pooled;             // dense [N][Hp][Wp][BN]: pool output (forward) / mask source (wgrad)
  uint8_t* amax;            // dense [N][Hp][Wp][BN] window slots
  const bf16* dpool;        // dense [N][Hp][Wp][BN] gradient w.r.t. the pool output
  // pooled forward, inference: slim.batch_norm with the moving statistics between the ReLU and
  // the pool (models/deconvolution.py:109-118), y = (relu(conv) - mean) * rsqrt(var + eps) +
  // beta, each stage rounded to bf16 as the unfused entries store it.  Null: none.
  const float* bn_mean; const float* bn_var; const float* bn_beta; float bn_eps;
};

constexpr int kFcThreads = 448;          // MMA issuer, alloc/TMA warp, 8 epilogue, 4 builder warps
constexpr int kFcABytes = 128 * 128;     // one K-block of a patch tile: 128 pixels x 128 bytes
constexpr int kFcRows = 37;              // accumulator rows of the weight gradient that are used
constexpr int kFcRawDepth = 4;           // pooled weight gradient: raw tiles in flight per thread
constexpr int kFcRawBytes = 128 * 48;    // dpool 16 B + pooled 16 B + slots 8 B (+8) per thread
constexpr int kFcPoolStg = 32 * 80;      // pooled forward: one warp's transpose staging (80-byte rows)

// KH x KH filter taps, stride ST.  k-slot of (tap, channel) = tap*4 + c; the constant-1
// slots (bias hi / lo) follow the taps; K is padded to a multiple of 16 and split into
// K-blocks of 64 slots (one 128-byte SWIZZLE_128B row per pixel and block).
template <int BN, bool WGRAD, bool POOL, int KH = 3, int ST = 1>
struct FconvCfg {
  static constexpr int kOne = KH * KH * 4;                       // first constant-1 slot
  static constexpr int kKpad = (kOne + 2 + 15) / 16 * 16;        // 48 (3x3), 112 (5x5)
  static constexpr int kKB = (kKpad + 63) / 64;                  // K-blocks per tile
  static constexpr int kABytes = kKB * kFcABytes;                // one patch-tile stage
  static constexpr int S = (WGRAD && POOL) ? 3 : (kKB > 1 ? 2 : 4);   // pipeline stages
  static constexpr int rowB = BN * 2;
  static constexpr int kZBytes = 128 * rowB;            // one dZ tile (wgrad)
  static constexpr int kWBytes = kKB * BN * 128;        // weights, K-major 128-byte rows (fwd)
  static constexpr int kStgBytes = 32 * rowB;           // one epilogue warp's store box (fwd)
  static constexpr int kOffRaw = S * kABytes + S * kZBytes;
  static constexpr int kOffBars =
      S * kABytes + (WGRAD ? S * kZBytes + (POOL ? kFcRawDepth * kFcRawBytes : 0)
                           : kWBytes + (POOL ? 8 * kFcPoolStg : 16 * kStgBytes));
  static constexpr int kSmemBytes = kOffBars + 256 + 1024 /*base alignment*/;
  static constexpr int kTmemCols = WGRAD ? (BN < 32 ? 32 : BN) : 2 * BN;
  static_assert(!WGRAD || kKB == 1, "weight gradient: 3x3 only");
};

__device__ __forceinline__ uint32_t fc_div(uint32_t x, uint32_t mul, uint32_t shr) {
  return mul ? (__umulhi(x, mul) >> shr) : x;
}
// per-half (a > b) ? 0xffff : 0 on packed bf16 pairs
__device__ __forceinline__ uint32_t bf16x2_gt_mask(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ void cp_async_16(uint32_t sdst, const void* gsrc, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdst), "l"(gsrc),
               "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int BN, bool WGRAD, bool POOL, int KH = 3, int ST = 1>
__global__ void __launch_bounds__(kFcThreads, 2)
fconv_kernel(const __grid_constant__ CUtensorMap tmIO, const FconvParams P) {
  using Cfg = FconvCfg<BN, WGRAD, POOL, KH, ST>;
  constexpr int S = Cfg::S;
  constexpr int kAB = Cfg::kABytes;
  constexpr int rowB = Cfg::rowB;
  static_assert(BN == 32 || BN == 64, "first-layer kernel: 32 or 64 output channels per tile");
  static_assert(!POOL || BN == 32, "pooled variants: 32 output channels");

  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* a_ring = smem;
  uint8_t* z_ring = smem + S * kAB;                     // wgrad
  uint8_t* raw_ring = smem + Cfg::kOffRaw;              // wgrad + pool
  uint8_t* w_smem = smem + S * kAB;                     // fwd
  uint8_t* stg = w_smem + Cfg::kWBytes;                 // fwd (plain)
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
  auto tile_of = [&](int i) { return (int)blockIdx.x + i * (int)gridDim.x; };
  // POOL: tile -> (image, pooled row, column block)
  auto tile_pos = [&](int tile, int& n, int& Y, int& tx) {
    n = (int)fc_div((uint32_t)tile, P.div_b_mul, P.div_b_shr);
    const uint32_t rem = (uint32_t)tile - (uint32_t)n * P.div_b;
    Y = (int)fc_div(rem, P.div_a_mul, P.div_a_shr);
    tx = (int)(rem - (uint32_t)Y * P.div_a);
  };

  if (warp == 0 && lane == 0) {
    if (!POOL) tma_prefetch_desc(&tmIO);
    for (int i = 0; i < S; ++i) {
      mbar_init(&a_full[i], 128);
      mbar_init(&a_empty[i], 1);
      mbar_init(&z_full[i], POOL ? 4 : 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  // initialise the patch ring once: the builders only ever write the first 72 bytes of a row
  // (k-slots 0..35); slots 36 and 37 are the constant 1, slots 38..47 (read by the third
  // k-step) stay zero
  for (int i = threadIdx.x; i < S * kAB / 16; i += kFcThreads)
    sts128(smem_u32(a_ring) + i * 16, make_uint4(0u, 0u, 0u, 0u));
  __syncthreads();
  {
    // the two constant-1 slots: byte kOne*2 of the pixel's K row = K-block kOne/64, 16-byte
    // chunk (kOne%64)/8 (swizzled), second half
    constexpr uint32_t kBlk = Cfg::kOne / 64, kChunk = (Cfg::kOne % 64) / 8;
    static_assert((Cfg::kOne % 8) == 4, "constant-1 slots sit in the second half of a chunk");
    for (int i = threadIdx.x; i < S * 128; i += kFcThreads) {
      const uint32_t st_ = (uint32_t)i / 128u, t_ = (uint32_t)i % 128u;
      const uint32_t row = smem_u32(a_ring) + st_ * kAB + kBlk * kFcABytes + t_ * 128u;
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(row + ((kChunk ^ (t_ & 7u)) << 4) + 8u),
                   "r"(0x3F803F80u) : "memory");
    }
  }
  pdl_wait();                 // everything above overlaps the previous kernel's tail
  if (!WGRAD) {
    // weights -> K-major SWIZZLE_128B rows: row = output channel, k-slot = r*12 + s*4 + c
    for (int idx = threadIdx.x; idx < BN * Cfg::kKB * 64; idx += kFcThreads) {
      const int co = idx / (Cfg::kKB * 64), k = idx % (Cfg::kKB * 64);
      const int tap = k >> 2, c = k & 3;
      bf16 v = __float2bfloat16(0.f);
      if (k < Cfg::kOne && c < 3 && co < P.cout_pad) {
        v = P.w[((int64_t)tap * P.cin_pad + c) * P.cout_pad + co];
      } else if ((k == Cfg::kOne || k == Cfg::kOne + 1) && (P.flags & SEG_EPI_BIAS) && co < P.cout) {
        const float bv = __ldg(P.bias + co);
        const bf16 hi = __float2bfloat16(bv);
        v = k == Cfg::kOne ? hi : __float2bfloat16(bv - __bfloat162float(hi));
      }
      const int blk = k >> 6, kk = k & 63;
      *reinterpret_cast<bf16*>(w_smem + blk * (BN * 128) + co * 128 +
                               (((kk >> 3) ^ (co & 7)) << 4) + (kk & 7) * 2) = v;
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
          const uint32_t a0 = umma_desc_lo(smem_u32(a_ring + s * kAB), 0);
#pragma unroll
          for (int kk = 0; kk < Cfg::kKpad / 16; ++kk)
            umma_f16(tmem_base + as * BN,
                     umma_desc_pack(hi, a0 + (kk >> 2) * (kFcABytes >> 4) + (kk & 3) * 2),
                     umma_desc_pack(hi, b0 + (kk >> 2) * ((BN * 128) >> 4) + (kk & 3) * 2), idesc,
                     kk != 0 ? 1u : 0u);
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
          const uint32_t a_addr = smem_u32(a_ring + s * kAB);
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
    // ==================== dZ tile producer (plain wgrad: TMA) ====================
    if (WGRAD && !POOL && elect_one()) {
      for (int i = 0; i < n_my; ++i) {
        const int s = i % S;
        mbar_wait(&a_empty[s], ((uint32_t)(i / S) & 1u) ^ 1u);
        mbar_expect_tx(&z_full[s], Cfg::kZBytes);
        tma_load_2d(&tmIO, &z_full[s], z_ring + s * Cfg::kZBytes, 0, tile_of(i) * 128);
      }
    }
  } else if (warp < 10) {
    const int quad = warp & 3;                       // TMEM lanes [32*quad, +32)
    const int half = (warp - 2) >> 2;                // which warp of the quadrant's pair
    if (!WGRAD && !POOL) {
      // ====================== epilogue: plain forward ======================
      // warp (quad, half) takes the tiles i with (i & 1) == half: accumulator stage `half`
      constexpr int NCH = BN / 32;
      uint8_t* my_stg = stg + (warp - 2) * 2 * Cfg::kStgBytes;
      const uint32_t tsrc = tmem_base + ((uint32_t)(quad * 32) << 16) + half * BN;
      int j = 0;
      for (int i = half; i < n_my; i += 2, ++j) {
        const int tile = tile_of(i);
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
    } else if (!WGRAD && POOL) {
      // ================= epilogue: forward + 2x2 max-pool =================
      // lane = accumulator row 4*xp + 2*dy + dx of its quadrant: 8 pool windows per warp
      const uint32_t tsrc = tmem_base + ((uint32_t)(quad * 32) << 16) + half * BN;
      const bool dx = (lane & 1) != 0, dy = (lane & 2) != 0;
      const int k = lane & 3;
      const int xp = quad * 8 + (lane >> 2);
      float bn_m[8], bn_s[8], bn_b[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const bool on = P.bn_mean != nullptr && 8 * k + e < P.cout;
        bn_m[e] = on ? __ldg(P.bn_mean + 8 * k + e) : 0.f;
        bn_s[e] = on ? rsqrtf(__ldg(P.bn_var + 8 * k + e) + P.bn_eps) : 1.f;
        bn_b[e] = on ? __ldg(P.bn_beta + 8 * k + e) : 0.f;
      }
      int j = 0;
      for (int i = half; i < n_my; i += 2, ++j) {
        int n, Y, tx;
        tile_pos(tile_of(i), n, Y, tx);
        mbar_wait(&tfull[half], (uint32_t)j & 1u);
        tc_fence_after();
        uint32_t r[32];
        tmem_ld_32x32(tsrc, r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[half]);
        uint32_t p[16];
#pragma unroll
        for (int w = 0; w < 16; ++w)
          p[w] = (P.flags & SEG_EPI_RELU)
                     ? pack_bf16x2_relu(__uint_as_float(r[2 * w]), __uint_as_float(r[2 * w + 1]))
                     : pack_bf16x2(__uint_as_float(r[2 * w]), __uint_as_float(r[2 * w + 1]));
        const int oy = 2 * Y + (dy ? 1 : 0);
        const int ox = 64 * tx + 2 * xp + (dx ? 1 : 0);
        if (ox < P.Wo && oy >= P.win_y0 && oy < P.win_y1 && ox >= P.win_x0 && ox < P.win_x1) {
          uint4* dst = reinterpret_cast<uint4*>(P.y + n * P.y_sn + oy * P.y_sh + ox * P.y_sw);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[q] = make_uint4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
        }
        // transpose through shared memory: lane k of a window gets chunk k of its 4 pixels
        const uint32_t sw = smem_u32(stg) + (uint32_t)(warp - 2) * kFcPoolStg;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          sts128(sw + (uint32_t)lane * 80u + q * 16,
                 make_uint4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]));
        __syncwarp();
        uint4 v[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          v[jj] = lds128(sw + (uint32_t)((lane & ~3) + jj) * 80u + (uint32_t)k * 16u);
        __syncwarp();
        if (P.bn_mean != nullptr) {
          // batch-norm (moving statistics) on this lane's 8 channels of the four candidates
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            uint32_t c4[4] = {v[jj].x, v[jj].y, v[jj].z, v[jj].w};
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const int ch = 8 * k + 2 * w;
              const float lo = (bf16_lo(c4[w]) - bn_m[2 * w]) * bn_s[2 * w] + bn_b[2 * w];
              const float hi = (bf16_hi(c4[w]) - bn_m[2 * w + 1]) * bn_s[2 * w + 1] + bn_b[2 * w + 1];
              (void)ch;
              c4[w] = pack_bf16x2(lo, hi);
            }
            v[jj] = make_uint4(c4[0], c4[1], c4[2], c4[3]);
          }
        }
        // maximum and slot of the first maximum in scan order (0,0),(0,1),(1,0),(1,1): a
        // later candidate wins only if strictly greater
        uint32_t best[4] = {v[0].x, v[0].y, v[0].z, v[0].w};
        uint32_t slot[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int jj = 1; jj < 4; ++jj) {
          const uint32_t c[4] = {v[jj].x, v[jj].y, v[jj].z, v[jj].w};
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const uint32_t g = bf16x2_gt_mask(c[w], best[w]);
            best[w] = (c[w] & g) | (best[w] & ~g);
            slot[w] = ((0x00010001u * (uint32_t)jj) & g) | (slot[w] & ~g);
          }
        }
        const int col = 32 * tx + xp;
        if (col < P.Wp) {
          const int64_t pix = ((int64_t)(n * P.Hp + Y) * P.Wp + col) * BN;
          *reinterpret_cast<uint4*>(P.pooled + pix + 8 * k) =
              make_uint4(best[0], best[1], best[2], best[3]);
          if (P.amax != nullptr) {
            uint32_t two[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) two[w] = (slot[w] & 0xffu) | ((slot[w] >> 8) & 0xff00u);
            *reinterpret_cast<uint2*>(P.amax + pix + 8 * k) =
                make_uint2(two[0] | (two[1] << 16), two[2] | (two[3] << 16));
          }
        }
      }
    } else if (WGRAD && POOL && half == 1) {
      // ============ dZ tile builders: max-pool backward into the operand ============
      // thread = (pool window xp, 16-byte chunk q of its 32 channels): the pooled gradient,
      // the window slots and the pooled activation (ReLU mask: the input at the argmax IS
      // the pooled value) arrive through per-thread cp.async groups kFcRawDepth tiles ahead;
      // the thread expands them to its four pixels' 16-byte chunks of the dZ tile (rows of
      // 64 bytes, SWIZZLE_64B, the MN-major B operand).
      const int u = (warp - 6) * 32 + lane;
      const int xp = u >> 2, q = u & 3;
      const uint32_t raw0 = smem_u32(raw_ring) + (uint32_t)u * 48u;
      auto fetch = [&](int i) {
        if (i < n_my) {
          int n, Y, tx;
          tile_pos(tile_of(i), n, Y, tx);
          const int col = 32 * tx + xp;
          const bool live = col < P.Wp;
          const int64_t idx = ((int64_t)(n * P.Hp + Y) * P.Wp + (live ? col : 0)) * BN + 8 * q;
          const uint32_t dst = raw0 + (uint32_t)(i % kFcRawDepth) * kFcRawBytes;
          cp_async_16(dst, P.dpool + idx, live ? 16u : 0u);
          cp_async_16(dst + 16, P.pooled + idx, live ? 16u : 0u);
          cp_async_8(dst + 32, P.amax + idx, live ? 8u : 0u);
        }
        cp_async_commit();
      };
#pragma unroll
      for (int i = 0; i < kFcRawDepth; ++i) fetch(i);
      for (int i = 0; i < n_my; ++i) {
        cp_async_wait<kFcRawDepth - 1>();
        const uint32_t src = raw0 + (uint32_t)(i % kFcRawDepth) * kFcRawBytes;
        const uint4 dp = lds128(src), pl = lds128(src + 16);
        uint2 am;
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(am.x), "=r"(am.y) : "r"(src + 32)
                     : "memory");
        int n, Y, tx;
        tile_pos(tile_of(i), n, Y, tx);
        const uint32_t dpw[4] = {dp.x, dp.y, dp.z, dp.w};
        const uint32_t plw[4] = {pl.x, pl.y, pl.z, pl.w};
        // ReLU mask from the pooled activation, per 16-bit half
        uint32_t pos[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) pos[w] = bf16x2_gt_mask(plw[w], 0u);
        const int s = i % S;
        mbar_wait(&a_empty[s], ((uint32_t)(i / S) & 1u) ^ 1u);
        const uint32_t zt = smem_u32(z_ring + s * Cfg::kZBytes);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          // slot == kk, per byte -> per 16-bit half masks
          uint32_t o[4];
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const uint32_t two = ((w < 2 ? am.x : am.y) >> (16 * (w & 1))) & 0xffffu;
            const uint32_t lo = (two & 0xffu) == (uint32_t)kk ? 0x0000ffffu : 0u;
            const uint32_t hi = (two >> 8) == (uint32_t)kk ? 0xffff0000u : 0u;
            o[w] = dpw[w] & (lo | hi) & pos[w];
          }
          const uint32_t m = (uint32_t)(4 * xp + kk);            // dZ tile row
          sts128(zt + m * rowB + (((uint32_t)q ^ ((m >> 1) & 3u)) << 4),
                 make_uint4(o[0], o[1], o[2], o[3]));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&z_full[s]);
        fetch(i + kFcRawDepth);                      // this thread's raw slot is free again
      }
    } else if (WGRAD && n_my > 0 && half == 0) {
      // partial sums of this CTA: accumulator rows 0..36 (k-slots + the constant-1 slot) ->
      // shared memory [37][BN] fp32 (the patch ring is idle once every MMA has completed);
      // the reduction below adds them to dW / db
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
    // here, so the builders run up to S tiles ahead of the tensor core.
    const int t = (warp - 10) * 32 + lane;
    const uint32_t x7 = (uint32_t)t & 7u;
    for (int i = 0; i < n_my; ++i) {
      int n, oy, ox;
      bool live;
      if (POOL) {
        int Y, tx;
        tile_pos(tile_of(i), n, Y, tx);
        oy = 2 * Y + ((t >> 1) & 1);
        ox = 64 * tx + 2 * (t >> 2) + (t & 1);
        live = ox < P.Wo;
      } else {
        const int m = tile_of(i) * 128 + t;
        live = m < P.M_total;
        const uint32_t mm = live ? (uint32_t)m : 0u;
        n = (int)fc_div(mm, P.div_b_mul, P.div_b_shr);
        const uint32_t rem = mm - (uint32_t)n * P.div_b;
        oy = (int)fc_div(rem, P.div_a_mul, P.div_a_shr);
        ox = (int)(rem - (uint32_t)oy * P.div_a);
      }
      const int iy0 = oy * ST - P.pad_t, ix0 = ox * ST - P.pad_l;
      const int s = i % S;
      mbar_wait(&a_empty[s], ((uint32_t)(i / S) & 1u) ^ 1u);
      const uint32_t row = smem_u32(a_ring + s * kAB) + (uint32_t)t * 128u;
#pragma unroll
      for (int r = 0; r < KH; ++r) {
        const int iy = iy0 + r;
        const bool yok = live && (unsigned)iy < (unsigned)P.H;
        const int iyc = yok ? iy : 0;
#pragma unroll
        for (int sx = 0; sx < KH; ++sx) {
          const int ix = ix0 + sx;
          const bool ok = yok && (unsigned)ix < (unsigned)P.W;
          const uint2* src = P.x4 + (n * P.x_sn + iyc * P.x_sh + (ok ? ix : 0));
          const int q = r * KH + sx;                 // 8-byte piece q of the pixel's K row
          cp_async_8(row + (uint32_t)(q >> 4) * kFcABytes +
                         (((((uint32_t)q & 15u) >> 1) ^ x7) << 4) + (uint32_t)(q & 1) * 8u,
                     src, ok ? 8u : 0u);
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
