// Memory-bound kernels of the segmentation hot path: max-pool with argmax,
// bilinear upsampling / resize, batch-norm, dropout, softmax cross-entropy,
// inference head, MC statistics, Adam, layout helpers.  All are HBM-bound:
// 16-byte vector accesses on the channel axis where the channel count allows,
// grid-stride loops sized from the SM count, warp-shuffle reductions.
#include "common.cuh"
#include "ptx.cuh"

namespace segb {

static int grid_for(int64_t total, int block) {
  int64_t g = ceil_div64(total, block);
  const int64_t cap = (int64_t)num_sms() * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}
#define GRID_STRIDE(i, total)                                                   \
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (total); \
       i += (int64_t)gridDim.x * blockDim.x)

// ---------------------------------------------------------------- max-pool
template <int VEC>
__global__ void maxpool_fwd_kernel(seg_view x, int k, int s, seg_view y, uint8_t* argmax) {
  pdl_trigger();
  pdl_wait();
  const int cv = y.c / VEC;
  const int64_t total = (int64_t)y.n * y.h * y.w * cv;
  GRID_STRIDE(idx, total) {
    const int c0 = (idx % cv) * VEC;
    int64_t m = idx / cv;
    const int q = m % y.w;
    m /= y.w;
    const int p = m % y.h;
    const int n = m / y.h;
    float best[VEC];
    uint8_t slot[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { best[j] = -INFINITY; slot[j] = 0; }
    for (int dy = 0; dy < k; ++dy)
      for (int dx = 0; dx < k; ++dx) {
        const bf16* xp = view_at(x, n, p * s + dy, q * s + dx) + c0;
        float v[VEC];
        if (VEC == 8) {
          const uint4 u = *reinterpret_cast<const uint4*>(xp);
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) { v[2 * j] = bf16_lo(w[j]); v[2 * j + 1] = bf16_hi(w[j]); }
        } else {
#pragma unroll
          for (int j = 0; j < VEC; ++j) v[j] = __bfloat162float(xp[j]);
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          if (v[j] > best[j] || (dy == 0 && dx == 0)) {   // strict >: first max wins
            best[j] = v[j];
            slot[j] = (uint8_t)(dy * k + dx);
          }
      }
    bf16* yp = view_at_mut(y, n, p, q) + c0;
    uint8_t* ap = argmax + (((int64_t)n * y.h + p) * y.w + q) * y.c + c0;
    if (VEC == 8) {
      uint4 o;
      o.x = pack_bf16x2(best[0], best[1]);
      o.y = pack_bf16x2(best[2], best[3]);
      o.z = pack_bf16x2(best[4], best[5]);
      o.w = pack_bf16x2(best[6], best[7]);
      *reinterpret_cast<uint4*>(yp) = o;
      uint2 a;
      a.x = slot[0] | (slot[1] << 8) | (slot[2] << 16) | ((uint32_t)slot[3] << 24);
      a.y = slot[4] | (slot[5] << 8) | (slot[6] << 16) | ((uint32_t)slot[7] << 24);
      *reinterpret_cast<uint2*>(ap) = a;
    } else {
#pragma unroll
      for (int j = 0; j < VEC; ++j) { yp[j] = __float2bfloat16(best[j]); ap[j] = slot[j]; }
    }
  }
}

// Vector form of the pool backward (8 channels per thread, k == s): one thread per pool
// window reads its dy / argmax once (16 + 8 bytes) and produces the k*k input-gradient
// pixels of the window, each with one 16-byte mask load, optional 16-byte add load and one
// 16-byte store.  The window grid is extended by one row / column so that input pixels
// beyond the last full window (VALID pooling drops them) still get add/mask/zero.
__global__ void maxpool_bwd_cell8_kernel(seg_view dy, seg_view dy2, const uint8_t* argmax, int k,
                                         seg_view add, int add_y0, int add_x0, seg_view mask,
                                         seg_view pooled, seg_view dx) {
  pdl_trigger();
  pdl_wait();
  const int cv = dx.c / 8;
  const int ch = (dx.h + k - 1) / k, cw = (dx.w + k - 1) / k;
  const int64_t total = (int64_t)dx.n * ch * cw * cv;
  GRID_STRIDE(idx, total) {
    const int c0 = (idx % cv) * 8;
    int64_t m = idx / cv;
    const int q = m % cw;
    m /= cw;
    const int p = m % ch;
    const int n = m / ch;
    float gsel[8];
    uint32_t sl[8];
    const bool in_pool = p < dy.h && q < dy.w;
    if (in_pool) {
      const uint4 u = *reinterpret_cast<const uint4*>(view_at(dy, n, p, q) + c0);
      const uint2 a = *reinterpret_cast<const uint2*>(
          argmax + (((int64_t)n * dy.h + p) * dy.w + q) * dy.c + c0);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { gsel[2 * j] = bf16_lo(w[j]); gsel[2 * j + 1] = bf16_hi(w[j]); }
      if (dy2.ptr) {
        const uint4 u2 = *reinterpret_cast<const uint4*>(view_at(dy2, n, p, q) + c0);
        const uint32_t w2[4] = {u2.x, u2.y, u2.z, u2.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { gsel[2 * j] += bf16_lo(w2[j]); gsel[2 * j + 1] += bf16_hi(w2[j]); }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { sl[j] = (a.x >> (8 * j)) & 0xffu; sl[4 + j] = (a.y >> (8 * j)) & 0xffu; }
      if (pooled.ptr) {
        // the ReLU mask of the routed gradient: x at the argmax IS the pooled value
        const uint4 pv = *reinterpret_cast<const uint4*>(view_at(pooled, n, p, q) + c0);
        const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (!(bf16_lo(pw[j]) > 0.f)) gsel[2 * j] = 0.f;
          if (!(bf16_hi(pw[j]) > 0.f)) gsel[2 * j + 1] = 0.f;
        }
      }
    }
    for (int wy = 0; wy < k; ++wy) {
      const int yy = p * k + wy;
      if (yy >= dx.h) break;
      for (int wx = 0; wx < k; ++wx) {
        const int xx = q * k + wx;
        if (xx >= dx.w) break;
        const uint32_t me = (uint32_t)(wy * k + wx);
        float g[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = (in_pool && sl[j] == me) ? gsel[j] : 0.f;
        bool added = false;
        if (add.ptr) {
          const int ay = yy - add_y0, ax = xx - add_x0;
          if (ay >= 0 && ay < add.h && ax >= 0 && ax < add.w) {
            const uint4 u = *reinterpret_cast<const uint4*>(view_at(add, n, ay, ax) + c0);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { g[2 * j] += bf16_lo(w[j]); g[2 * j + 1] += bf16_hi(w[j]); }
            added = true;
          }
        }
        // with `pooled` the routed part is already masked: x is read only under `add`
        if (mask.ptr && (added || !pooled.ptr)) {
          const uint4 u = *reinterpret_cast<const uint4*>(view_at(mask, n, yy, xx) + c0);
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (!(bf16_lo(w[j]) > 0.f)) g[2 * j] = 0.f;
            if (!(bf16_hi(w[j]) > 0.f)) g[2 * j + 1] = 0.f;
          }
        }
        uint4 o;
        o.x = pack_bf16x2(g[0], g[1]);
        o.y = pack_bf16x2(g[2], g[3]);
        o.z = pack_bf16x2(g[4], g[5]);
        o.w = pack_bf16x2(g[6], g[7]);
        *reinterpret_cast<uint4*>(view_at_mut(dx, n, yy, xx) + c0) = o;
      }
    }
  }
}

// Row-mapped forms of the two pool kernels (8 channels per thread, k == s).  blockIdx.y is
// one (image, pool row); the threads of a block walk that row with 32-bit indices, so the
// only divisions left are by the channel-vector count and k (compile-time for k = 2, 3).
// Backward: a thread owns ONE input column and the k input rows of the window row, so a
// warp's 16-byte stores are contiguous (512 bytes per row); the k threads of a window read
// the same dy / argmax / pooled vectors (L1 hits).
static int g_pool_rows = 1;     // seg_set_option key 11
void pool_set_rows(int on) { g_pool_rows = on != 0; }
static int g_tail_mma = 1;      // seg_set_option key 18
void tail_set_mma(int on) { g_tail_mma = on != 0; }

// per-halfword mask (0xffff / 0): bf16 pair a > b (ordered compare, false on NaN)
__device__ __forceinline__ uint32_t bf16x2_gt_mask(uint32_t a, uint32_t b) {
  __nv_bfloat162 va, vb;
  memcpy(&va, &a, 4);
  memcpy(&vb, &b, 4);
  return __hgt2_mask(va, vb);                                   // one HSET2.BF16
}

template <int K>
__global__ void __launch_bounds__(256)
maxpool_fwd_row8_kernel(seg_view x, int k_rt, seg_view y, uint8_t* argmax) {
  pdl_trigger();
  pdl_wait();
  const int k = K ? K : k_rt;
  const int cv = y.c >> 3;
  const int rowlen = y.w * cv;
  const int n = blockIdx.y / y.h, p = blockIdx.y - n * y.h;
  const bf16* xrow = view_at(x, n, p * k, 0);
  bf16* yrow = view_at_mut(y, n, p, 0);
  uint8_t* arow = argmax + ((int64_t)n * y.h + p) * y.w * y.c;
  const int x_sw = (int)x.sw, y_sw = (int)y.sw, yc = y.c;
  // (column, channel-vector) of the thread's items advance by a fixed step: one division
  // per thread instead of one per item
  const int step = gridDim.x * blockDim.x;
  const int step_q = step / cv, step_c = step - step_q * cv;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int q = i / cv, c8 = i - q * cv;
  for (; i < rowlen; i += step, q += step_q, c8 += step_c) {
    if (c8 >= cv) { c8 -= cv; ++q; }
    const int c0 = c8 * 8;
    // running max and its window slot per bf16 pair, selected with halfword masks
    uint32_t best[4] = {0u, 0u, 0u, 0u}, slot[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int dy = 0; dy < (K ? K : 8); ++dy) {
      if (dy >= k) break;
#pragma unroll
      for (int dx = 0; dx < (K ? K : 8); ++dx) {
        if (dx >= k) break;
        const uint4 u = *reinterpret_cast<const uint4*>(xrow + dy * x.sh + (q * k + dx) * x_sw + c0);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
        if (dy == 0 && dx == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) best[j] = w[j];
        } else {
          const uint32_t sc = (uint32_t)(dy * k + dx) * 0x00010001u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t m = bf16x2_gt_mask(w[j], best[j]);     // strict >: first max wins
            best[j] = (w[j] & m) | (best[j] & ~m);
            slot[j] = (sc & m) | (slot[j] & ~m);
          }
        }
      }
    }
    *reinterpret_cast<uint4*>(yrow + q * y_sw + c0) = make_uint4(best[0], best[1], best[2], best[3]);
    uint2 a;
    a.x = __byte_perm(slot[0], slot[1], 0x6420);
    a.y = __byte_perm(slot[2], slot[3], 0x6420);
    *reinterpret_cast<uint2*>(arow + q * yc + c0) = a;
  }
}

// slim.batch_norm at inference (moving statistics, centre only) on one value: the single
// expression every kernel that applies it shares, so a fused and an unfused evaluation round
// the same way
__device__ __forceinline__ float bn_infer_value(float v, float mean, float rstd, float beta) {
  return (v - mean) * rstd + beta;
}

// max-pool (k == stride) followed by the inference batch-norm of the layer BEFORE it:
// y = bn(maxpool(x)).  The affine map has a positive slope (rsqrt(var + eps)) and every
// rounding on the way is monotonic, so this equals maxpool(bn(x)) - the order the model
// states (models/deconvolution.py:126-138: conv -> bn -> pool) - bit for bit, while the
// normalised full-resolution tensor is neither written nor read.  No argmax (inference).
template <int K>
__global__ void __launch_bounds__(256)
maxpool_bn_infer_row8_kernel(seg_view x, int k_rt, const float* __restrict__ mean,
                             const float* __restrict__ var, float eps,
                             const float* __restrict__ beta, int bn_c, seg_view y) {
  pdl_trigger();
  pdl_wait();
  const int k = K ? K : k_rt;
  const int cv = y.c >> 3;
  const int rowlen = y.w * cv;
  const int n = blockIdx.y / y.h, p = blockIdx.y - n * y.h;
  const bf16* xrow = view_at(x, n, p * k, 0);
  bf16* yrow = view_at_mut(y, n, p, 0);
  const int x_sw = (int)x.sw, y_sw = (int)y.sw;
  const int step = gridDim.x * blockDim.x;
  const int step_q = step / cv, step_c = step - step_q * cv;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int q = i / cv, c8 = i - q * cv;
  for (; i < rowlen; i += step, q += step_q, c8 += step_c) {
    if (c8 >= cv) { c8 -= cv; ++q; }
    const int c0 = c8 * 8;
    uint32_t best[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int dy = 0; dy < (K ? K : 8); ++dy) {
      if (dy >= k) break;
#pragma unroll
      for (int dx = 0; dx < (K ? K : 8); ++dx) {
        if (dx >= k) break;
        const uint4 u = *reinterpret_cast<const uint4*>(xrow + dy * x.sh + (q * k + dx) * x_sw + c0);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
        if (dy == 0 && dx == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) best[j] = w[j];
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t m = bf16x2_gt_mask(w[j], best[j]);
            best[j] = (w[j] & m) | (best[j] & ~m);
          }
        }
      }
    }
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float r[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = c0 + 2 * j + e;
        const float v = e ? bf16_hi(best[j]) : bf16_lo(best[j]);
        r[e] = c < bn_c ? bn_infer_value(v, __ldg(mean + c), rsqrtf(__ldg(var + c) + eps),
                                         __ldg(beta + c))
                        : v;
      }
      o[j] = pack_bf16x2(r[0], r[1]);
    }
    *reinterpret_cast<uint4*>(yrow + q * y_sw + c0) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// per-halfword mask (0xffff / 0) of the bf16 pairs in `w` that are > 0 (NaN: not > 0)
__device__ __forceinline__ uint32_t bf16x2_gt0_mask(uint32_t w) {
  __nv_bfloat162 v;
  memcpy(&v, &w, 4);
  return __hgt2_mask(v, __floats2bfloat162_rn(0.f, 0.f));       // one HSET2.BF16
}

template <int K, bool TWO>
__global__ void __launch_bounds__(256)
maxpool_bwd_row8_kernel(seg_view dy, seg_view dy2, const uint8_t* argmax, int k_rt, seg_view add,
                        int add_y0, int add_x0, seg_view mask, seg_view pooled, seg_view dx) {
  pdl_trigger();
  pdl_wait();
  const int k = K ? K : k_rt;
  const int cv = dx.c >> 3;
  const int ch = (dx.h + k - 1) / k;
  const int rowlen = dx.w * cv;
  const int n = blockIdx.y / ch, p = blockIdx.y - n * ch;
  const bool p_ok = p < dy.h;
  // row bases (64-bit once per block); inside the row everything is 32-bit
  const bf16* dy_row = p_ok ? view_at(dy, n, p, 0) : nullptr;
  const bf16* dy2_row = (p_ok && dy2.ptr) ? view_at(dy2, n, p, 0) : nullptr;
  const bf16* po_row = (p_ok && pooled.ptr) ? view_at(pooled, n, p, 0) : nullptr;
  const uint8_t* am_row = p_ok ? argmax + ((int64_t)n * dy.h + p) * dy.w * dy.c : nullptr;
  bf16* dx_row = view_at_mut(dx, n, p * k, 0);
  const bf16* mk_row = mask.ptr ? view_at(mask, n, p * k, 0) : nullptr;
  const int dy_sw = (int)dy.sw, dy2_sw = (int)dy2.sw, po_sw = (int)pooled.sw, dyc = dy.c;
  const int dx_sw = (int)dx.sw, mk_sw = (int)mask.sw;
  const int step = gridDim.x * blockDim.x;
  const int step_x = step / cv, step_c = step - step_x * cv;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int xx = i / cv, c8 = i - xx * cv;
  for (; i < rowlen; i += step, xx += step_x, c8 += step_c) {
    if (c8 >= cv) { c8 -= cv; ++xx; }
    const int c0 = c8 * 8;
    const int q = xx / k;
    const uint32_t wx = (uint32_t)(xx - q * k);
    uint32_t w[4] = {0u, 0u, 0u, 0u};      // routed gradient of the window, bf16 pairs
    uint2 a = make_uint2(0u, 0u);          // outside the pool grid w == 0, any slot will do
    float gsel[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) gsel[j] = 0.f;
    const bool in_pool = p_ok && q < dy.w;
    constexpr bool two = TWO;               // a second gradient (dy2) is summed in fp32
    if (in_pool) {
      const uint4 u = *reinterpret_cast<const uint4*>(dy_row + q * dy_sw + c0);
      a = *reinterpret_cast<const uint2*>(am_row + q * dyc + c0);
      w[0] = u.x; w[1] = u.y; w[2] = u.z; w[3] = u.w;
      if (two) {
        const uint4 u2 = *reinterpret_cast<const uint4*>(dy2_row + q * dy2_sw + c0);
        const uint32_t w2[4] = {u2.x, u2.y, u2.z, u2.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          gsel[2 * j] = bf16_lo(w[j]) + bf16_lo(w2[j]);
          gsel[2 * j + 1] = bf16_hi(w[j]) + bf16_hi(w2[j]);
        }
      }
      if (po_row) {
        // the ReLU mask of the routed gradient: x at the argmax IS the pooled value
        const uint4 pv = *reinterpret_cast<const uint4*>(po_row + q * po_sw + c0);
        const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t m = bf16x2_gt0_mask(pw[j]);
          w[j] &= m;
          if (two) {
            if (!(m & 0xffffu)) gsel[2 * j] = 0.f;
            if (!(m >> 16)) gsel[2 * j + 1] = 0.f;
          }
        }
      }
    }
#pragma unroll
    for (int wy = 0; wy < (K ? K : 8); ++wy) {
      const int yy = p * k + wy;
      if (wy >= k || yy >= dx.h) break;
      const uint32_t me = (uint32_t)(wy * k) + wx;
      // sel[j]: 0xffff per channel of pair j whose argmax slot is this pixel
      uint32_t sel[4];
      if (K == 2) {
        // slots 0..3 used as PRMT byte selectors into a word whose byte `me` is 0xff;
        // slot * 0x11 doubles each selector nibble (one mask byte per bf16 half)
        const uint32_t cm = 0xffu << (8 * me);
        const uint32_t t0 = a.x * 0x11u, t1 = a.y * 0x11u;
        sel[0] = __byte_perm(cm, 0u, t0);
        sel[1] = __byte_perm(cm, 0u, t0 >> 16);
        sel[2] = __byte_perm(cm, 0u, t1);
        sel[3] = __byte_perm(cm, 0u, t1 >> 16);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t word = j < 2 ? a.x : a.y;
          const uint32_t s0 = (word >> (16 * (j & 1))) & 0xffu, s1 = (word >> (16 * (j & 1) + 8)) & 0xffu;
          sel[j] = (s0 == me ? 0xffffu : 0u) | (s1 == me ? 0xffff0000u : 0u);
        }
      }
      bool added = false;
      const int ay = yy - add_y0, ax = xx - add_x0;
      if (add.ptr && ay >= 0 && ay < add.h && ax >= 0 && ax < add.w) added = true;
      uint4 o;
      if (!added && !two) {
        // integer path: select, then (without `pooled`) the ReLU mask of the pool input
        uint32_t r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) r[j] = w[j] & sel[j];
        if (mk_row && !po_row) {
          const uint4 u = *reinterpret_cast<const uint4*>(mk_row + wy * mask.sh + xx * mk_sw + c0);
          r[0] &= bf16x2_gt0_mask(u.x); r[1] &= bf16x2_gt0_mask(u.y);
          r[2] &= bf16x2_gt0_mask(u.z); r[3] &= bf16x2_gt0_mask(u.w);
        }
        o = make_uint4(r[0], r[1], r[2], r[3]);
      } else {
        float g[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float lo = two ? gsel[2 * j] : bf16_lo(w[j]);
          const float hi = two ? gsel[2 * j + 1] : bf16_hi(w[j]);
          g[2 * j] = (sel[j] & 0xffffu) ? lo : 0.f;
          g[2 * j + 1] = (sel[j] >> 16) ? hi : 0.f;
        }
        if (added) {
          const uint4 u = *reinterpret_cast<const uint4*>(view_at(add, n, ay, ax) + c0);
          const uint32_t wa[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) { g[2 * j] += bf16_lo(wa[j]); g[2 * j + 1] += bf16_hi(wa[j]); }
        }
        // with `pooled` the routed part is already masked: x is read only under `add`
        if (mk_row && (added || !po_row)) {
          const uint4 u = *reinterpret_cast<const uint4*>(mk_row + wy * mask.sh + xx * mk_sw + c0);
          const uint32_t wm[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (!(bf16_lo(wm[j]) > 0.f)) g[2 * j] = 0.f;
            if (!(bf16_hi(wm[j]) > 0.f)) g[2 * j + 1] = 0.f;
          }
        }
        o.x = pack_bf16x2(g[0], g[1]);
        o.y = pack_bf16x2(g[2], g[3]);
        o.z = pack_bf16x2(g[4], g[5]);
        o.w = pack_bf16x2(g[6], g[7]);
      }
      *reinterpret_cast<uint4*>(dx_row + wy * dx.sh + xx * dx_sw + c0) = o;
    }
  }
}

// dx = relu_mask(route(dy, argmax) + add).  Non-overlapping windows (k == s).
template <int VEC>
__global__ void maxpool_bwd_kernel(seg_view dy, seg_view dy2, const uint8_t* argmax, int k, int s,
                                   seg_view add, int add_y0, int add_x0, seg_view mask,
                                   seg_view dx) {
  pdl_trigger();
  pdl_wait();
  const int cv = dx.c / VEC;
  const int64_t total = (int64_t)dx.n * dx.h * dx.w * cv;
  GRID_STRIDE(idx, total) {
    const int c0 = (idx % cv) * VEC;
    int64_t m = idx / cv;
    const int xx = m % dx.w;
    m /= dx.w;
    const int yy = m % dx.h;
    const int n = m / dx.h;
    float g[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) g[j] = 0.f;
    const int p = yy / s, q = xx / s;
    if (p < dy.h && q < dy.w) {
      const uint8_t me = (uint8_t)((yy - p * s) * k + (xx - q * s));
      const bf16* gp = view_at(dy, n, p, q) + c0;
      const uint8_t* ap = argmax + (((int64_t)n * dy.h + p) * dy.w + q) * dy.c + c0;
#pragma unroll
      for (int j = 0; j < VEC; ++j)
        if (ap[j] == me) g[j] = __bfloat162float(gp[j]);
      if (dy2.ptr) {
        const bf16* gp2 = view_at(dy2, n, p, q) + c0;
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          if (ap[j] == me) g[j] += __bfloat162float(gp2[j]);
      }
    }
    if (add.ptr) {
      const int ay = yy - add_y0, ax = xx - add_x0;
      if (ay >= 0 && ay < add.h && ax >= 0 && ax < add.w) {
        const bf16* p2 = view_at(add, n, ay, ax) + c0;
#pragma unroll
        for (int j = 0; j < VEC; ++j) g[j] += __bfloat162float(p2[j]);
      }
    }
    if (mask.ptr) {
      const bf16* mp = view_at(mask, n, yy, xx) + c0;
#pragma unroll
      for (int j = 0; j < VEC; ++j)
        if (!(__bfloat162float(mp[j]) > 0.f)) g[j] = 0.f;
    }
    bf16* op = view_at_mut(dx, n, yy, xx) + c0;
    if (VEC == 8) {
      uint4 o;
      o.x = pack_bf16x2(g[0], g[1]);
      o.y = pack_bf16x2(g[2], g[3]);
      o.z = pack_bf16x2(g[4], g[5]);
      o.w = pack_bf16x2(g[6], g[7]);
      *reinterpret_cast<uint4*>(op) = o;
    } else {
#pragma unroll
      for (int j = 0; j < VEC; ++j) op[j] = __float2bfloat16(g[j]);
    }
  }
}

__global__ void relu_grad_kernel(seg_view dy, seg_view y, seg_view dz) {
  const int64_t total = (int64_t)dz.n * dz.h * dz.w * dz.c;
  GRID_STRIDE(idx, total) {
    const int c = idx % dz.c;
    int64_t m = idx / dz.c;
    const int xx = m % dz.w;
    m /= dz.w;
    const int yy = m % dz.h;
    const int n = m / dz.h;
    const float g = __bfloat162float(view_at(dy, n, yy, xx)[c]);
    const float a = __bfloat162float(view_at(y, n, yy, xx)[c]);
    view_at_mut(dz, n, yy, xx)[c] = __float2bfloat16(a > 0.f ? g : 0.f);
  }
}

static bool vec8_ok(const seg_view& v) {
  return v.c % 8 == 0 && v.sw % 8 == 0 && v.sh % 8 == 0 && v.sn % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(v.ptr) % 16) == 0;
}

// ------------------------------------------------- FCN depthwise bilinear
// weight of tap a (0..k-1) of upsample_filt(k): 1 - |a - center| / F   (double)
__device__ __forceinline__ double bil_w(int a, int k) {
  const int F = (k + 1) / 2;
  const double center = (k & 1) ? (double)(F - 1) : (double)F - 0.5;
  return 1.0 - fabs((double)a - center) / (double)F;
}

// The k tap weights are tabulated once per block (in double, as the reference's numpy code
// computes them, utils/upsampling.py:13-24): evaluating bil_w per tap costs two double
// divisions and made these kernels 10-30x slower than their memory traffic.
constexpr int kBilMaxK = 64;

// One thread = one output pixel x G consecutive channels (G <= 8 divides C): the tap
// geometry and weights are shared by the G channels; 32-bit index arithmetic.
__global__ void bilinear_up_fwd_kernel(seg_view x, int f, seg_view add, seg_view y, int y_f32,
                                       int G) {
  pdl_trigger();
  const int k = 2 * f - (f & 1);
  const int before = (k - f) / 2;
  __shared__ double swt[kBilMaxK];
  for (int a = threadIdx.x; a < k; a += blockDim.x) swt[a] = bil_w(a, k);
  __syncthreads();
  pdl_wait();
  const uint32_t cg = (uint32_t)(y.c / G);
  const uint32_t total = (uint32_t)y.n * y.h * y.w * cg;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += gridDim.x * blockDim.x) {
    const int c0 = (int)(idx % cg) * G;
    uint32_t m = idx / cg;
    const int ox = m % y.w;
    m /= y.w;
    const int oy = m % y.h;
    const int n = m / y.h;
    const int ty = oy + before, tx = ox + before;
    float acc[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) acc[g] = 0.f;
    // contributing inputs i: 0 <= ty - i*f < k
    for (int i = ty / f; i >= 0 && ty - i * f < k; --i) {
      if (i >= x.h) continue;
      const double wy = swt[ty - i * f];
      for (int j = tx / f; j >= 0 && tx - j * f < k; --j) {
        if (j >= x.w) continue;
        const float wgt = (float)(wy * swt[tx - j * f]);
        const bf16* xp = view_at(x, n, i, j) + c0;
#pragma unroll
        for (int g = 0; g < 8; ++g)
          if (g < G) acc[g] += __bfloat162float(xp[g]) * wgt;
      }
    }
    if (add.ptr) {
      const bf16* ap = view_at(add, n, oy, ox) + c0;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        if (g < G) acc[g] += __bfloat162float(ap[g]);
    }
    const int64_t off = n * y.sn + oy * y.sh + ox * y.sw + c0;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (g < G) {
        if (y_f32)
          reinterpret_cast<float*>(y.ptr)[off + g] = acc[g];
        else
          reinterpret_cast<bf16*>(y.ptr)[off + g] = __float2bfloat16(acc[g]);
      }
    }
  }
}

__global__ void bilinear_up_bwd_kernel(seg_view dy, int dy_f32, int f, seg_view mask, seg_view dx) {
  pdl_trigger();
  const int k = 2 * f - (f & 1);
  const int before = (k - f) / 2;
  __shared__ double swt[kBilMaxK];
  for (int a = threadIdx.x; a < k; a += blockDim.x) swt[a] = bil_w(a, k);
  __syncthreads();
  pdl_wait();
  const int64_t total = (int64_t)dx.n * dx.h * dx.w * dx.c;
  GRID_STRIDE(idx, total) {
    const int c = idx % dx.c;
    int64_t m = idx / dx.c;
    const int j = m % dx.w;
    m /= dx.w;
    const int i = m % dx.h;
    const int n = m / dx.h;
    float acc = 0.f;
    for (int a = 0; a < k; ++a) {
      const int oy = i * f + a - before;
      if (oy < 0 || oy >= dy.h) continue;
      const double wy = swt[a];
      const int64_t row = n * dy.sn + oy * dy.sh + c;
      for (int b = 0; b < k; ++b) {
        const int ox = j * f + b - before;
        if (ox < 0 || ox >= dy.w) continue;
        const float wgt = (float)(wy * swt[b]);
        const int64_t off = row + ox * dy.sw;
        const float g = dy_f32 ? reinterpret_cast<const float*>(dy.ptr)[off]
                               : __bfloat162float(reinterpret_cast<const bf16*>(dy.ptr)[off]);
        acc += g * wgt;
      }
    }
    if (mask.ptr && !(__bfloat162float(view_at(mask, n, i, j)[c]) > 0.f)) acc = 0.f;
    view_at_mut(dx, n, i, j)[c] = __float2bfloat16(acc);
  }
}

// ------------------------------------------- tf.image.resize_bilinear legacy
__device__ __forceinline__ void legacy_src(int o, float scale, int n_in, int& lo, int& hi,
                                           float& frac) {
  const float src = (float)o * scale;
  lo = (int)floorf(src);
  hi = min((int)ceilf(src), n_in - 1);
  frac = src - (float)lo;
}

__global__ void resize_bilinear_fwd_kernel(seg_view x, seg_view y) {
  const float sy = (float)x.h / (float)y.h, sx = (float)x.w / (float)y.w;
  const int64_t total = (int64_t)y.n * y.h * y.w * y.c;
  GRID_STRIDE(idx, total) {
    const int c = idx % y.c;
    int64_t m = idx / y.c;
    const int ox = m % y.w;
    m /= y.w;
    const int oy = m % y.h;
    const int n = m / y.h;
    int y0, y1, x0, x1;
    float ly, lx;
    legacy_src(oy, sy, x.h, y0, y1, ly);
    legacy_src(ox, sx, x.w, x0, x1, lx);
    const float tl = __bfloat162float(view_at(x, n, y0, x0)[c]);
    const float tr = __bfloat162float(view_at(x, n, y0, x1)[c]);
    const float bl = __bfloat162float(view_at(x, n, y1, x0)[c]);
    const float br = __bfloat162float(view_at(x, n, y1, x1)[c]);
    const float top = tl + (tr - tl) * lx;
    const float bot = bl + (br - bl) * lx;
    view_at_mut(y, n, oy, ox)[c] = __float2bfloat16(top + (bot - top) * ly);
  }
}

// 8 channels per thread, 16-byte accesses (same arithmetic per element as above)
__global__ void resize_bilinear_fwd_vec8_kernel(seg_view x, seg_view y) {
  pdl_trigger();
  pdl_wait();
  const float sy = (float)x.h / (float)y.h, sx = (float)x.w / (float)y.w;
  const int cv = y.c / 8;
  const int64_t total = (int64_t)y.n * y.h * y.w * cv;
  GRID_STRIDE(idx, total) {
    const int c0 = (idx % cv) * 8;
    int64_t m = idx / cv;
    const int ox = m % y.w;
    m /= y.w;
    const int oy = m % y.h;
    const int n = m / y.h;
    int y0, y1, x0, x1;
    float ly, lx;
    legacy_src(oy, sy, x.h, y0, y1, ly);
    legacy_src(ox, sx, x.w, x0, x1, lx);
    const uint4 a = *reinterpret_cast<const uint4*>(view_at(x, n, y0, x0) + c0);
    const uint4 b = *reinterpret_cast<const uint4*>(view_at(x, n, y0, x1) + c0);
    const uint4 c = *reinterpret_cast<const uint4*>(view_at(x, n, y1, x0) + c0);
    const uint4 d = *reinterpret_cast<const uint4*>(view_at(x, n, y1, x1) + c0);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    const uint32_t cw[4] = {c.x, c.y, c.z, c.w}, dw[4] = {d.x, d.y, d.z, d.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float r[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float tl = e ? bf16_hi(aw[j]) : bf16_lo(aw[j]);
        const float tr = e ? bf16_hi(bw[j]) : bf16_lo(bw[j]);
        const float bl = e ? bf16_hi(cw[j]) : bf16_lo(cw[j]);
        const float br = e ? bf16_hi(dw[j]) : bf16_lo(dw[j]);
        const float top = tl + (tr - tl) * lx;
        const float bot = bl + (br - bl) * lx;
        r[e] = top + (bot - top) * ly;
      }
      o[j] = pack_bf16x2(r[0], r[1]);
    }
    *reinterpret_cast<uint4*>(view_at_mut(y, n, oy, ox) + c0) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// gather form of ResizeBilinearGrad: each input pixel sums the outputs that read it
__global__ void resize_bilinear_bwd_kernel(seg_view dy, seg_view dx) {
  const float sy = (float)dx.h / (float)dy.h, sx = (float)dx.w / (float)dy.w;
  const int64_t total = (int64_t)dx.n * dx.h * dx.w * dx.c;
  GRID_STRIDE(idx, total) {
    const int c = idx % dx.c;
    int64_t m = idx / dx.c;
    const int ix = m % dx.w;
    m /= dx.w;
    const int iy = m % dx.h;
    const int n = m / dx.h;
    const int oy_lo = max(0, (int)floorf((float)(iy - 1) / sy) - 1);
    const int oy_hi = min(dy.h - 1, (int)ceilf((float)(iy + 1) / sy) + 1);
    const int ox_lo = max(0, (int)floorf((float)(ix - 1) / sx) - 1);
    const int ox_hi = min(dy.w - 1, (int)ceilf((float)(ix + 1) / sx) + 1);
    float acc = 0.f;
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      int y0, y1;
      float ly;
      legacy_src(oy, sy, dx.h, y0, y1, ly);
      float wy = 0.f;
      if (y0 == iy) wy += 1.f - ly;
      if (y1 == iy) wy += ly;
      if (wy == 0.f) continue;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        int x0, x1;
        float lx;
        legacy_src(ox, sx, dx.w, x0, x1, lx);
        float wx = 0.f;
        if (x0 == ix) wx += 1.f - lx;
        if (x1 == ix) wx += lx;
        if (wx == 0.f) continue;
        acc += __bfloat162float(view_at(dy, n, oy, ox)[c]) * wy * wx;
      }
    }
    view_at_mut(dx, n, iy, ix)[c] = __float2bfloat16(acc);
  }
}

// -------------------------------------------------------------- batch-norm
// per-channel partial sums: thread t owns channel (t % C) when C <= blockDim.
__global__ void channel_reduce2_kernel(seg_view a, seg_view b, const float* mean,
                                       const float* rstd, int mode, float* out0, float* out1) {
  // mode 0: out0 += sum a, out1 += sum a^2
  // mode 1: out0 += sum a (dy), out1 += sum a * xhat(b)
  // mode 2: out0 += sum a           (bias grad)
  extern __shared__ float sh[];
  const int C = a.c;
  const int64_t pixels = (int64_t)a.n * a.h * a.w;
  const int lanes = blockDim.x / C > 0 ? blockDim.x / C : 1;   // pixel lanes per block
  const int c = threadIdx.x % C;
  const int pl = threadIdx.x / C;
  float s0 = 0.f, s1 = 0.f;
  if (pl < lanes && threadIdx.x < lanes * C) {
    for (int cc = c; cc < C; cc += (C > (int)blockDim.x ? blockDim.x : C)) {
      float t0 = 0.f, t1 = 0.f;
      for (int64_t m = (int64_t)blockIdx.x * lanes + pl; m < pixels;
           m += (int64_t)gridDim.x * lanes) {
        const int xx = m % a.w;
        const int64_t t = m / a.w;
        const int yy = t % a.h;
        const int n = t / a.h;
        const float av = __bfloat162float(view_at(a, n, yy, xx)[cc]);
        t0 += av;
        if (mode == 0)
          t1 += av * av;
        else if (mode == 1)
          t1 += av * ((__bfloat162float(view_at(b, n, yy, xx)[cc]) - mean[cc]) * rstd[cc]);
      }
      if (C > (int)blockDim.x) {   // one thread owns several channels: flush directly
        atomicAdd(out0 + cc, t0);
        if (mode != 2) atomicAdd(out1 + cc, t1);
      } else {
        s0 = t0;
        s1 = t1;
      }
    }
  }
  if (C <= (int)blockDim.x) {
    sh[threadIdx.x] = s0;
    sh[blockDim.x + threadIdx.x] = s1;
    __syncthreads();
    if (threadIdx.x < C) {
      float r0 = 0.f, r1 = 0.f;
      for (int l = 0; l < lanes; ++l) {
        r0 += sh[l * C + threadIdx.x];
        r1 += sh[blockDim.x + l * C + threadIdx.x];
      }
      atomicAdd(out0 + threadIdx.x, r0);
      if (mode != 2) atomicAdd(out1 + threadIdx.x, r1);
    }
  }
}

__global__ void bn_finalize_kernel(const float* sum, const float* sumsq, float inv_count, int C,
                                   float eps, float decay, float* mean, float* rstd,
                                   float* mmean, float* mvar) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mu = sum[c] * inv_count;
  const float var = fmaxf(sumsq[c] * inv_count - mu * mu, 0.f);
  mean[c] = mu;
  rstd[c] = rsqrtf(var + eps);
  if (mmean) mmean[c] = mmean[c] * decay + mu * (1.f - decay);
  if (mvar) mvar[c] = mvar[c] * decay + var * (1.f - decay);
}

__global__ void bn_apply_kernel(seg_view x, const float* mean, const float* rstd_or_var, float eps,
                                int is_var, const float* beta, seg_view y) {
  const int64_t total = (int64_t)x.n * x.h * x.w * x.c;
  GRID_STRIDE(idx, total) {
    const int c = idx % x.c;
    int64_t m = idx / x.c;
    const int xx = m % x.w;
    m /= x.w;
    const int yy = m % x.h;
    const int n = m / x.h;
    const float r = is_var ? rsqrtf(rstd_or_var[c] + eps) : rstd_or_var[c];
    const float v = (__bfloat162float(view_at(x, n, yy, xx)[c]) - mean[c]) * r + beta[c];
    view_at_mut(y, n, yy, xx)[c] = __float2bfloat16(v);
  }
}

// inference batch-norm folded to y = x * scale + shift for a conv epilogue; columns beyond
// the layer's c channels (the padded ones) map to zero
__global__ void bn_fold_kernel(const float* __restrict__ mean, const float* __restrict__ var,
                               float eps, const float* __restrict__ beta, int c, int c_pad,
                               float* __restrict__ scale, float* __restrict__ shift) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c_pad) return;
  float sc = 0.f, sh = 0.f;
  if (i < c) {
    sc = rsqrtf(var[i] + eps);
    sh = beta[i] - mean[i] * sc;
  }
  scale[i] = sc;
  shift[i] = sh;
}

// 8 channels per thread, 16-byte accesses; per-channel scale / shift from L1
__global__ void bn_apply_vec8_kernel(seg_view x, const float* __restrict__ mean,
                                     const float* __restrict__ rstd_or_var, float eps, int is_var,
                                     const float* __restrict__ beta, seg_view y) {
  pdl_trigger();
  pdl_wait();
  const int cv = x.c / 8;
  const int64_t total = (int64_t)x.n * x.h * x.w * cv;
  GRID_STRIDE(idx, total) {
    const int c0 = (idx % cv) * 8;
    int64_t m = idx / cv;
    const int xx = m % x.w;
    m /= x.w;
    const int yy = m % x.h;
    const int n = m / x.h;
    const uint4 u = *reinterpret_cast<const uint4*>(view_at(x, n, yy, xx) + c0);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float r[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = c0 + 2 * j + e;
        const float sc = is_var ? rsqrtf(__ldg(rstd_or_var + c) + eps) : __ldg(rstd_or_var + c);
        const float v = e ? bf16_hi(w[j]) : bf16_lo(w[j]);
        r[e] = bn_infer_value(v, __ldg(mean + c), sc, __ldg(beta + c));
      }
      o[j] = pack_bf16x2(r[0], r[1]);
    }
    *reinterpret_cast<uint4*>(view_at_mut(y, n, yy, xx) + c0) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void bn_bwd_apply_kernel(seg_view dy, seg_view x, const float* mean, const float* rstd,
                                    const float* dbeta, const float* dxhat, float inv_count,
                                    int relu_mask, seg_view dx) {
  const int64_t total = (int64_t)x.n * x.h * x.w * x.c;
  GRID_STRIDE(idx, total) {
    const int c = idx % x.c;
    int64_t m = idx / x.c;
    const int xx = m % x.w;
    m /= x.w;
    const int yy = m % x.h;
    const int n = m / x.h;
    const float xv = __bfloat162float(view_at(x, n, yy, xx)[c]);
    const float xh = (xv - mean[c]) * rstd[c];
    const float g = __bfloat162float(view_at(dy, n, yy, xx)[c]);
    float v = rstd[c] * (g - dbeta[c] * inv_count - xh * dxhat[c] * inv_count);
    if (relu_mask && !(xv > 0.f)) v = 0.f;
    view_at_mut(dx, n, yy, xx)[c] = __float2bfloat16(v);
  }
}

// ----------------------------------------------------------------- dropout
struct Philox4 { uint32_t v[4]; };
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  Philox4 r;
  r.v[0] = c0; r.v[1] = c1; r.v[2] = c2; r.v[3] = c3;
  return r;
}

// x, y dense with identical geometry; element index e is the dense NHWC index.
__global__ void dropout_kernel(const bf16* x, bf16* y, int64_t numel, uint32_t seed_lo,
                               uint32_t seed_hi, uint32_t stream_id, float keep, float inv_keep) {
  const int64_t nq = (numel + 3) / 4;
  GRID_STRIDE(qi, nq) {
    const Philox4 r = philox4x32_10((uint32_t)qi, (uint32_t)((uint64_t)qi >> 32), stream_id, 0u,
                                    seed_lo, seed_hi);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t e = qi * 4 + j;
      if (e >= numel) break;
      const float u = (float)(r.v[j] >> 8) * 5.9604644775390625e-8f;   // 2^-24
      const float v = __bfloat162float(x[e]);
      y[e] = __float2bfloat16(u < keep ? v * inv_keep : 0.f);
    }
  }
}

// Batched / graph-replayable form: image n of x uses Philox stream
//   stream_id0 + n * per_image_step + (step_dev ? *step_dev * step_mul : 0);
// per_image_step != 0: the element index restarts at every image (T MC passes of one tile
// in one launch, each pass its own stream); per_image_step == 0: one stream, dense index
// over the whole batch (identical to dropout_kernel).  One thread = 8 consecutive elements
// (two Philox counters, one 16-byte load / store); img_elems % 8 == 0.
__global__ void dropout_ex_kernel(const bf16* x, bf16* y, int64_t img_elems, int n_img,
                                  uint32_t seed_lo, uint32_t seed_hi, uint32_t stream_id0,
                                  uint32_t per_image_step, const uint32_t* step_dev,
                                  uint32_t step_mul, float keep, float inv_keep) {
  const uint32_t base = stream_id0 + (step_dev ? __ldg(step_dev) * step_mul : 0u);
  const int64_t per_img = img_elems / 8;
  const int64_t total = per_img * n_img;
  GRID_STRIDE(t, total) {
    int64_t q8 = t;                   // 8-element group index inside its Philox stream
    uint32_t sid = base;
    if (per_image_step) {
      const int64_t n = t / per_img;
      q8 = t - n * per_img;
      sid = base + (uint32_t)n * per_image_step;
    }
    const uint4 in = *reinterpret_cast<const uint4*>(x + t * 8);
    const uint32_t w[4] = {in.x, in.y, in.z, in.w};
    uint32_t o[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint64_t ctr = (uint64_t)q8 * 2 + h;
      const Philox4 r = philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), sid, 0u, seed_lo,
                                      seed_hi);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint32_t pair = w[h * 2 + j];
        const float u0 = (float)(r.v[2 * j] >> 8) * 5.9604644775390625e-8f;
        const float u1 = (float)(r.v[2 * j + 1] >> 8) * 5.9604644775390625e-8f;
        const float v0 = __uint_as_float(pair << 16), v1 = __uint_as_float(pair & 0xffff0000u);
        const __nv_bfloat162 pk = __floats2bfloat162_rn(u0 < keep ? v0 * inv_keep : 0.f,
                                                        u1 < keep ? v1 * inv_keep : 0.f);
        o[h * 2 + j] = *reinterpret_cast<const uint32_t*>(&pk);
      }
    }
    *reinterpret_cast<uint4*>(y + t * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------ loss / heads
// one thread per pixel, C <= 64 channels held in registers chunk-wise
__global__ void softmax_xent_kernel(seg_view logits, seg_view labels, float* loss_sum,
                                    seg_view dlogits, float inv_pixels) {
  pdl_trigger();
  pdl_wait();
  const int C = logits.c;
  const int64_t pixels = (int64_t)logits.n * logits.h * logits.w;
  // fast path: the pixel's logits live in registers (one read), exp evaluated once, the
  // bf16 gradient row leaves as 16-byte stores
  const bool reg_path = C <= 32 && (!dlogits.ptr || (dlogits.c <= 32 && dlogits.c % 8 == 0 &&
                                                     dlogits.sw % 8 == 0 && dlogits.sh % 8 == 0 &&
                                                     dlogits.sn % 8 == 0 &&
                                                     (reinterpret_cast<uintptr_t>(dlogits.ptr) & 15) == 0));
  float local = 0.f;
  GRID_STRIDE(m, pixels) {
    const int xx = m % logits.w;
    const int64_t t = m / logits.w;
    const int yy = t % logits.h;
    const int n = t / logits.h;
    const float* lp =
        reinterpret_cast<const float*>(logits.ptr) + n * logits.sn + yy * logits.sh + xx * logits.sw;
    const int lab = reinterpret_cast<const uint8_t*>(labels.ptr)[n * labels.sn + yy * labels.sh +
                                                                 xx * labels.sw];
    if (reg_path) {
      float v[32];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        v[c] = c < C ? lp[c] : -INFINITY;
        mx = fmaxf(mx, v[c]);
      }
      float se = 0.f, picked = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        if (c == lab) picked = v[c];
        v[c] = c < C ? expf(v[c] - mx) : 0.f;
        se += v[c];
      }
      local += mx + logf(se) - (lab < C ? picked : 0.f);
      if (dlogits.ptr) {
        bf16* dp = view_at_mut(dlogits, n, yy, xx);
        const float inv_se = 1.f / se;
#pragma unroll
        for (int c8 = 0; c8 < 32; c8 += 8) {
          if (c8 < dlogits.c) {
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = c8 + 2 * j;
              const float g0 = c < C ? (v[c] * inv_se - (c == lab ? 1.f : 0.f)) * inv_pixels : 0.f;
              const float g1 = c + 1 < C ? (v[c + 1] * inv_se - (c + 1 == lab ? 1.f : 0.f)) * inv_pixels : 0.f;
              o[j] = pack_bf16x2(g0, g1);
            }
            *reinterpret_cast<uint4*>(dp + c8) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      continue;
    }
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, lp[c]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(lp[c] - mx);
    const float lse = mx + logf(se);
    const float picked = lab < C ? lp[lab] : 0.f;
    local += lse - picked;
    if (dlogits.ptr) {
      bf16* dp = view_at_mut(dlogits, n, yy, xx);
      const float inv_se = 1.f / se;
      for (int c = 0; c < dlogits.c; ++c) {
        float g = 0.f;
        if (c < C) g = (expf(lp[c] - mx) * inv_se - (c == lab ? 1.f : 0.f)) * inv_pixels;
        dp[c] = __float2bfloat16(g);
      }
    }
  }
  local = warp_sum(local);
  __shared__ float sh[32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(loss_sum, v);
  }
}

// ------------------------------------ FCN-8s training head in one launch
// A train step of FCN-8s ends with  logits = bilinear x8 transposed conv (fixed weights) of
// the fused score map -> softmax cross-entropy -> dlogits -> the transposed conv's input
// gradient (reference models/fcn.py:207-220, models/basemodel.py:59-70,360).  Unfused that
// is three launches around two full-resolution class tensors (config 2, 16 x 512^2 x 21:
// 352 MB fp32 logits written and read, 268 MB bf16 dlogits written and read as 256 scattered
// 2-byte loads per input pixel and channel: 240 + 154 + 292 us of a 1.79 ms step).  Here a
// block owns 4 x 4 pixels of the score map: it evaluates logits, softmax and dlogits for the
// 40 x 40 outputs those pixels reach (the 32 x 32 it owns for the loss plus a 4-pixel ring
// recomputed by the neighbours' blocks too, x 1.56), keeps the bf16-rounded dlogits in shared
// memory, and gathers each input pixel's 16 x 16 window from there.  No full-resolution
// tensor is written (logits only when the caller asks for them), no atomics on dx.
// Every value is computed by the same expression, in the same order, as in
// bilinear_up_fwd_kernel, softmax_xent_kernel and bilinear_up_bwd_kernel (taps that fall
// outside the image enter as exact zeros instead of being skipped): dx is bit-identical,
// the loss differs by the order of its final sum.
constexpr int kUxF = 8, kUxK = 2 * kUxF, kUxTI = 4;
constexpr int kUxR = (kUxTI + 1) * kUxF;             // 40: output region side per block
constexpr int kUxXT = kUxTI + 2;                     // 6: input tile side incl. the ring
// dlogits in shared memory: one plane per channel PAIR, a 32-bit word (bf16 c, bf16 c+1) per
// region pixel; plane pitch an odd multiple of 16 bytes, so the 16-byte window loads of lanes
// that differ in the pair hit different banks
constexpr int kUxPitch = kUxR * kUxR + 4;            // words per plane: 1604 * 4 B = 401 * 16 B
struct UpXentArgs {
  seg_view x;            // [n, h, w, C] bf16 score map
  seg_view labels;       // [n, 8h, 8w, 1] uint8
  seg_view dx;           // [n, h, w, C] bf16
  seg_view mask;         // nullable: ReLU-grad source for dx
  float* loss_sum;
  float* logits;         // nullable fp32 [n, 8h, 8w, C] (dense)
  float inv_pixels;
  int C;
  int tiles_x, tiles_y;
};
static inline int ux_xpitch(int c) { return (c + 3) & ~3; }   // floats per input-tile pixel
static inline size_t ux_smem_bytes(int c) {
  return (size_t)((c + 1) / 2) * kUxPitch * 4 +
         (kUxK * kUxK + kUxXT * kUxXT * ux_xpitch(c)) * sizeof(float);
}

template <int CT>
__global__ void __launch_bounds__(256) upscore8_xent_kernel(const UpXentArgs A) {
  constexpr int CMAX = CT ? CT : 32;
  constexpr int C4MAX = (CMAX + 3) / 4;
  extern __shared__ __align__(16) uint8_t ux_smem[];
  const int C = CT ? CT : A.C;
  const int XP = (C + 3) & ~3;                                        // s_x pixel pitch (floats)
  const int NP = (C + 1) >> 1;                                        // channel pairs
  uint32_t* s_g = reinterpret_cast<uint32_t*>(ux_smem);               // [NP][kUxPitch]
  float* s_w2 = reinterpret_cast<float*>(ux_smem + (size_t)NP * kUxPitch * 4);   // [16][16]
  float* s_x = s_w2 + kUxK * kUxK;                                    // [6*6][XP]
  __shared__ float s_red[8];
  pdl_trigger();
  const int tid = threadIdx.x;
  {
    const int a = tid >> 4, b = tid & 15;
    s_w2[tid] = (float)(bil_w(a, kUxK) * bil_w(b, kUxK));
  }
  const int n = blockIdx.x / (A.tiles_x * A.tiles_y);
  const int trem = blockIdx.x - n * A.tiles_x * A.tiles_y;
  const int i0 = (trem / A.tiles_x) * kUxTI, j0 = (trem % A.tiles_x) * kUxTI;
  const int H = A.x.h * kUxF, W = A.x.w * kUxF;
  pdl_wait();
  for (int idx = tid; idx < kUxXT * kUxXT * XP; idx += 256) {
    const int c = idx % XP, pix = idx / XP;
    const int ii = i0 - 1 + pix / kUxXT, jj = j0 - 1 + pix % kUxXT;
    float v = 0.f;
    if (c < C && ii >= 0 && ii < A.x.h && jj >= 0 && jj < A.x.w)
      v = __bfloat162float(view_at(A.x, n, ii, jj)[c]);
    s_x[idx] = v;
  }
  __syncthreads();
  // ---- phase 1: logits -> softmax -> dlogits of the 40 x 40 region
  float local = 0.f;
  for (int p = tid; p < kUxR * kUxR; p += 256) {
    const int py = p / kUxR, px = p - py * kUxR;
    const int oy = kUxF * i0 - kUxF / 2 + py, ox = kUxF * j0 - kUxF / 2 + px;
    if (oy < 0 || oy >= H || ox < 0 || ox >= W) {
      for (int q = 0; q < NP; ++q) s_g[q * kUxPitch + p] = 0u;
      continue;
    }
    // taps in the order of bilinear_up_fwd_kernel: (i_hi, j_hi), (i_hi, j_lo), (i_lo, j_hi),
    // (i_lo, j_lo); filter tap a = ty - i * f with ty = oy + f/2
    const int a_hi = py % kUxF, b_hi = px % kUxF;
    const int li = py / kUxF, lj = px / kUxF;                 // tile row / column of i_lo, j_lo
    const float w_hh = s_w2[a_hi * kUxK + b_hi], w_hl = s_w2[a_hi * kUxK + b_hi + kUxF];
    const float w_lh = s_w2[(a_hi + kUxF) * kUxK + b_hi], w_ll = s_w2[(a_hi + kUxF) * kUxK + b_hi + kUxF];
    const float* x_hh = s_x + ((li + 1) * kUxXT + lj + 1) * XP;
    const float* x_hl = s_x + ((li + 1) * kUxXT + lj) * XP;
    const float* x_lh = s_x + (li * kUxXT + lj + 1) * XP;
    const float* x_ll = s_x + (li * kUxXT + lj) * XP;
    float v[C4MAX * 4];
    float mx = -INFINITY;
#pragma unroll
    for (int c4 = 0; c4 < C4MAX; ++c4) {
      if (4 * c4 < C) {
        const float4 hh = reinterpret_cast<const float4*>(x_hh)[c4];
        const float4 hl = reinterpret_cast<const float4*>(x_hl)[c4];
        const float4 lh = reinterpret_cast<const float4*>(x_lh)[c4];
        const float4 ll = reinterpret_cast<const float4*>(x_ll)[c4];
        const float xa[4] = {hh.x, hh.y, hh.z, hh.w}, xb[4] = {hl.x, hl.y, hl.z, hl.w};
        const float xc[4] = {lh.x, lh.y, lh.z, lh.w}, xd[4] = {ll.x, ll.y, ll.z, ll.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float acc = 0.f;
          acc += xa[e] * w_hh;
          acc += xb[e] * w_hl;
          acc += xc[e] * w_lh;
          acc += xd[e] * w_ll;
          const int c = 4 * c4 + e;
          v[c] = c < C ? acc : -INFINITY;
          mx = fmaxf(mx, v[c]);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[4 * c4 + e] = -INFINITY;
      }
    }
    const bool owned = py >= kUxF / 2 && py < kUxR - kUxF / 2 && px >= kUxF / 2 && px < kUxR - kUxF / 2;
    if (owned && A.logits) {
      float* lp = A.logits + (((int64_t)n * H + oy) * W + ox) * C;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) lp[c] = v[c];
    }
    const int lab = reinterpret_cast<const uint8_t*>(A.labels.ptr)[n * A.labels.sn + oy * A.labels.sh +
                                                                   ox * A.labels.sw];
    // the label's logit, re-evaluated from the tile (same expression, same bits) instead of a
    // 21-way register select
    float picked = 0.f;
    if (lab < C) {
      float acc = 0.f;
      acc += x_hh[lab] * w_hh;
      acc += x_hl[lab] * w_hl;
      acc += x_lh[lab] * w_lh;
      acc += x_ll[lab] * w_ll;
      picked = acc;
    }
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      v[c] = c < C ? expf(v[c] - mx) : 0.f;
      se += v[c];
    }
    if (owned) local += mx + logf(se) - picked;
    const float inv_se = 1.f / se;
#pragma unroll
    for (int q = 0; q < (CMAX + 1) / 2; ++q) {
      if (q < NP) {
        const float g0 = (v[2 * q] * inv_se - 0.f) * A.inv_pixels;
        const float g1 = 2 * q + 1 < C ? (v[2 * q + 1 < CMAX ? 2 * q + 1 : 0] * inv_se - 0.f) * A.inv_pixels : 0.f;
        s_g[q * kUxPitch + p] = pack_bf16x2(g0, g1);
      }
    }
    if (lab < C) {
      const float e_lab = expf(picked - mx);
      reinterpret_cast<bf16*>(s_g)[((lab >> 1) * kUxPitch + p) * 2 + (lab & 1)] =
          __float2bfloat16((e_lab * inv_se - 1.f) * A.inv_pixels);
    }
  }
  local = warp_sum(local);
  if ((tid & 31) == 0) s_red[tid >> 5] = local;
  __syncthreads();
  if (tid < 32) {
    float t = tid < 8 ? s_red[tid] : 0.f;
    t = warp_sum(t);
    if (tid == 0) atomicAdd(A.loss_sum, t);
  }
  // ---- phase 2: each (input pixel, channel pair) gathers its 16 x 16 window, rows then
  // columns, one accumulator per channel (the summation order of bilinear_up_bwd_kernel)
  for (int it = tid; it < kUxTI * kUxTI * NP; it += 256) {
    const int q = it % NP, cell = it / NP;
    const int ci = cell / kUxTI, cj = cell - ci * kUxTI;
    const int i = i0 + ci, j = j0 + cj;
    if (i >= A.x.h || j >= A.x.w) continue;
    const uint32_t base = smem_u32(s_g) + (uint32_t)((q * kUxPitch + (kUxF * ci) * kUxR + kUxF * cj) * 4);
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 2
    for (int a = 0; a < kUxK; ++a) {
      const float4* wr = reinterpret_cast<const float4*>(s_w2 + a * kUxK);
#pragma unroll
      for (int b4 = 0; b4 < 4; ++b4) {
        const uint4 g = lds128(base + (uint32_t)((a * kUxR + 4 * b4) * 4));
        const float4 w = wr[b4];
        acc0 += bf16_lo(g.x) * w.x; acc1 += bf16_hi(g.x) * w.x;
        acc0 += bf16_lo(g.y) * w.y; acc1 += bf16_hi(g.y) * w.y;
        acc0 += bf16_lo(g.z) * w.z; acc1 += bf16_hi(g.z) * w.z;
        acc0 += bf16_lo(g.w) * w.w; acc1 += bf16_hi(g.w) * w.w;
      }
    }
    const int c0 = 2 * q;
    if (A.mask.ptr) {
      const bf16* mp = view_at(A.mask, n, i, j);
      if (!(__bfloat162float(mp[c0]) > 0.f)) acc0 = 0.f;
      if (c0 + 1 < C && !(__bfloat162float(mp[c0 + 1]) > 0.f)) acc1 = 0.f;
    }
    bf16* dp = view_at_mut(A.dx, n, i, j);
    dp[c0] = __float2bfloat16(acc0);
    if (c0 + 1 < C) dp[c0 + 1] = __float2bfloat16(acc1);
  }
}

__global__ void sigmoid_argmax_kernel(seg_view logits, float* probs, float* labelmap) {
  const int C = logits.c;
  const int64_t pixels = (int64_t)logits.n * logits.h * logits.w;
  GRID_STRIDE(m, pixels) {
    const int xx = m % logits.w;
    const int64_t t = m / logits.w;
    const int yy = t % logits.h;
    const int n = t / logits.h;
    const float* lp =
        reinterpret_cast<const float*>(logits.ptr) + n * logits.sn + yy * logits.sh + xx * logits.sw;
    float best = -1.f;
    int bi = 0;
    for (int c = 0; c < C; ++c) {
      // fp32 sigmoid, saturating to exactly 1.0f for large logits like TF's
      const float sg = 1.f / (1.f + expf(-lp[c]));
      probs[m * C + c] = sg;
      if (sg > best) { best = sg; bi = c; }   // strict >: first index on ties
    }
    labelmap[m] = (float)bi;
  }
}

// ------------------------------------------------- class-map tail (inference), one launch
// DeconvModel ends with  resize_bilinear(H/2) -> 2x2/s2 transposed conv (+bias, ReLU) to
// n_classes -> batch-norm -> 3x3 SAME conv to n_classes -> sigmoid / argmax
// (reference models/deconvolution.py:163-174, head :79-82).  Unfused, the n_classes <= 4
// channels travel through HBM as 16-channel padded full-resolution tensors four times
// (3.5 of config 4's 6.4 ms).  Here a block owns 14 x 14 pixels of the half-resolution grid
// (28 x 28 outputs): phase A, one thread per half-resolution pixel of the 16 x 16 haloed
// tile: lerp of the C input channels (legacy resize arithmetic, bf16 rounding as the stored
// tensor would have), the four sub-pixels of the transposed conv, ReLU, batch-norm (moving
// statistics), each rounded to bf16 where the unfused path stores bf16, into a 32 x 32 x NC
// shared-memory tile (zero outside the image: the SAME padding of the 3x3 conv); phase B:
// 3x3 conv + sigmoid + first-index argmax per output pixel.  HBM traffic: the small input
// and the outputs.
constexpr int kTailT = 14;                 // half-resolution pixels per tile side
constexpr int kTailH = kTailT + 2;         // + halo
constexpr int kTailO = 2 * kTailH;         // bn tile side (outputs incl. halo ring)
struct TailArgs {
  seg_view x;                              // [n, hs, ws, C] bf16
  int rh, rw;                              // half-resolution grid (resize target)
  const bf16* w_up; int up_cop, up_cip;    // [2][2][cout_pad][cin_pad]
  const float* b_up;
  const float* bn_mean; const float* bn_var; float bn_eps; const float* bn_beta;
  const bf16* w_out; int out_cip, out_cop; // [3][3][cin_pad][cout_pad]
  const float* b_out;
  float* logits; float* probs; float* labelmap;
};

template <int NC, int C>
__global__ void __launch_bounds__(256) classmap_tail_kernel(const TailArgs A) {
  __shared__ __align__(16) float s_wup[4 * NC * C];   // [sub][co][ci]
  __shared__ float s_wout[9 * NC * NC];           // [tap][ci][co]
  __shared__ float s_aff[4 * NC];                 // b_up, bn scale, bn shift, b_out
  __shared__ bf16 s_t[kTailO * kTailO * NC];      // bn output tile
  pdl_trigger();
  pdl_wait();
  for (int i = threadIdx.x; i < 4 * NC * C; i += 256) {
    const int ci = i % C, co = (i / C) % NC, sub = i / (C * NC);
    s_wup[i] = __bfloat162float(A.w_up[((int64_t)sub * A.up_cop + co) * A.up_cip + ci]);
  }
  for (int i = threadIdx.x; i < 9 * NC * NC; i += 256) {
    const int co = i % NC, ci = (i / NC) % NC, tap = i / (NC * NC);
    s_wout[i] = __bfloat162float(A.w_out[((int64_t)tap * A.out_cip + ci) * A.out_cop + co]);
  }
  if (threadIdx.x < NC) {
    const int c = threadIdx.x;
    const float sc = rsqrtf(__ldg(A.bn_var + c) + A.bn_eps);
    s_aff[c] = __ldg(A.b_up + c);
    s_aff[NC + c] = sc;
    s_aff[2 * NC + c] = __ldg(A.bn_beta + c) - __ldg(A.bn_mean + c) * sc;   // informational
    s_aff[3 * NC + c] = __ldg(A.b_out + c);
  }
  __syncthreads();
  const int tiles_x = (A.rw + kTailT - 1) / kTailT, tiles_y = (A.rh + kTailT - 1) / kTailT;
  const int n = blockIdx.x / (tiles_x * tiles_y);
  const int trem = blockIdx.x - n * tiles_x * tiles_y;
  const int ty = trem / tiles_x, tx = trem - ty * tiles_x;
  const int ry0 = ty * kTailT - 1, rx0 = tx * kTailT - 1;   // haloed tile origin (half-res)
  // ---- phase A: thread = one half-resolution pixel of the 16 x 16 haloed tile
  {
    const int ly_ = threadIdx.x / kTailH, lx_ = threadIdx.x % kTailH;
    const int ry = ry0 + ly_, rx = rx0 + lx_;
    const bool inside = ry >= 0 && ry < A.rh && rx >= 0 && rx < A.rw;
    float v[C];
    if (inside) {
      const float sy = (float)A.x.h / (float)A.rh, sx = (float)A.x.w / (float)A.rw;
      int y0, y1, x0, x1;
      float fy, fx;
      legacy_src(ry, sy, A.x.h, y0, y1, fy);
      legacy_src(rx, sx, A.x.w, x0, x1, fx);
      const bf16* p00 = view_at(A.x, n, y0, x0);
      const bf16* p01 = view_at(A.x, n, y0, x1);
      const bf16* p10 = view_at(A.x, n, y1, x0);
      const bf16* p11 = view_at(A.x, n, y1, x1);
#pragma unroll
      for (int c0 = 0; c0 < C; c0 += 8) {
        const uint4 a = *reinterpret_cast<const uint4*>(p00 + c0);
        const uint4 b = *reinterpret_cast<const uint4*>(p01 + c0);
        const uint4 c = *reinterpret_cast<const uint4*>(p10 + c0);
        const uint4 d = *reinterpret_cast<const uint4*>(p11 + c0);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
        const uint32_t cw[4] = {c.x, c.y, c.z, c.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float tl = e ? bf16_hi(aw[j]) : bf16_lo(aw[j]);
            const float tr = e ? bf16_hi(bw[j]) : bf16_lo(bw[j]);
            const float bl = e ? bf16_hi(cw[j]) : bf16_lo(cw[j]);
            const float br = e ? bf16_hi(dw[j]) : bf16_lo(dw[j]);
            const float top = tl + (tr - tl) * fx;
            const float bot = bl + (br - bl) * fx;
            v[c0 + 2 * j + e] = __bfloat162float(__float2bfloat16(top + (bot - top) * fy));
          }
      }
    }
#pragma unroll
    for (int sub = 0; sub < 4; ++sub) {
      const int oyl = 2 * ly_ + (sub >> 1), oxl = 2 * lx_ + (sub & 1);
#pragma unroll
      for (int co = 0; co < NC; ++co) {
        float r = 0.f;
        if (inside) {
          // weights read as float4 (one shared-memory load per four products, all lanes the
          // same address); four partial sums shorten the dependency chain
          const float4* wp = reinterpret_cast<const float4*>(s_wup + (sub * NC + co) * C);
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int c4 = 0; c4 < C / 4; ++c4) {
            const float4 w4 = wp[c4];
            a0 += v[4 * c4] * w4.x;
            a1 += v[4 * c4 + 1] * w4.y;
            a2 += v[4 * c4 + 2] * w4.z;
            a3 += v[4 * c4 + 3] * w4.w;
          }
          float acc = (a0 + a1) + (a2 + a3);
          acc = fmaxf(acc + s_aff[co], 0.f);
          const float dq = __bfloat162float(__float2bfloat16(acc));    // stored deconv output
          r = (dq - __ldg(A.bn_mean + co)) * s_aff[NC + co] + __ldg(A.bn_beta + co);
        }
        s_t[(oyl * kTailO + oxl) * NC + co] = __float2bfloat16(r);
      }
    }
  }
  __syncthreads();
  // ---- phase B: 28 x 28 output pixels of the tile interior
  const int H = 2 * A.rh, W = 2 * A.rw;
  float wo[9 * NC * NC];
#pragma unroll
  for (int i = 0; i < 9 * NC * NC; ++i) wo[i] = s_wout[i];
  for (int i = threadIdx.x; i < 4 * kTailT * kTailT; i += 256) {
    const int oyl = i / (2 * kTailT), oxl = i - oyl * (2 * kTailT);
    const int oy = 2 * ty * kTailT + oyl, ox = 2 * tx * kTailT + oxl;
    if (oy >= H || ox >= W) continue;
    float acc[NC];
#pragma unroll
    for (int co = 0; co < NC; ++co) acc[co] = s_aff[3 * NC + co];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int sx = 0; sx < 3; ++sx) {
        const bf16* tp = s_t + ((oyl + 1 + r) * kTailO + (oxl + 1 + sx)) * NC;
#pragma unroll
        for (int ci = 0; ci < NC; ++ci) {
          const float xv = __bfloat162float(tp[ci]);
#pragma unroll
          for (int co = 0; co < NC; ++co) acc[co] += xv * wo[((r * 3 + sx) * NC + ci) * NC + co];
        }
      }
    const int64_t m = ((int64_t)n * H + oy) * W + ox;
    float best = -1.f;
    int bi = 0;
#pragma unroll
    for (int co = 0; co < NC; ++co) {
      if (A.logits) A.logits[m * NC + co] = acc[co];
      const float sg = 1.f / (1.f + expf(-acc[co]));
      A.probs[m * NC + co] = sg;
      if (sg > best) { best = sg; bi = co; }
    }
    A.labelmap[m] = (float)bi;
  }
}

// ---- the same tail with the transposed conv on the tensor cores
// Source-level profile of the kernel above (profiles/r02_ncu_tail.md): instruction-issue and
// LSU bound (L1 88 %, DRAM 1.0 TB/s), 63 % of its instructions in phase A - per
// half-resolution pixel 256 FFMA + 64 broadcast LDS.128 for the 32 -> 4*NC transposed conv -
// and 32 % in phase B (one output pixel per thread: 9*NC 16-bit LDS, a quarter of the
// threads idle in the last pass).  Here:
//   phase A: a thread still interpolates one half-resolution pixel, but leaves its 32 bf16
//     channels in shared memory (64-byte rows, 16-byte chunks XOR-swizzled so that both the
//     row-per-thread stores and ldmatrix are conflict-free); each warp then multiplies its own
//     32 pixels with the [32][4*NC] weight matrix by mma.sync.m16n8k16 (bf16 products are
//     exact, fp32 accumulation: the same value as the FFMA chain up to summation order), the
//     weight fragments living in registers for the whole kernel.  tcgen05 is the wrong tool
//     for a 32 x 8 product per pixel group: its M = 128 tile, TMEM allocation and
//     commit / tcgen05.ld round trip cost more than the 4 warp-level MMAs they would replace.
//   phase B: a thread owns the 2 x 2 outputs of one half-resolution pixel: 16 shared loads
//     feed four 3x3 windows, every store is 8 or 16 bytes wide.
// Same rounding points and the same 3x3 summation order as the kernel above.
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1,
                                            uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_m16n8k16(float (&d)[4], uint32_t a0, uint32_t a1,
                                                  uint32_t a2, uint32_t a3, uint32_t b0,
                                                  uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
      "{%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int NC>
__global__ void __launch_bounds__(256) classmap_tail_mma_kernel(const TailArgs A) {
  constexpr int C = 32;
  constexpr int NT = (4 * NC + 7) / 8;             // 8-column tiles of the [32][4*NC] product
  static_assert(kTailH * kTailH == 256, "one thread per haloed half-resolution pixel");
  __shared__ __align__(16) uint8_t s_v[256 * C * 2];   // interpolated pixels, bf16 [256][32]
  __shared__ float s_wout[9 * NC * NC];            // [tap][ci][co]
  __shared__ float s_aff[5 * NC];                  // b_up, bn rstd, bn mean, bn beta, b_out
  __shared__ __align__(16) bf16 s_t[kTailO * kTailO * NC];   // bn output tile
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  for (int i = tid; i < 9 * NC * NC; i += 256) {
    const int co = i % NC, ci = (i / NC) % NC, tap = i / (NC * NC);
    s_wout[i] = __bfloat162float(A.w_out[((int64_t)tap * A.out_cip + ci) * A.out_cop + co]);
  }
  if (tid < NC) {
    const int c = tid;
    s_aff[c] = __ldg(A.b_up + c);
    s_aff[NC + c] = rsqrtf(__ldg(A.bn_var + c) + A.bn_eps);
    s_aff[2 * NC + c] = __ldg(A.bn_mean + c);
    s_aff[3 * NC + c] = __ldg(A.bn_beta + c);
    s_aff[4 * NC + c] = __ldg(A.b_out + c);
  }
  // B fragments of the transposed conv: column n = sub * NC + co, rows = input channels
  uint32_t bw[NT][2][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int nn = nt * 8 + g;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      bw[nt][ks][0] = bw[nt][ks][1] = 0u;
      if (nn < 4 * NC) {
        const bf16* wp = A.w_up + ((int64_t)(nn / NC) * A.up_cop + (nn % NC)) * A.up_cip + ks * 16 + 2 * t;
        bw[nt][ks][0] = __ldg(reinterpret_cast<const unsigned int*>(wp));
        bw[nt][ks][1] = __ldg(reinterpret_cast<const unsigned int*>(wp + 8));
      }
    }
  }
  const int tiles_x = (A.rw + kTailT - 1) / kTailT, tiles_y = (A.rh + kTailT - 1) / kTailT;
  const int n = blockIdx.x / (tiles_x * tiles_y);
  const int trem = blockIdx.x - n * tiles_x * tiles_y;
  const int ty = trem / tiles_x, tx = trem - ty * tiles_x;
  const int ry0 = ty * kTailT - 1, rx0 = tx * kTailT - 1;   // haloed tile origin (half-res)
  // ---- phase A.1: thread = one half-resolution pixel of the 16 x 16 haloed tile
  {
    const int ly_ = tid / kTailH, lx_ = tid % kTailH;
    const int ry = ry0 + ly_, rx = rx0 + lx_;
    const bool inside = ry >= 0 && ry < A.rh && rx >= 0 && rx < A.rw;
    const uint32_t row = smem_u32(s_v) + (uint32_t)tid * (C * 2);
    const uint32_t swz = (uint32_t)((tid >> 1) & 3);
    if (inside) {
      const float sy = (float)A.x.h / (float)A.rh, sx = (float)A.x.w / (float)A.rw;
      int y0, y1, x0, x1;
      float fy, fx;
      legacy_src(ry, sy, A.x.h, y0, y1, fy);
      legacy_src(rx, sx, A.x.w, x0, x1, fx);
      const bf16* p00 = view_at(A.x, n, y0, x0);
      const bf16* p01 = view_at(A.x, n, y0, x1);
      const bf16* p10 = view_at(A.x, n, y1, x0);
      const bf16* p11 = view_at(A.x, n, y1, x1);
#pragma unroll
      for (int c0 = 0; c0 < C; c0 += 8) {
        const uint4 a = *reinterpret_cast<const uint4*>(p00 + c0);
        const uint4 b = *reinterpret_cast<const uint4*>(p01 + c0);
        const uint4 c = *reinterpret_cast<const uint4*>(p10 + c0);
        const uint4 d = *reinterpret_cast<const uint4*>(p11 + c0);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bq[4] = {b.x, b.y, b.z, b.w};
        const uint32_t cw[4] = {c.x, c.y, c.z, c.w}, dw[4] = {d.x, d.y, d.z, d.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float r[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float tl = e ? bf16_hi(aw[j]) : bf16_lo(aw[j]);
            const float tr = e ? bf16_hi(bq[j]) : bf16_lo(bq[j]);
            const float bl = e ? bf16_hi(cw[j]) : bf16_lo(cw[j]);
            const float br = e ? bf16_hi(dw[j]) : bf16_lo(dw[j]);
            const float top = tl + (tr - tl) * fx;
            const float bot = bl + (br - bl) * fx;
            r[e] = top + (bot - top) * fy;
          }
          o[j] = pack_bf16x2(r[0], r[1]);           // the bf16 the stored resize output holds
        }
        sts128(row + ((((uint32_t)c0 >> 3) ^ swz) << 4), make_uint4(o[0], o[1], o[2], o[3]));
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) sts128(row + ((uint32_t)q << 4), make_uint4(0u, 0u, 0u, 0u));
    }
  }
  __syncthreads();     // (also publishes s_aff / s_wout)
  // ---- phase A.2: each warp multiplies its 32 pixels with the weight matrix
  {
    float acc[2][NT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
    const uint32_t sv = smem_u32(s_v) + (uint32_t)warp * 32 * (C * 2);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int row = mt * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
        const int chunk = 2 * ks + (lane >> 4);
        uint32_t a0, a1, a2, a3;
        ldmatrix_x4(sv + (uint32_t)row * (C * 2) + (uint32_t)((chunk ^ ((row >> 1) & 3)) << 4), a0,
                    a1, a2, a3);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
          mma_bf16_m16n8k16(acc[mt][nt], a0, a1, a2, a3, bw[nt][ks][0], bw[nt][ks][1]);
      }
    // bias, ReLU, bf16, batch-norm, bf16 -> the four sub-pixels of each half-resolution pixel
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int p = warp * 32 + mt * 16 + h * 8 + g;
        const int ly_ = p / kTailH, lx_ = p % kTailH;
        const int ry = ry0 + ly_, rx = rx0 + lx_;
        const bool inside = ry >= 0 && ry < A.rh && rx >= 0 && rx < A.rw;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          float r[2];
          const int nn0 = nt * 8 + 2 * t;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int nn = nn0 + e;
            const int co = nn % NC;
            float v = 0.f;
            if (inside && nn < 4 * NC) {
              const float z = fmaxf(acc[mt][nt][2 * h + e] + s_aff[co], 0.f);
              const float dq = __bfloat162float(__float2bfloat16(z));   // stored deconv output
              v = bn_infer_value(dq, s_aff[2 * NC + co], s_aff[NC + co], s_aff[3 * NC + co]);
            }
            r[e] = v;
          }
          if (NC == 2) {
            // columns (2t, 2t+1) = sub-pixel t, classes 0 and 1: one 4-byte store
            const int oyl = 2 * ly_ + (t >> 1), oxl = 2 * lx_ + (t & 1);
            *reinterpret_cast<uint32_t*>(s_t + (oyl * kTailO + oxl) * NC) = pack_bf16x2(r[0], r[1]);
          } else {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int nn = nn0 + e;
              if (nn < 4 * NC) {
                const int sub = nn / NC, co = nn % NC;
                const int oyl = 2 * ly_ + (sub >> 1), oxl = 2 * lx_ + (sub & 1);
                s_t[(oyl * kTailO + oxl) * NC + co] = __float2bfloat16(r[e]);
              }
            }
          }
        }
      }
  }
  __syncthreads();
  // ---- phase B: thread = the 2 x 2 outputs of one interior half-resolution pixel
  if (tid < kTailT * kTailT) {
    const int hy = tid / kTailT, hx = tid - hy * kTailT;
    const int H = 2 * A.rh, W = 2 * A.rw;
    const int oy0 = 2 * (ty * kTailT + hy), ox0 = 2 * (tx * kTailT + hx);
    if (oy0 < H && ox0 < W) {          // H, W, oy0, ox0 even: the block is inside or outside
      float win[4][4][NC];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const bf16* tp = s_t + ((2 * hy + 1 + r) * kTailO + (2 * hx + 1 + q)) * NC;
          if (NC == 2) {
            const uint32_t u = *reinterpret_cast<const uint32_t*>(tp);
            win[r][q][0] = bf16_lo(u);
            win[r][q][NC - 1] = bf16_hi(u);
          } else if (NC == 4) {
            const uint2 u = *reinterpret_cast<const uint2*>(tp);
            win[r][q][0] = bf16_lo(u.x);
            win[r][q][1] = bf16_hi(u.x);
            win[r][q][NC - 2] = bf16_lo(u.y);
            win[r][q][NC - 1] = bf16_hi(u.y);
          } else {
#pragma unroll
            for (int ci = 0; ci < NC; ++ci) win[r][q][ci] = __bfloat162float(tp[ci]);
          }
        }
      // the four 3x3 windows advance together so that every weight is read once; each
      // output still accumulates in (row, column, ci) order
      float out[2][2][NC];
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx)
#pragma unroll
          for (int co = 0; co < NC; ++co) out[dy][dx][co] = s_aff[4 * NC + co];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int ci = 0; ci < NC; ++ci) {
            float wv[NC];
#pragma unroll
            for (int co = 0; co < NC; ++co) wv[co] = s_wout[((r * 3 + q) * NC + ci) * NC + co];
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
              for (int dx = 0; dx < 2; ++dx) {
                const float xv = win[dy + r][dx + q][ci];
#pragma unroll
                for (int co = 0; co < NC; ++co) out[dy][dx][co] += xv * wv[co];
              }
          }
      float sg[2][2][NC], lab[2][2];
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          float best = -1.f;
          int bi = 0;
#pragma unroll
          for (int co = 0; co < NC; ++co) {
            // fp32 sigmoid, saturating to exactly 1.0f for large logits like TF's
            const float s1 = 1.f / (1.f + expf(-out[dy][dx][co]));
            sg[dy][dx][co] = s1;
            if (s1 > best) { best = s1; bi = co; }   // strict >: first index on ties
          }
          lab[dy][dx] = (float)bi;
        }
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const int64_t m = ((int64_t)n * H + oy0 + dy) * W + ox0;
        if (NC == 2) {
          if (A.logits)
            *reinterpret_cast<float4*>(A.logits + m * 2) =
                make_float4(out[dy][0][0], out[dy][0][NC - 1], out[dy][1][0], out[dy][1][NC - 1]);
          *reinterpret_cast<float4*>(A.probs + m * 2) =
              make_float4(sg[dy][0][0], sg[dy][0][NC - 1], sg[dy][1][0], sg[dy][1][NC - 1]);
        } else {
#pragma unroll
          for (int dx = 0; dx < 2; ++dx)
#pragma unroll
            for (int co = 0; co < NC; ++co) {
              if (A.logits) A.logits[(m + dx) * NC + co] = out[dy][dx][co];
              A.probs[(m + dx) * NC + co] = sg[dy][dx][co];
            }
        }
        *reinterpret_cast<float2*>(A.labelmap + m) = make_float2(lab[dy][0], lab[dy][1]);
      }
    }
  }
}

__global__ void mc_mean_var_kernel(const float* probs, int T, int64_t count, float* mean,
                                   float* var) {
  GRID_STRIDE(i, count) {
    float mu = 0.f, m2 = 0.f;
    for (int t = 0; t < T; ++t) {       // Welford
      const float v = probs[(int64_t)t * count + i];
      const float d = v - mu;
      mu += d / (float)(t + 1);
      m2 += d * (v - mu);
    }
    mean[i] = mu;
    var[i] = m2 / (float)T;
  }
}

// -------------------------------------------------------------------- Adam
// One block per chunk; a chunk is a contiguous run of <= kAdamChunk master elements
// inside ONE segment (parameter tensor).  The host precomputes the chunk table
// {segment, first element}, so no block searches.  Segments whose shadow needs no
// channel padding take the identity-mapping 16-byte vector path.
constexpr int kAdamChunk = 4096;
constexpr int kAdamThreads = 256;

// sqrt.approx / div.approx (<= 2 ulp each): the update term is <= lr in magnitude, so the
// deviation from IEEE is ~1e-7 * lr — far below the bf16 shadow's resolution — and the
// straight-line code lets the compiler keep every load of the chunk in flight.
__device__ __forceinline__ float adam_one(float g, float& m, float& v, float p, float lr_t,
                                          float b1, float b2, float eps) {
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * g * g;
  float sq;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(sq) : "f"(v));
  return p - lr_t * __fdividef(m, sq + eps);
}

__global__ void __launch_bounds__(kAdamThreads)
adam_multi_kernel(float* __restrict__ param, float* __restrict__ grad, float* __restrict__ m,
                  float* __restrict__ v, bf16* __restrict__ shadow,
                  const int32_t* __restrict__ seg, const int64_t* __restrict__ shadow_off,
                  const int32_t* __restrict__ chunks, float lr_t, const float* lr_t_dev, float b1,
                  float b2, float eps, float gscale) {
  pdl_trigger();
  pdl_wait();
  if (lr_t_dev) lr_t = __ldg(lr_t_dev);
  const int sidx = __ldg(chunks + 2 * blockIdx.x);
  const int begin = __ldg(chunks + 2 * blockIdx.x + 1);
  const int32_t* s = seg + sidx * 6;
  const int64_t base = s[0];
  const int numel = s[1], inner = s[2], inner_pad = s[3], mid = s[4], mid_pad = s[5];
  const int64_t so = shadow_off[sidx];
  const bool identity = inner == inner_pad && mid == mid_pad;
  const int end = min(numel, begin + kAdamChunk);
  const int64_t i0 = base + begin;
  const bool vec_ok = identity && ((i0 & 3) == 0) && (so < 0 || ((so + begin) & 3) == 0);
  if (vec_ok) {
    const int n4 = (end - begin) >> 2;
    float4* g4 = reinterpret_cast<float4*>(grad + i0);
    float4* m4 = reinterpret_cast<float4*>(m + i0);
    float4* v4 = reinterpret_cast<float4*>(v + i0);
    float4* p4 = reinterpret_cast<float4*>(param + i0);
    uint2* s4 = so >= 0 ? reinterpret_cast<uint2*>(shadow + so + begin) : nullptr;
    constexpr int kIt = kAdamChunk / 4 / kAdamThreads;
    float4 g[kIt], mm[kIt], vv[kIt], pp[kIt];
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int q = threadIdx.x + it * kAdamThreads;
      if (q < n4) { g[it] = g4[q]; mm[it] = m4[q]; vv[it] = v4[q]; pp[it] = p4[q]; }
    }
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int q = threadIdx.x + it * kAdamThreads;
      if (q < n4) {
        g4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        pp[it].x = adam_one(g[it].x * gscale, mm[it].x, vv[it].x, pp[it].x, lr_t, b1, b2, eps);
        pp[it].y = adam_one(g[it].y * gscale, mm[it].y, vv[it].y, pp[it].y, lr_t, b1, b2, eps);
        pp[it].z = adam_one(g[it].z * gscale, mm[it].z, vv[it].z, pp[it].z, lr_t, b1, b2, eps);
        pp[it].w = adam_one(g[it].w * gscale, mm[it].w, vv[it].w, pp[it].w, lr_t, b1, b2, eps);
        m4[q] = mm[it];
        v4[q] = vv[it];
        p4[q] = pp[it];
        if (s4) s4[q] = make_uint2(pack_bf16x2(pp[it].x, pp[it].y), pack_bf16x2(pp[it].z, pp[it].w));
      }
    }
    // tail (< 4 elements) falls through to the scalar loop below
    const int done = begin + (n4 << 2);
    for (int e = done + threadIdx.x; e < end; e += kAdamThreads) {
      const int64_t i = base + e;
      float mi = m[i], vi = v[i];
      const float p = adam_one(grad[i] * gscale, mi, vi, param[i], lr_t, b1, b2, eps);
      grad[i] = 0.f; m[i] = mi; v[i] = vi; param[i] = p;
      if (so >= 0) shadow[so + e] = __float2bfloat16(p);
    }
    return;
  }
  for (int e = begin + threadIdx.x; e < end; e += kAdamThreads) {
    const int64_t i = base + e;
    float mi = m[i], vi = v[i];
    const float p = adam_one(grad[i] * gscale, mi, vi, param[i], lr_t, b1, b2, eps);
    grad[i] = 0.f;
    m[i] = mi;
    v[i] = vi;
    param[i] = p;
    if (so >= 0) {
      int64_t si;
      if (identity) {
        si = so + e;
      } else {
        const int in_i = e % inner;
        const int t = e / inner;
        const int mid_i = t % mid;
        const int outer = t / mid;
        si = so + ((int64_t)outer * mid_pad + mid_i) * inner_pad + in_i;
      }
      shadow[si] = __float2bfloat16(p);
    }
  }
}

// ----------------------------------------------------------------- layout
__global__ void pack_input_kernel(const float* x, int C, seg_view y) {
  const int64_t total = (int64_t)y.n * y.h * y.w * y.c;
  GRID_STRIDE(idx, total) {
    const int c = idx % y.c;
    int64_t m = idx / y.c;
    const int xx = m % y.w;
    m /= y.w;
    const int yy = m % y.h;
    const int n = m / y.h;
    const float v = c < C ? x[(((int64_t)n * y.h + yy) * y.w + xx) * C + c] : 0.f;
    view_at_mut(y, n, yy, xx)[c] = __float2bfloat16(v);
  }
}

// dense RGB(A) fp32 -> 16-channel bf16: one thread per pixel, two 16-byte stores
__global__ void pack_input16_kernel(const float* __restrict__ x, int C, bf16* __restrict__ y,
                                    int64_t pixels) {
  pdl_trigger();
  pdl_wait();
  GRID_STRIDE(m, pixels) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < C; ++c) v[c] = __ldg(x + m * C + c);
    uint4 lo, hi;
    lo.x = pack_bf16x2(v[0], v[1]);
    lo.y = pack_bf16x2(v[2], v[3]);
    lo.z = lo.w = 0u;
    hi.x = hi.y = hi.z = hi.w = 0u;
    uint4* dst = reinterpret_cast<uint4*>(y + m * 16);
    dst[0] = lo;
    dst[1] = hi;
  }
}

// Input staging of one step (seg_stage_input): source pixels (fp32 in [0,1], or uint8 to be
// divided by 255 as utils/datasets.py:176-178 does) -> 4 bf16 per pixel (R, G, B, 1) - the
// layout the first-layer kernel (fconv.cuh) reads - with an optional per-image crop window
// (utils/datasets.py:184-185) and the mask that goes with it; thread 0 also runs the step's
// scalar housekeeping so that a training step needs no other node in front of its graph.
struct StageArgs {
  const void* x;
  int kind, C, src_h, src_w;
  const int32_t* crop_yx;
  uint2* y;
  int n, H, W;
  const uint8_t* mask_src;
  int mask_kind;
  uint8_t* mask_dst;
  seg_stage_ctl ctl;
  int has_ctl;
  int vec4;
};

__global__ void stage_input4_kernel(const StageArgs A) {
  pdl_trigger();
  pdl_wait();
  if (A.has_ctl && blockIdx.x == 0 && threadIdx.x == 0) {
    const seg_stage_ctl& c = A.ctl;
    if (c.loss_sum) {
      if (c.host_ring && c.publish_step >= 0) {
        // loss of the step that just finished -> pinned host ring {loss_sum, step id}
        volatile float* slot = c.host_ring + 2 * (c.publish_step & 3);
        slot[0] = *c.loss_sum;
        slot[1] = __int_as_float(c.publish_step);
      }
      *c.loss_sum = 0.f;
    }
    if (c.lr_t_dev) *c.lr_t_dev = c.lr_t;
    if (c.step_dev) *c.step_dev = c.step;
  }
  const int64_t pixels = (int64_t)A.n * A.H * A.W;
  const float one = 1.0f;
  if (A.vec4) {
    // dense fp32 RGB, no crop: four pixels per thread = three 16-byte loads, two 16-byte
    // stores (+ 4 mask bytes)
    const float4* x4 = reinterpret_cast<const float4*>(A.x);
    GRID_STRIDE(q, pixels / 4) {
      const float4 a = __ldg(x4 + 3 * q), b = __ldg(x4 + 3 * q + 1), c = __ldg(x4 + 3 * q + 2);
      uint4* dst = reinterpret_cast<uint4*>(A.y + 4 * q);
      dst[0] = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, one), pack_bf16x2(a.w, b.x),
                          pack_bf16x2(b.y, one));
      dst[1] = make_uint4(pack_bf16x2(b.z, b.w), pack_bf16x2(c.x, one), pack_bf16x2(c.y, c.z),
                          pack_bf16x2(c.w, one));
      if (A.mask_src) {
        uint32_t mk = __ldg(reinterpret_cast<const uint32_t*>(A.mask_src) + q);
        if (A.mask_kind) {
          uint32_t o = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) o |= (uint32_t)(((mk >> (8 * k)) & 0xffu) == 255u) << (8 * k);
          mk = o;
        }
        reinterpret_cast<uint32_t*>(A.mask_dst)[q] = mk;
      }
    }
    return;
  }
  GRID_STRIDE(m, pixels) {
    const int xx = (int)(m % A.W);
    const int64_t t = m / A.W;
    const int yy = (int)(t % A.H);
    const int n = (int)(t / A.H);
    int cy = 0, cx = 0;
    if (A.crop_yx) { cy = __ldg(A.crop_yx + 2 * n); cx = __ldg(A.crop_yx + 2 * n + 1); }
    const int64_t src = ((int64_t)n * A.src_h + yy + cy) * A.src_w + xx + cx;
    float v[3] = {0.f, 0.f, 0.f};
    if (A.kind == 0) {
      const float* xp = reinterpret_cast<const float*>(A.x) + src * A.C;
      for (int c = 0; c < A.C; ++c) v[c] = __ldg(xp + c);
    } else {
      const uint8_t* xp = reinterpret_cast<const uint8_t*>(A.x) + src * A.C;
      for (int c = 0; c < A.C; ++c) v[c] = (float)__ldg(xp + c) / 255.0f;
    }
    A.y[m] = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], one));
    if (A.mask_src) {
      const uint8_t mk = __ldg(A.mask_src + src);
      A.mask_dst[m] = A.mask_kind ? (uint8_t)(mk == 255) : mk;
    }
  }
}

// fp32 NHWC [n,H,W,C] -> bf16 [n,Ho,Wo,CP]: the kh x kw x C patch of every output pixel is
// packed into the channel axis, channel (r*kw+s)*C+c of pixel (oy,ox) =
// x[n, oy*stride+r-pad_t, ox*stride+s-pad_l, c] (zero outside the image and beyond
// kh*kw*C).  A first conv on few input channels then is a 1x1 conv over CP channels.
// Thread = (output pixel, 8-channel group); the 9-fold reuse of x is served by L1/L2.
__global__ void pack_patches_kernel(const float* __restrict__ x, int C, int H, int W, int kh, int kw,
                                    int stride, int pad_t, int pad_l, seg_view y) {
  pdl_trigger();
  pdl_wait();
  const int groups = y.c / 8;
  const int real = kh * kw * C;
  const int64_t total = (int64_t)y.n * y.h * y.w * groups;
  GRID_STRIDE(idx, total) {
    const int g = idx % groups;
    int64_t m = idx / groups;
    const int ox = m % y.w;
    m /= y.w;
    const int oy = m % y.h;
    const int n = m / y.h;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = g * 8 + j;
      v[j] = 0.f;
      if (ch < real) {
        const int t = ch / C, c = ch - t * C;
        const int r = t / kw, sx = t - r * kw;
        const int yy = oy * stride + r - pad_t, xx = ox * stride + sx - pad_l;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W)
          v[j] = __ldg(x + (((int64_t)n * H + yy) * W + xx) * C + c);
      }
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]);
    o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]);
    o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(view_at_mut(y, n, oy, ox) + g * 8) = o;
  }
}

// Compile-time geometry: one thread builds the whole packed pixel in registers (KH*KW*C
// values, fully unrolled, no index arithmetic per channel) and writes it as CP/8 16-byte
// stores; neighbouring threads read overlapping input rows out of L1.
template <int KH, int KW, int C>
__global__ void pack_patches_fixed_kernel(const float* __restrict__ x, int H, int W, int stride,
                                          int pad_t, int pad_l, seg_view y) {
  pdl_trigger();
  pdl_wait();
  constexpr int R = KH * KW * C;
  constexpr int CP = (R + 15) / 16 * 16;
  const int64_t total = (int64_t)y.n * y.h * y.w;
  GRID_STRIDE(m, total) {
    const int ox = m % y.w;
    const int64_t t = m / y.w;
    const int oy = t % y.h;
    const int n = t / y.h;
    float v[CP];
#pragma unroll
    for (int i = R; i < CP; ++i) v[i] = 0.f;
    const int y0 = oy * stride - pad_t, x0 = ox * stride - pad_l;
#pragma unroll
    for (int r = 0; r < KH; ++r) {
      const int yy = y0 + r;
      const bool row_in = yy >= 0 && yy < H;
      const float* xr = x + (((int64_t)n * H + (row_in ? yy : 0)) * W) * C;
#pragma unroll
      for (int sx = 0; sx < KW; ++sx) {
        const int xx = x0 + sx;
        const bool in = row_in && xx >= 0 && xx < W;
#pragma unroll
        for (int c = 0; c < C; ++c) v[(r * KW + sx) * C + c] = in ? __ldg(xr + (int64_t)xx * C + c) : 0.f;
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(view_at_mut(y, n, oy, ox));
#pragma unroll
    for (int g = 0; g < CP / 8; ++g)
      dst[g] = make_uint4(pack_bf16x2(v[8 * g], v[8 * g + 1]), pack_bf16x2(v[8 * g + 2], v[8 * g + 3]),
                          pack_bf16x2(v[8 * g + 4], v[8 * g + 5]), pack_bf16x2(v[8 * g + 6], v[8 * g + 7]));
  }
}

// per-channel sums with 16-byte loads: thread = (pixel lane, 8-channel group)
template <int MODE>   // 0: sum + sumsq   2: sum only
__global__ void channel_sum_vec8_kernel(seg_view a, float* out0, float* out1) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sh[];
  const int C = a.c;
  const int groups = C / 8;
  const int lanes = blockDim.x / groups;
  const int g = threadIdx.x % groups;
  const int pl = threadIdx.x / groups;
  const int64_t pixels = (int64_t)a.n * a.h * a.w;
  float s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s0[j] = 0.f; s1[j] = 0.f; }
  if (pl < lanes) {
    // four pixels per trip, their 16-byte loads issued before any is consumed: these are
    // small tensors (a few MB), so the kernel is bound by load latency, not bandwidth
    const int64_t step = (int64_t)gridDim.x * lanes;
    for (int64_t m0 = (int64_t)blockIdx.x * lanes + pl; m0 < pixels; m0 += 4 * step) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t m = m0 + k * step;
        u[k] = make_uint4(0u, 0u, 0u, 0u);
        if (m < pixels) {
          const int xx = m % a.w;
          const int64_t t = m / a.w;
          const int yy = t % a.h;
          const int n = t / a.h;
          u[k] = *reinterpret_cast<const uint4*>(view_at(a, n, yy, xx) + g * 8);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float lo = bf16_lo(w[j]), hi = bf16_hi(w[j]);
          s0[2 * j] += lo;
          s0[2 * j + 1] += hi;
          if (MODE == 0) { s1[2 * j] += lo * lo; s1[2 * j + 1] += hi * hi; }
        }
      }
    }
  }
  // block reduce over pixel lanes through shared memory: sh[lane][C]
  float* sh0 = sh;
  float* sh1 = sh + blockDim.x * 8;
  if (pl < lanes) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sh0[pl * C + g * 8 + j] = s0[j];
      if (MODE == 0) sh1[pl * C + g * 8 + j] = s1[j];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float r0 = 0.f, r1 = 0.f;
    for (int l = 0; l < lanes; ++l) {
      r0 += sh0[l * C + c];
      if (MODE == 0) r1 += sh1[l * C + c];
    }
    atomicAdd(out0 + c, r0);
    if (MODE == 0) atomicAdd(out1 + c, r1);
  }
}

// ---------------------------------------------------- fused 1x1 head + loss + backward
// The classification head of U-Net is a 1x1 convolution to <= 4 classes (reference
// models/unet.py:166-167): per pixel, logits = x . W + b, softmax cross-entropy against the
// label, dlogits, the head's weight / bias gradient and the input gradient
// dx = relu_mask(dlogits . W^T) are all functions of that pixel's CIN channels.  One pass
// over x (bf16) replaces five launches (conv fwd, loss, wgrad, dgrad, memset) at the
// forward/backward turning point where nothing else can overlap.  Arithmetic matches the
// unfused path: bf16 operands, fp32 accumulation, dlogits rounded to bf16 before use.
template <int CIN, int CO>
__global__ void __launch_bounds__(256)
head1x1_xent_kernel(seg_view x, const bf16* __restrict__ w, int cout_pad,
                    const float* __restrict__ bias, seg_view labels, seg_view logits,
                    float* loss_sum, seg_view dx, float* dw, float* db, float inv_pixels) {
  // two threads (a lane pair) per pixel, CIN/2 channels each: halves the per-thread
  // weight-gradient accumulators so two blocks fit an SM
  constexpr int HC = CIN / 2;
  pdl_trigger();
  __shared__ float sw[CIN * CO];
  __shared__ float sb[CO];
  __shared__ float red[8][CIN * CO + CO + 1];
  pdl_wait();
  for (int i = threadIdx.x; i < CIN * CO; i += blockDim.x)
    sw[i] = __bfloat162float(w[(int64_t)(i / CO) * cout_pad + (i % CO)]);
  if (threadIdx.x < CO) sb[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
  __syncthreads();
  const int h = threadIdx.x & 1;                  // which half of the channels
  const float* swh = sw + h * HC * CO;
  float gw[HC * CO];
  float gb[CO];
#pragma unroll
  for (int i = 0; i < HC * CO; ++i) gw[i] = 0.f;
#pragma unroll
  for (int c = 0; c < CO; ++c) gb[c] = 0.f;
  float local = 0.f;
  const int64_t pixels = (int64_t)x.n * x.h * x.w;
  // every lane runs the same number of iterations (the pair shuffle below is full-warp);
  // lanes past the end recompute the last pixel with a zero gradient and no stores
  const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 1;
  const int64_t first = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 1;
  const int64_t iters = (pixels + stride - 1) / stride;
  for (int64_t it = 0; it < iters; ++it) {
    int64_t m = first + it * stride;
    const bool live = m < pixels;
    if (!live) m = pixels - 1;
    const int xx = m % x.w;
    const int64_t t = m / x.w;
    const int yy = t % x.h;
    const int n = t / x.h;
    float xv[HC];
    const uint4* xp = reinterpret_cast<const uint4*>(view_at(x, n, yy, xx) + h * HC);
#pragma unroll
    for (int q = 0; q < HC / 8; ++q) {
      const uint4 u = xp[q];
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { xv[8 * q + 2 * e] = bf16_lo(w4[e]); xv[8 * q + 2 * e + 1] = bf16_hi(w4[e]); }
    }
    float lg[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) lg[c] = 0.f;
#pragma unroll
    for (int k = 0; k < HC; ++k)
#pragma unroll
      for (int c = 0; c < CO; ++c) lg[c] += xv[k] * swh[k * CO + c];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < CO; ++c) {
      // the pair's partial sums in a fixed order (even-lane half first) on both lanes
      const float other = __shfl_xor_sync(0xffffffffu, lg[c], 1);
      lg[c] = (h == 0 ? lg[c] + other : other + lg[c]) + sb[c];
      mx = fmaxf(mx, lg[c]);
    }
    if (logits.ptr && h == 0 && live) {
      float* lp = reinterpret_cast<float*>(logits.ptr) + n * logits.sn + yy * logits.sh + xx * logits.sw;
#pragma unroll
      for (int c = 0; c < CO; ++c) lp[c] = lg[c];
    }
    const int lab = reinterpret_cast<const uint8_t*>(labels.ptr)[n * labels.sn + yy * labels.sh +
                                                                 xx * labels.sw];
    float e[CO], se = 0.f, picked = 0.f;
#pragma unroll
    for (int c = 0; c < CO; ++c) {
      e[c] = expf(lg[c] - mx);
      se += e[c];
      if (c == lab) picked = lg[c];
    }
    if (h == 0 && live) local += mx + logf(se) - picked;
    const float inv_se = 1.f / se;
    float dl[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) {
      dl[c] = live ? __bfloat162float(__float2bfloat16((e[c] * inv_se - (c == lab ? 1.f : 0.f)) * inv_pixels))
                   : 0.f;
      if (h == 0) gb[c] += dl[c];
    }
    uint32_t o[HC / 2];
#pragma unroll
    for (int k = 0; k < HC; k += 2) {
      float g0 = 0.f, g1 = 0.f;
#pragma unroll
      for (int c = 0; c < CO; ++c) {
        g0 += dl[c] * swh[k * CO + c];
        g1 += dl[c] * swh[(k + 1) * CO + c];
        gw[k * CO + c] += xv[k] * dl[c];
        gw[(k + 1) * CO + c] += xv[k + 1] * dl[c];
      }
      o[k / 2] = pack_bf16x2(xv[k] > 0.f ? g0 : 0.f, xv[k + 1] > 0.f ? g1 : 0.f);
    }
    if (live) {
      uint4* dp = reinterpret_cast<uint4*>(view_at_mut(dx, n, yy, xx) + h * HC);
#pragma unroll
      for (int q = 0; q < HC / 8; ++q) dp[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
    }
  }
  // block reduction of dW, db and the loss: lanes of equal parity hold the same channel half
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < HC * CO; ++i) {
    float v = gw[i];
#pragma unroll
    for (int o2 = 16; o2 >= 2; o2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o2);
    if (lane < 2) red[warp][lane * HC * CO + i] = v;
  }
#pragma unroll
  for (int c = 0; c < CO; ++c) {
    const float v = warp_sum(gb[c]);
    if (lane == 0) red[warp][CIN * CO + c] = v;
  }
  local = warp_sum(local);
  if (lane == 0) red[warp][CIN * CO + CO] = local;
  __syncthreads();
  for (int i = threadIdx.x; i < CIN * CO + CO + 1; i += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int wq = 0; wq < 8; ++wq) v += red[wq][i];
    if (i < CIN * CO) atomicAdd(dw + i, v);
    else if (i < CIN * CO + CO) atomicAdd(db + (i - CIN * CO), v);
    else atomicAdd(loss_sum, v);
  }
}

template <int CIN>
static int launch_head1x1(int co, const seg_view& x, const bf16* w, int cout_pad, const float* bias,
                          const seg_view& labels, const seg_view& logits, float* loss_sum,
                          const seg_view& dx, float* dw, float* db, cudaStream_t st) {
  const int64_t pixels = (int64_t)x.n * x.h * x.w;
  int grid = num_sms() * 2;                                                  // one resident wave
  if ((int64_t)grid * 128 > pixels) grid = (int)ceil_div64(pixels, 128);     // 2 threads per pixel
  const float inv = 1.f / (float)pixels;
  switch (co) {
    case 2: SEG_CHECK_CUDA(launch_k(head1x1_xent_kernel<CIN, 2>, dim3(grid), dim3(256), (size_t)0, st, x, w, cout_pad, bias, labels, logits, loss_sum, dx, dw, db, inv)); return SEG_OK;
    case 3: SEG_CHECK_CUDA(launch_k(head1x1_xent_kernel<CIN, 3>, dim3(grid), dim3(256), (size_t)0, st, x, w, cout_pad, bias, labels, logits, loss_sum, dx, dw, db, inv)); return SEG_OK;
    case 4: SEG_CHECK_CUDA(launch_k(head1x1_xent_kernel<CIN, 4>, dim3(grid), dim3(256), (size_t)0, st, x, w, cout_pad, bias, labels, logits, loss_sum, dx, dw, db, inv)); return SEG_OK;
  }
  set_error("head1x1_xent: 2..4 classes supported, got %d", co);
  return SEG_E_UNSUPPORTED;
}

}  // namespace segb

using namespace segb;

// =========================================================================
// C ABI
// =========================================================================
extern "C" {

// Launch geometry of the row-mapped pool kernels: `rows` block rows of `rowlen` 16-byte
// items.  A thread's setup (row base pointers) is amortised over up to 8 items of its row
// as long as the grid still fills the machine with ~2048 threads per SM.
static inline void pool_row_geometry(int64_t rows, int rowlen, dim3* grid, dim3* block) {
  int64_t ipt = rows * rowlen / ((int64_t)num_sms() * 2048);
  ipt = ipt < 1 ? 1 : (ipt > 8 ? 8 : ipt);
  int bt = (int)(((rowlen + ipt - 1) / ipt + 31) / 32 * 32);
  bt = bt > 256 ? 256 : bt;
  *block = dim3(bt);
  *grid = dim3((unsigned)((rowlen + bt * ipt - 1) / (bt * ipt)), (unsigned)rows);
}

SEG_API int32_t seg_maxpool_fwd(const seg_view* x, int32_t k, int32_t s, const seg_view* y,
                        uint8_t* argmax, void* stream) {
  SEG_REQUIRE(x && y && argmax, SEG_E_BAD_SHAPE, "maxpool_fwd: null argument");
  SEG_REQUIRE(k >= 1 && s >= 1 && y->h == (x->h - k) / s + 1 && y->w == (x->w - k) / s + 1 &&
                  y->c == x->c && y->n == x->n,
              SEG_E_BAD_SHAPE, "maxpool_fwd: bad geometry");
  cudaStream_t st = (cudaStream_t)stream;
  const bool v8 = vec8_ok(*x) && vec8_ok(*y) && (reinterpret_cast<uintptr_t>(argmax) % 8) == 0;
  const int64_t total = (int64_t)y->n * y->h * y->w * (v8 ? y->c / 8 : y->c);
  const int64_t f_rows = (int64_t)y->n * y->h;
  if (v8 && g_pool_rows && k == s && k <= 8 && f_rows <= 65535 && f_rows > 0) {
    const int rowlen = y->w * (y->c / 8);
    dim3 grid, block;
    pool_row_geometry(f_rows, rowlen, &grid, &block);
    if (k == 2)
      SEG_CHECK_CUDA(launch_k(maxpool_fwd_row8_kernel<2>, grid, block, (size_t)(0), st, *x, k, *y, argmax));
    else if (k == 3)
      SEG_CHECK_CUDA(launch_k(maxpool_fwd_row8_kernel<3>, grid, block, (size_t)(0), st, *x, k, *y, argmax));
    else
      SEG_CHECK_CUDA(launch_k(maxpool_fwd_row8_kernel<0>, grid, block, (size_t)(0), st, *x, k, *y, argmax));
  } else if (v8)
    SEG_CHECK_CUDA(launch_k(maxpool_fwd_kernel<8>, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), st, *x, k, s, *y, argmax));
  else
    maxpool_fwd_kernel<1><<<grid_for(total, 256), 256, 0, st>>>(*x, k, s, *y, argmax);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

SEG_API int32_t seg_maxpool_bn_infer(const seg_view* x, int32_t k, const float* moving_mean,
                                     const float* moving_var, float eps, const float* beta,
                                     int32_t bn_c, const seg_view* y, void* stream) {
  SEG_REQUIRE(x && y && moving_mean && moving_var && beta, SEG_E_BAD_SHAPE,
              "maxpool_bn_infer: null argument");
  SEG_REQUIRE(k >= 1 && k <= 8 && y->h == x->h / k && y->w == x->w / k && y->c == x->c &&
                  y->n == x->n && bn_c >= 0 && bn_c <= x->c,
              SEG_E_BAD_SHAPE, "maxpool_bn_infer: bad geometry");
  const int64_t f_rows = (int64_t)y->n * y->h;
  SEG_REQUIRE(vec8_ok(*x) && vec8_ok(*y) && f_rows > 0 && f_rows <= 65535, SEG_E_UNSUPPORTED,
              "maxpool_bn_infer: needs 16-byte aligned channel vectors and n*h_out <= 65535");
  cudaStream_t st = (cudaStream_t)stream;
  const int rowlen = y->w * (y->c / 8);
  dim3 grid, block;
  pool_row_geometry(f_rows, rowlen, &grid, &block);
  if (k == 2)
    SEG_CHECK_CUDA(launch_k(maxpool_bn_infer_row8_kernel<2>, grid, block, (size_t)0, st, *x, k,
                            moving_mean, moving_var, eps, beta, bn_c, *y));
  else if (k == 3)
    SEG_CHECK_CUDA(launch_k(maxpool_bn_infer_row8_kernel<3>, grid, block, (size_t)0, st, *x, k,
                            moving_mean, moving_var, eps, beta, bn_c, *y));
  else
    SEG_CHECK_CUDA(launch_k(maxpool_bn_infer_row8_kernel<0>, grid, block, (size_t)0, st, *x, k,
                            moving_mean, moving_var, eps, beta, bn_c, *y));
  return SEG_OK;
}

// row-mapped backward launch; false: geometry outside its limits (caller uses the cell form)
static bool launch_pool_bwd_rows(const seg_view& dy, const seg_view& dy2, const uint8_t* argmax,
                                 int k, const seg_view& add, int add_y0, int add_x0,
                                 const seg_view& mask, const seg_view& pooled, const seg_view& dx,
                                 cudaStream_t st, cudaError_t* err) {
  const int64_t rows = (int64_t)dx.n * ((dx.h + k - 1) / k);
  if (!g_pool_rows || k > 8 || rows > 65535 || rows <= 0) return false;
  const int rowlen = dx.w * (dx.c / 8);
  dim3 grid, block;
  pool_row_geometry(rows, rowlen, &grid, &block);
  const bool two = dy2.ptr != nullptr;
#define SEG_POOL_BWD(KK, TT) \
  *err = launch_k(maxpool_bwd_row8_kernel<KK, TT>, grid, block, (size_t)(0), st, dy, dy2, argmax, k, \
                  add, add_y0, add_x0, mask, pooled, dx)
  if (k == 2) { if (two) SEG_POOL_BWD(2, true); else SEG_POOL_BWD(2, false); }
  else if (k == 3) { if (two) SEG_POOL_BWD(3, true); else SEG_POOL_BWD(3, false); }
  else { if (two) SEG_POOL_BWD(0, true); else SEG_POOL_BWD(0, false); }
#undef SEG_POOL_BWD
  return true;
}

static int maxpool_bwd_impl(const seg_view* dy, const uint8_t* argmax, int32_t k, int32_t s,
                            const seg_view* add, int32_t add_y0, int32_t add_x0,
                            const seg_view* mask_src, const seg_view* pooled, const seg_view* dx,
                            void* stream) {
  SEG_REQUIRE(dy && argmax && dx, SEG_E_BAD_SHAPE, "maxpool_bwd: null argument");
  SEG_REQUIRE(k == s, SEG_E_UNSUPPORTED, "maxpool_bwd: only non-overlapping windows (k == s)");
  cudaStream_t st = (cudaStream_t)stream;
  const seg_view a = add ? *add : null_view();
  const seg_view mk = mask_src ? *mask_src : null_view();
  const bool use_y = pooled && pooled->ptr && mask_src && vec8_ok(*pooled) && pooled->h == dy->h &&
                     pooled->w == dy->w && pooled->c == dy->c && pooled->n == dy->n;
  const seg_view py = use_y ? *pooled : null_view();
  const bool v8 = vec8_ok(*dx) && vec8_ok(*dy) && (reinterpret_cast<uintptr_t>(argmax) % 8) == 0 &&
                  (!add || vec8_ok(a)) && (!mask_src || vec8_ok(mk));
  const int64_t total = (int64_t)dx->n * dx->h * dx->w * (v8 ? dx->c / 8 : dx->c);
  if (v8)
  {
    cudaError_t e = cudaSuccess;
    if (launch_pool_bwd_rows(*dy, null_view(), argmax, k, a, add_y0, add_x0, mk, py, *dx, st, &e)) {
      SEG_CHECK_CUDA(e);
    } else {
      const int64_t cells = (int64_t)dx->n * ((dx->h + k - 1) / k) * ((dx->w + k - 1) / k) * (dx->c / 8);
      SEG_CHECK_CUDA(launch_k(maxpool_bwd_cell8_kernel, dim3(grid_for(cells, 256)), dim3(256), (size_t)(0), st, *dy, null_view(), argmax, k, a, add_y0, add_x0, mk, py, *dx));
    }
  }
  else
    maxpool_bwd_kernel<1><<<grid_for(total, 256), 256, 0, st>>>(*dy, null_view(), argmax, k, s, a,
                                                               add_y0, add_x0, mk, *dx);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

SEG_API int32_t seg_maxpool_bwd(const seg_view* dy, const uint8_t* argmax, int32_t k, int32_t s,
                        const seg_view* add, int32_t add_y0, int32_t add_x0,
                        const seg_view* mask_src, const seg_view* dx, void* stream) {
  return maxpool_bwd_impl(dy, argmax, k, s, add, add_y0, add_x0, mask_src, nullptr, dx, stream);
}

SEG_API int32_t seg_maxpool_bwd_y(const seg_view* dy, const uint8_t* argmax, int32_t k, int32_t s,
                          const seg_view* add, int32_t add_y0, int32_t add_x0,
                          const seg_view* mask_src, const seg_view* pooled_y, const seg_view* dx,
                          void* stream) {
  return maxpool_bwd_impl(dy, argmax, k, s, add, add_y0, add_x0, mask_src, pooled_y, dx, stream);
}

static int maxpool_bwd2_impl(const seg_view* dy, const seg_view* dy2, const uint8_t* argmax,
                             int32_t k, int32_t s, const seg_view* mask_src,
                             const seg_view* pooled, const seg_view* dx, void* stream) {
  SEG_REQUIRE(dy && dy2 && argmax && dx, SEG_E_BAD_SHAPE, "maxpool_bwd2: null argument");
  SEG_REQUIRE(k == s, SEG_E_UNSUPPORTED, "maxpool_bwd2: only non-overlapping windows (k == s)");
  cudaStream_t st = (cudaStream_t)stream;
  const seg_view mk = mask_src ? *mask_src : null_view();
  const bool v8 = vec8_ok(*dx) && vec8_ok(*dy) && vec8_ok(*dy2) &&
                  (reinterpret_cast<uintptr_t>(argmax) % 8) == 0 && (!mask_src || vec8_ok(mk));
  // the pool output as the ReLU mask of the routed gradient (x at the argmax IS the pooled
  // value): the pool input is then not read at all
  const bool use_y = v8 && pooled && pooled->ptr && mask_src && vec8_ok(*pooled) &&
                     pooled->h == dy->h && pooled->w == dy->w && pooled->c == dy->c &&
                     pooled->n == dy->n;
  const seg_view py = use_y ? *pooled : null_view();
  const int64_t total = (int64_t)dx->n * dx->h * dx->w * (v8 ? dx->c / 8 : dx->c);
  if (v8)
  {
    cudaError_t e = cudaSuccess;
    if (launch_pool_bwd_rows(*dy, *dy2, argmax, k, null_view(), 0, 0, mk, py, *dx, st, &e)) {
      SEG_CHECK_CUDA(e);
    } else {
      const int64_t cells = (int64_t)dx->n * ((dx->h + k - 1) / k) * ((dx->w + k - 1) / k) * (dx->c / 8);
      SEG_CHECK_CUDA(launch_k(maxpool_bwd_cell8_kernel, dim3(grid_for(cells, 256)), dim3(256), (size_t)(0), st, *dy, *dy2, argmax, k, null_view(), 0, 0, mk, py, *dx));
    }
  }
  else
    maxpool_bwd_kernel<1><<<grid_for(total, 256), 256, 0, st>>>(*dy, *dy2, argmax, k, s,
                                                               null_view(), 0, 0, mk, *dx);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

SEG_API int32_t seg_maxpool_bwd2(const seg_view* dy, const seg_view* dy2, const uint8_t* argmax,
                         int32_t k, int32_t s, const seg_view* mask_src, const seg_view* dx,
                         void* stream) {
  return maxpool_bwd2_impl(dy, dy2, argmax, k, s, mask_src, nullptr, dx, stream);
}

SEG_API int32_t seg_maxpool_bwd2_y(const seg_view* dy, const seg_view* dy2, const uint8_t* argmax,
                                   int32_t k, int32_t s, const seg_view* mask_src,
                                   const seg_view* pooled_y, const seg_view* dx, void* stream) {
  return maxpool_bwd2_impl(dy, dy2, argmax, k, s, mask_src, pooled_y, dx, stream);
}

SEG_API int32_t seg_relu_grad(const seg_view* dy, const seg_view* y, const seg_view* dz, void* stream) {
  SEG_REQUIRE(dy && y && dz, SEG_E_BAD_SHAPE, "relu_grad: null argument");
  const int64_t total = (int64_t)dz->n * dz->h * dz->w * dz->c;
  relu_grad_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*dy, *y, *dz);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

SEG_API int32_t seg_bilinear_upsample_fwd(const seg_view* x, int32_t factor, const seg_view* add,
                                  const seg_view* y, int32_t y_is_f32, void* stream) {
  SEG_REQUIRE(x && y && factor >= 1 && factor <= 32, SEG_E_BAD_SHAPE,
              "bilinear_upsample_fwd: bad argument");
  SEG_REQUIRE(y->h == x->h * factor && y->w == x->w * factor && y->c == x->c, SEG_E_BAD_SHAPE,
              "bilinear_upsample_fwd: y must be x scaled by factor");
  int G = 1;
  for (int g = 8; g >= 2; --g)
    if (y->c % g == 0) { G = g; break; }
  const int64_t total = (int64_t)y->n * y->h * y->w * (y->c / G);
  SEG_REQUIRE(total < (int64_t)1 << 31, SEG_E_UNSUPPORTED, "bilinear_upsample_fwd: tensor too large");
  bilinear_up_fwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      *x, factor, add ? *add : null_view(), *y, y_is_f32, G);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

SEG_API int32_t seg_bilinear_upsample_bwd(const seg_view* dy, int32_t dy_is_f32, int32_t factor,
                                  const seg_view* mask_src, const seg_view* dx, void* stream) {
  SEG_REQUIRE(dy && dx && factor >= 1 && factor <= 32, SEG_E_BAD_SHAPE,
              "bilinear_upsample_bwd: bad argument");
  const int64_t total = (int64_t)dx->n * dx->h * dx->w * dx->c;
  bilinear_up_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      *dy, dy_is_f32, factor, mask_src ? *mask_src : null_view(), *dx);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

SEG_API int32_t seg_resize_bilinear_fwd(const seg_view* x, const seg_view* y, void* stream) {
  SEG_REQUIRE(x && y && x->c == y->c && x->n == y->n, SEG_E_BAD_SHAPE, "resize_bilinear_fwd");
  const int64_t total = (int64_t)y->n * y->h * y->w * y->c;
  if (vec8_ok(*x) && vec8_ok(*y))
    SEG_CHECK_CUDA(launch_k(resize_bilinear_fwd_vec8_kernel, dim3(grid_for(total / 8, 256)), dim3(256),
                            (size_t)0, (cudaStream_t)stream, *x, *y));
  else
    resize_bilinear_fwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*x, *y);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

SEG_API int32_t seg_resize_bilinear_bwd(const seg_view* dy, const seg_view* dx, void* stream) {
  SEG_REQUIRE(dy && dx && dx->c == dy->c && dx->n == dy->n, SEG_E_BAD_SHAPE, "resize_bilinear_bwd");
  const int64_t total = (int64_t)dx->n * dx->h * dx->w * dx->c;
  resize_bilinear_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*dy, *dx);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

static int launch_channel_reduce(const seg_view& a, const seg_view& b, const float* mean,
                                 const float* rstd, int mode, float* o0, float* o1,
                                 cudaStream_t st) {
  const int64_t pixels = (int64_t)a.n * a.h * a.w;
  const int block = 256;
  if ((mode == 0 || mode == 2) && vec8_ok(a) && a.c / 8 <= block) {
    const int lanes_v = block / (a.c / 8);
    int64_t gv = ceil_div64(pixels, (int64_t)lanes_v * 4);
    const int64_t capv = (int64_t)num_sms() * 8;
    if (gv > capv) gv = capv;
    if (gv < 1) gv = 1;
    const size_t shb = (size_t)2 * block * 8 * sizeof(float);
    if (mode == 0)
      SEG_CHECK_CUDA(launch_k(channel_sum_vec8_kernel<0>, dim3((int)gv), dim3(block), (size_t)(shb), st, a, o0, o1));
    else
      SEG_CHECK_CUDA(launch_k(channel_sum_vec8_kernel<2>, dim3((int)gv), dim3(block), (size_t)(shb), st, a, o0, o1));
    SEG_LAUNCH_CHECK();
    return SEG_OK;
  }
  const int lanes = block / a.c > 0 ? block / a.c : 1;
  int64_t g = ceil_div64(pixels, (int64_t)lanes * 64);
  const int64_t cap = (int64_t)num_sms() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  channel_reduce2_kernel<<<(int)g, block, 2 * block * sizeof(float), st>>>(a, b, mean, rstd, mode,
                                                                           o0, o1);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

SEG_API int32_t seg_bias_grad(const seg_view* dz, float* db, void* stream) {
  SEG_REQUIRE(dz && db, SEG_E_BAD_SHAPE, "bias_grad: null argument");
  return launch_channel_reduce(*dz, null_view(), nullptr, nullptr, 2, db, nullptr,
                               (cudaStream_t)stream);
}

SEG_API int32_t seg_batchnorm_stats(const seg_view* x, float* sum, float* sumsq, void* stream) {
  SEG_REQUIRE(x && sum && sumsq, SEG_E_BAD_SHAPE, "batchnorm_stats: null argument");
  return launch_channel_reduce(*x, null_view(), nullptr, nullptr, 0, sum, sumsq,
                               (cudaStream_t)stream);
}

SEG_API int32_t seg_batchnorm_finalize(const float* sum, const float* sumsq, int64_t count, int32_t c,
                               float eps, float decay, float* mean, float* rstd,
                               float* moving_mean, float* moving_var, void* stream) {
  SEG_REQUIRE(sum && sumsq && mean && rstd && count > 0, SEG_E_BAD_SHAPE, "batchnorm_finalize");
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      sum, sumsq, 1.f / (float)count, c, eps, decay, mean, rstd, moving_mean, moving_var);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

SEG_API int32_t seg_batchnorm_apply(const seg_view* x, const float* mean, const float* rstd,
                            const float* beta, const seg_view* y, void* stream) {
  SEG_REQUIRE(x && y && mean && rstd && beta, SEG_E_BAD_SHAPE, "batchnorm_apply");
  const int64_t total = (int64_t)x->n * x->h * x->w * x->c;
  if (vec8_ok(*x) && vec8_ok(*y))
    SEG_CHECK_CUDA(launch_k(bn_apply_vec8_kernel, dim3(grid_for(total / 8, 256)), dim3(256), (size_t)0,
                            (cudaStream_t)stream, *x, mean, rstd, 0.f, 0, beta, *y));
  else
    bn_apply_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*x, mean, rstd, 0.f, 0,
                                                                            beta, *y);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

SEG_API int32_t seg_batchnorm_infer(const seg_view* x, const float* moving_mean, const float* moving_var,
                            float eps, const float* beta, const seg_view* y, void* stream) {
  SEG_REQUIRE(x && y && moving_mean && moving_var && beta, SEG_E_BAD_SHAPE, "batchnorm_infer");
  const int64_t total = (int64_t)x->n * x->h * x->w * x->c;
  if (vec8_ok(*x) && vec8_ok(*y))
    SEG_CHECK_CUDA(launch_k(bn_apply_vec8_kernel, dim3(grid_for(total / 8, 256)), dim3(256), (size_t)0,
                            (cudaStream_t)stream, *x, moving_mean, moving_var, eps, 1, beta, *y));
  else
    bn_apply_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        *x, moving_mean, moving_var, eps, 1, beta, *y);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

SEG_API int32_t seg_batchnorm_fold(const float* moving_mean, const float* moving_var, float eps,
                                   const float* beta, int32_t c, int32_t c_pad, float* scale,
                                   float* shift, void* stream) {
  SEG_REQUIRE(moving_mean && moving_var && beta && scale && shift && c >= 0 && c <= c_pad,
              SEG_E_BAD_SHAPE, "batchnorm_fold: bad argument");
  if (c_pad == 0) return SEG_OK;
  SEG_CHECK_CUDA(launch_k(bn_fold_kernel, dim3((c_pad + 127) / 128), dim3(128), (size_t)0,
                          (cudaStream_t)stream, moving_mean, moving_var, eps, beta, (int)c,
                          (int)c_pad, scale, shift));
  return SEG_OK;
}

SEG_API int32_t seg_batchnorm_bwd_reduce(const seg_view* dy, const seg_view* x, const float* mean,
                                 const float* rstd, float* dbeta, float* dxhat, void* stream) {
  SEG_REQUIRE(dy && x && mean && rstd && dbeta && dxhat, SEG_E_BAD_SHAPE, "batchnorm_bwd_reduce");
  return launch_channel_reduce(*dy, *x, mean, rstd, 1, dbeta, dxhat, (cudaStream_t)stream);
}

SEG_API int32_t seg_batchnorm_bwd_apply(const seg_view* dy, const seg_view* x, const float* mean,
                                const float* rstd, const float* dbeta, const float* dxhat,
                                int64_t count, int32_t relu_mask, const seg_view* dx,
                                void* stream) {
  SEG_REQUIRE(dy && x && dx && count > 0, SEG_E_BAD_SHAPE, "batchnorm_bwd_apply");
  const int64_t total = (int64_t)x->n * x->h * x->w * x->c;
  bn_bwd_apply_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      *dy, *x, mean, rstd, dbeta, dxhat, 1.f / (float)count, relu_mask, *dx);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

SEG_API int32_t seg_dropout(const seg_view* x, uint64_t seed, uint32_t stream_id, float keep_prob,
                    const seg_view* y, void* stream) {
  SEG_REQUIRE(x && y && view_dense(*x) && view_dense(*y), SEG_E_BAD_SHAPE,
              "dropout: dense views required");
  const int64_t numel = (int64_t)x->n * x->h * x->w * x->c;
  dropout_kernel<<<grid_for((numel + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const bf16*>(x->ptr), reinterpret_cast<bf16*>(y->ptr), numel,
      (uint32_t)(seed & 0xFFFFFFFFu), (uint32_t)(seed >> 32), stream_id, keep_prob,
      1.f / keep_prob);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

SEG_API int32_t seg_dropout_ex(const seg_view* x, uint64_t seed, uint32_t stream_id0,
                               uint32_t per_image_step, const uint32_t* step_dev,
                               uint32_t step_mul, float keep_prob, const seg_view* y,
                               void* stream) {
  SEG_REQUIRE(x && y && view_dense(*x) && view_dense(*y), SEG_E_BAD_SHAPE,
              "dropout_ex: dense views required");
  const int64_t img = (int64_t)x->h * x->w * x->c;
  SEG_REQUIRE(img % 8 == 0 && (reinterpret_cast<uintptr_t>(x->ptr) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(y->ptr) & 15) == 0,
              SEG_E_ALIGN, "dropout_ex: h*w*c must be a multiple of 8 and the tensors 16-byte aligned");
  dropout_ex_kernel<<<grid_for(img / 8 * x->n, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const bf16*>(x->ptr), reinterpret_cast<bf16*>(y->ptr), img, x->n,
      (uint32_t)(seed & 0xFFFFFFFFu), (uint32_t)(seed >> 32), stream_id0, per_image_step, step_dev,
      step_mul, keep_prob, 1.f / keep_prob);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

SEG_API int32_t seg_softmax_xent_fwd_bwd(const seg_view* logits, const seg_view* labels, float* loss_sum,
                                 const seg_view* dlogits, void* stream) {
  SEG_REQUIRE(logits && labels && loss_sum, SEG_E_BAD_SHAPE, "softmax_xent: null argument");
  SEG_REQUIRE(labels->h == logits->h && labels->w == logits->w && labels->n == logits->n,
              SEG_E_BAD_SHAPE, "softmax_xent: labels must match logits spatially");
  const int64_t pixels = (int64_t)logits->n * logits->h * logits->w;
  SEG_CHECK_CUDA(launch_k(softmax_xent_kernel, dim3(grid_for(pixels, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, *logits, *labels, loss_sum, dlogits ? *dlogits : null_view(), 1.f / (float)pixels));
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

SEG_API int32_t seg_upscore8_xent_fwd_bwd(const seg_view* x, const seg_view* labels,
                                          float* loss_sum, const seg_view* dx,
                                          const seg_view* mask_src, float* logits, void* stream) {
  SEG_REQUIRE(x && labels && loss_sum && dx, SEG_E_BAD_SHAPE, "upscore8_xent: null argument");
  SEG_REQUIRE(labels->n == x->n && labels->h == x->h * kUxF && labels->w == x->w * kUxF &&
                  dx->n == x->n && dx->h == x->h && dx->w == x->w && dx->c == x->c &&
                  (!mask_src || (mask_src->n == x->n && mask_src->h == x->h &&
                                 mask_src->w == x->w && mask_src->c == x->c)),
              SEG_E_BAD_SHAPE, "upscore8_xent: labels must be 8x the score map, dx / mask its shape");
  SEG_REQUIRE(x->c >= 1 && x->c <= 32, SEG_E_UNSUPPORTED, "upscore8_xent: 1..32 classes");
  UpXentArgs A;
  memset(&A, 0, sizeof(A));
  A.x = *x; A.labels = *labels; A.dx = *dx;
  A.mask = mask_src ? *mask_src : null_view();
  A.loss_sum = loss_sum; A.logits = logits;
  const int64_t pixels = (int64_t)x->n * x->h * kUxF * x->w * kUxF;
  A.inv_pixels = 1.f / (float)pixels;
  A.C = x->c;
  A.tiles_x = (x->w + kUxTI - 1) / kUxTI; A.tiles_y = (x->h + kUxTI - 1) / kUxTI;
  const int64_t blocks = (int64_t)x->n * A.tiles_x * A.tiles_y;
  SEG_REQUIRE(blocks < ((int64_t)1 << 31), SEG_E_UNSUPPORTED, "upscore8_xent: tensor too large");
  const size_t smem = ux_smem_bytes(x->c);
  cudaStream_t st = (cudaStream_t)stream;
  if (x->c == 21) {
    static bool attr21 = false;
    if (!attr21) {
      SEG_CHECK_CUDA(cudaFuncSetAttribute(upscore8_xent_kernel<21>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr21 = true;
    }
    SEG_CHECK_CUDA(launch_k(upscore8_xent_kernel<21>, dim3((unsigned)blocks), dim3(256), smem, st, A));
  } else {
    static size_t attr_any = 0;
    if (smem > attr_any) {
      SEG_CHECK_CUDA(cudaFuncSetAttribute(upscore8_xent_kernel<0>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_any = smem;
    }
    SEG_CHECK_CUDA(launch_k(upscore8_xent_kernel<0>, dim3((unsigned)blocks), dim3(256), smem, st, A));
  }
  return SEG_OK;
}

SEG_API int32_t seg_head1x1_xent(const seg_view* x, const void* w_bf16, int32_t cout_pad,
                                 const float* bias, const seg_view* labels, int32_t n_classes,
                                 const seg_view* logits, float* loss_sum, const seg_view* dx,
                                 float* dw, float* db, void* stream) {
  SEG_REQUIRE(x && w_bf16 && labels && loss_sum && dx && dw && db, SEG_E_BAD_SHAPE,
              "head1x1_xent: null argument");
  SEG_REQUIRE(labels->h == x->h && labels->w == x->w && labels->n == x->n && dx->h == x->h &&
                  dx->w == x->w && dx->n == x->n && dx->c == x->c,
              SEG_E_BAD_SHAPE, "head1x1_xent: geometry mismatch");
  SEG_REQUIRE(vec8_ok(*x) && vec8_ok(*dx), SEG_E_ALIGN, "head1x1_xent: x / dx must allow 16-byte rows");
  const seg_view lg = logits ? *logits : null_view();
  const bf16* w = reinterpret_cast<const bf16*>(w_bf16);
  cudaStream_t st = (cudaStream_t)stream;
  switch (x->c) {
    case 16: return launch_head1x1<16>(n_classes, *x, w, cout_pad, bias, *labels, lg, loss_sum, *dx, dw, db, st);
    case 32: return launch_head1x1<32>(n_classes, *x, w, cout_pad, bias, *labels, lg, loss_sum, *dx, dw, db, st);
  }
  set_error("head1x1_xent: 16 or 32 input channels supported, got %d", x->c);
  return SEG_E_UNSUPPORTED;
}

SEG_API int32_t seg_sigmoid_argmax(const seg_view* logits, float* probs, float* labelmap, void* stream) {
  SEG_REQUIRE(logits && probs && labelmap, SEG_E_BAD_SHAPE, "sigmoid_argmax: null argument");
  const int64_t pixels = (int64_t)logits->n * logits->h * logits->w;
  sigmoid_argmax_kernel<<<grid_for(pixels, 256), 256, 0, (cudaStream_t)stream>>>(*logits, probs,
                                                                                 labelmap);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

SEG_API int32_t seg_classmap_tail_infer(const seg_view* x, int32_t rh, int32_t rw,
                                        const void* w_up_bf16, int32_t up_cout_pad,
                                        int32_t up_cin_pad, const float* b_up,
                                        const float* bn_mean, const float* bn_var, float bn_eps,
                                        const float* bn_beta, const void* w_out_bf16,
                                        int32_t out_cin_pad, int32_t out_cout_pad,
                                        const float* b_out, int32_t n_classes, float* logits,
                                        float* probs, float* labelmap, void* stream) {
  SEG_REQUIRE(x && w_up_bf16 && b_up && bn_mean && bn_var && bn_beta && w_out_bf16 && b_out &&
                  probs && labelmap && rh > 0 && rw > 0,
              SEG_E_BAD_SHAPE, "classmap_tail_infer: null argument");
  SEG_REQUIRE(n_classes >= 2 && n_classes <= 4 && x->c == 32 && up_cin_pad >= 32 &&
                  up_cout_pad >= n_classes && out_cin_pad >= n_classes &&
                  out_cout_pad >= n_classes && vec8_ok(*x),
              SEG_E_UNSUPPORTED, "classmap_tail_infer: 32 input channels (16-byte aligned view), "
              "2..4 classes");
  TailArgs A;
  memset(&A, 0, sizeof(A));
  A.x = *x; A.rh = rh; A.rw = rw;
  A.w_up = reinterpret_cast<const bf16*>(w_up_bf16); A.up_cop = up_cout_pad; A.up_cip = up_cin_pad;
  A.b_up = b_up;
  A.bn_mean = bn_mean; A.bn_var = bn_var; A.bn_eps = bn_eps; A.bn_beta = bn_beta;
  A.w_out = reinterpret_cast<const bf16*>(w_out_bf16); A.out_cip = out_cin_pad;
  A.out_cop = out_cout_pad; A.b_out = b_out;
  A.logits = logits; A.probs = probs; A.labelmap = labelmap;
  const int tiles = ((rw + kTailT - 1) / kTailT) * ((rh + kTailT - 1) / kTailT);
  const dim3 grid((unsigned)(tiles * x->n));
  cudaStream_t st = (cudaStream_t)stream;
  // the 16- and 8-byte stores of the tensor-core form need aligned class maps and 4-byte
  // aligned weight pairs; anything else takes the CUDA-core form
  const bool mma_ok = g_tail_mma && (up_cin_pad % 2) == 0 &&
                      (reinterpret_cast<uintptr_t>(w_up_bf16) & 3) == 0 &&
                      (reinterpret_cast<uintptr_t>(probs) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(logits) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(labelmap) & 7) == 0;
  if (mma_ok) {
    switch (n_classes) {
      case 2: SEG_CHECK_CUDA(launch_k(classmap_tail_mma_kernel<2>, grid, dim3(256), (size_t)0, st, A)); break;
      case 3: SEG_CHECK_CUDA(launch_k(classmap_tail_mma_kernel<3>, grid, dim3(256), (size_t)0, st, A)); break;
      default: SEG_CHECK_CUDA(launch_k(classmap_tail_mma_kernel<4>, grid, dim3(256), (size_t)0, st, A)); break;
    }
    return SEG_OK;
  }
  switch (n_classes) {
    case 2: SEG_CHECK_CUDA(launch_k(classmap_tail_kernel<2, 32>, grid, dim3(256), (size_t)0, st, A)); break;
    case 3: SEG_CHECK_CUDA(launch_k(classmap_tail_kernel<3, 32>, grid, dim3(256), (size_t)0, st, A)); break;
    default: SEG_CHECK_CUDA(launch_k(classmap_tail_kernel<4, 32>, grid, dim3(256), (size_t)0, st, A)); break;
  }
  return SEG_OK;
}

SEG_API int32_t seg_mc_mean_var(const float* probs, int32_t t, int64_t count, float* mean, float* var,
                        void* stream) {
  SEG_REQUIRE(probs && mean && var && t >= 1, SEG_E_BAD_SHAPE, "mc_mean_var: bad argument");
  mc_mean_var_kernel<<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(probs, t, count, mean,
                                                                             var);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

SEG_API int32_t seg_adam_chunk_elems(void) { return kAdamChunk; }

SEG_API int32_t seg_adam_multi(float* param, float* grad, float* m, float* v, void* shadow_bf16,
                       const int32_t* segments, const int64_t* shadow_offsets,
                       const int32_t* chunks, int32_t nchunks, float lr_t,
                       const float* lr_t_dev, float beta1, float beta2, float eps,
                       float grad_scale, void* stream) {
  SEG_REQUIRE(param && grad && m && v && segments && shadow_offsets && chunks && nchunks > 0,
              SEG_E_BAD_SHAPE, "adam_multi: null argument");
  SEG_CHECK_CUDA(launch_k(adam_multi_kernel, dim3((unsigned)nchunks), dim3(kAdamThreads), (size_t)(0), (cudaStream_t)stream, param, grad, m, v, reinterpret_cast<bf16*>(shadow_bf16), segments, shadow_offsets, chunks, lr_t, lr_t_dev, beta1, beta2, eps, grad_scale));
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

SEG_API int32_t seg_pack_input(const float* x, int32_t c, const seg_view* y, void* stream) {
  SEG_REQUIRE(x && y && c <= y->c, SEG_E_BAD_SHAPE, "pack_input: bad argument");
  const int64_t total = (int64_t)y->n * y->h * y->w * y->c;
  if (y->c == 16 && c <= 4 && view_dense(*y) && (reinterpret_cast<uintptr_t>(y->ptr) % 16) == 0) {
    const int64_t pixels = (int64_t)y->n * y->h * y->w;
    SEG_CHECK_CUDA(launch_k(pack_input16_kernel, dim3(grid_for(pixels, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, x, c, reinterpret_cast<bf16*>(y->ptr), pixels));
  } else {
    pack_input_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, c, *y);
  }
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

SEG_API int32_t seg_pack_patches(const float* x, int32_t c, int32_t h, int32_t w, int32_t kh,
                                 int32_t kw, int32_t stride, int32_t pad_t, int32_t pad_l,
                                 const seg_view* y, void* stream) {
  SEG_REQUIRE(x && y && c >= 1 && kh >= 1 && kw >= 1 && stride >= 1, SEG_E_BAD_SHAPE,
              "pack_patches: bad argument");
  SEG_REQUIRE(kh * kw * c <= y->c && vec8_ok(*y), SEG_E_BAD_SHAPE,
              "pack_patches: %d patch values do not fit %d channels (multiple of 8, 16-byte "
              "aligned view)", kh * kw * c, y->c);
  const int64_t pixels = (int64_t)y->n * y->h * y->w;
  const int cp = (kh * kw * c + 15) / 16 * 16;
  if (c == 3 && kh == 3 && kw == 3 && y->c == cp) {
    SEG_CHECK_CUDA(launch_k(pack_patches_fixed_kernel<3, 3, 3>, dim3(grid_for(pixels, 128)), dim3(128),
                            (size_t)0, (cudaStream_t)stream, x, h, w, stride, pad_t, pad_l, *y));
    return SEG_OK;
  }
  if (c == 3 && kh == 5 && kw == 5 && y->c == cp) {
    SEG_CHECK_CUDA(launch_k(pack_patches_fixed_kernel<5, 5, 3>, dim3(grid_for(pixels, 128)), dim3(128),
                            (size_t)0, (cudaStream_t)stream, x, h, w, stride, pad_t, pad_l, *y));
    return SEG_OK;
  }
  const int64_t total = pixels * (y->c / 8);
  SEG_CHECK_CUDA(launch_k(pack_patches_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)0,
                          (cudaStream_t)stream, x, c, h, w, kh, kw, stride, pad_t, pad_l, *y));
  return SEG_OK;
}

SEG_API int32_t seg_stage_input(const void* x, int32_t x_kind, int32_t c, int32_t src_h,
                                int32_t src_w, const int32_t* crop_yx, const seg_view* y4,
                                const uint8_t* mask_src, int32_t mask_kind, uint8_t* mask_dst,
                                const seg_stage_ctl* ctl, void* stream) {
  SEG_REQUIRE(x && y4 && y4->c == 4 && view_dense(*y4) && c >= 1 && c <= 3 &&
                  (x_kind == 0 || x_kind == 1) && src_h >= y4->h && src_w >= y4->w &&
                  (reinterpret_cast<uintptr_t>(y4->ptr) & 7) == 0,
              SEG_E_BAD_SHAPE, "stage_input: need a dense 4-channel bf16 destination, 1..3 source "
              "channels and a source at least as large as the destination");
  SEG_REQUIRE(!mask_src || mask_dst, SEG_E_BAD_SHAPE, "stage_input: mask_dst missing");
  SEG_REQUIRE(crop_yx || (src_h == y4->h && src_w == y4->w), SEG_E_BAD_SHAPE,
              "stage_input: a larger source needs crop offsets");
  StageArgs A;
  memset(&A, 0, sizeof(A));
  A.x = x; A.kind = x_kind; A.C = c; A.src_h = src_h; A.src_w = src_w; A.crop_yx = crop_yx;
  A.y = reinterpret_cast<uint2*>(y4->ptr); A.n = y4->n; A.H = y4->h; A.W = y4->w;
  A.mask_src = mask_src; A.mask_kind = mask_kind; A.mask_dst = mask_dst;
  if (ctl) { A.ctl = *ctl; A.has_ctl = 1; }
  const int64_t pixels = (int64_t)A.n * A.H * A.W;
  A.vec4 = x_kind == 0 && c == 3 && !crop_yx && pixels % 4 == 0 &&
           (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(y4->ptr) & 15) == 0 &&
           (!mask_src || ((reinterpret_cast<uintptr_t>(mask_src) & 3) == 0 &&
                          (reinterpret_cast<uintptr_t>(mask_dst) & 3) == 0));
  SEG_CHECK_CUDA(launch_k(stage_input4_kernel, dim3(grid_for(A.vec4 ? pixels / 4 : pixels, 256)),
                          dim3(256), (size_t)0, (cudaStream_t)stream, A));
  return SEG_OK;
}

SEG_API int32_t seg_fill_zero(void* ptr, int64_t bytes, void* stream) {
  SEG_CHECK_CUDA(cudaMemsetAsync(ptr, 0, (size_t)bytes, (cudaStream_t)stream));
  return SEG_OK;
}

}  // extern "C"
