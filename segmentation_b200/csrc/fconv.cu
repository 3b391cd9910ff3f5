// Host side of the first-layer kernel (fconv.cuh): eligibility and launches.
#include <cstdlib>

#include "fconv.cuh"

namespace segb {

// umma_conv.cu: loads the driver's tensor-map encoder and builds a dense 2-D bf16 map
int make_probe_tmap(CUtensorMap* tm, const void* ptr, int64_t cols, int64_t rows, int box_cols,
                    int box_rows, int swizzle_bytes);

static bool g_use_fconv = true;      // seg_set_option key 16
void fconv_enable(int on) { g_use_fconv = on != 0; }

// x: the (R, G, B, 1) bf16 input of seg_stage_input; 3x3 stride-1 convolution of 3 real
// channels onto 32 or 64 (padded) output channels.  out_h x out_w: the convolution's output
// grid.
static bool fconv_geom_ok(const seg_conv_desc& d, const seg_view& x, const seg_view* x2, int out_n,
                          int out_h, int out_w) {
  const bool k3 = d.kh == 3 && d.kw == 3 && d.stride == 1;
  const bool k5 = d.kh == 5 && d.kw == 5 && d.stride == 2;      // forward variants only
  return g_use_fconv && (k3 || k5) && d.cin == 3 && x.c == 4 &&
         !(x2 && x2->ptr) && x.sw == 4 && x.sh % 4 == 0 && x.sn % 4 == 0 &&
         (d.cout_pad == 32 || d.cout_pad == 64) &&
         (reinterpret_cast<uintptr_t>(x.ptr) & 7) == 0 && out_n == x.n &&
         (int64_t)out_n * out_h * out_w < (int64_t)1 << 30 &&
         out_h == (x.h + d.pad_t + d.pad_b - d.kh) / d.stride + 1 &&
         out_w == (x.w + d.pad_l + d.pad_r - d.kw) / d.stride + 1 && out_w > 1 && out_h > 1;
}
static bool fconv_shape_ok(const seg_conv_desc& d, const seg_view& x, const seg_view* x2,
                           const seg_view& out) {
  return fconv_geom_ok(d, x, x2, out.n, out.h, out.w) && view_dense(out) && out.c == d.cout_pad;
}

template <int BN, bool WGRAD, bool POOL, int KH = 3, int ST = 1>
static int launch_fconv_t(const FconvParams& P, const void* io, cudaStream_t st) {
  using Cfg = FconvCfg<BN, WGRAD, POOL, KH, ST>;
  static bool attr_done = false;
  if (!attr_done) {
    SEG_CHECK_CUDA(cudaFuncSetAttribute(fconv_kernel<BN, WGRAD, POOL, KH, ST>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::kSmemBytes));
    attr_done = true;
  }
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  if (!POOL) {
    // fwd: store boxes of 32 pixel rows (one per epilogue warp); wgrad: whole 128-pixel dZ tiles
    const int rc = make_probe_tmap(&tm, io, BN, P.M_total, BN, WGRAD ? 128 : 32, BN * 2);
    if (rc) return rc;
  }
  // two CTAs per SM when they fit: a tile's patch loads have ~1 us of latency to hide
  const int per_sm = 2 * Cfg::kSmemBytes <= 220 * 1024 ? 2 : 1;
  int grid = per_sm * num_sms();
  if (grid > P.tiles) grid = P.tiles;
  SEG_CHECK_CUDA(launch_k(fconv_kernel<BN, WGRAD, POOL, KH, ST>, dim3(grid), dim3(kFcThreads),
                          (size_t)Cfg::kSmemBytes, st, tm, P));
  return SEG_OK;
}

// exact unsigned division of values < 2^31 by multiply-high + shift (mul = 0: divisor 1)
static void magic(uint32_t d, uint32_t* dv, uint32_t* mul, uint32_t* shr) {
  *dv = d;
  if (d == 1) { *mul = 0; *shr = 0; return; }
  uint32_t l = 0;
  while ((1u << l) < d) ++l;
  const uint32_t p = 31 + l;
  *mul = (uint32_t)((((uint64_t)1 << p) + d - 1) / d);
  *shr = p - 32;
}

static void fill_params(FconvParams* P, const seg_conv_desc& d, const seg_view& x, int out_h,
                        int out_w, bool pool) {
  memset(P, 0, sizeof(*P));
  P->x4 = reinterpret_cast<const uint2*>(x.ptr);
  P->x_sn = x.sn / 4; P->x_sh = x.sh / 4;
  P->H = x.h; P->W = x.w; P->Ho = out_h; P->Wo = out_w;
  P->pad_t = d.pad_t; P->pad_l = d.pad_l;
  P->M_total = x.n * out_h * out_w;
  if (pool) {
    P->Hp = out_h / 2; P->Wp = out_w / 2;
    const int tiles_x = (out_w + 63) / 64;
    P->tiles = x.n * P->Hp * tiles_x;
    magic((uint32_t)tiles_x, &P->div_a, &P->div_a_mul, &P->div_a_shr);
    magic((uint32_t)(P->Hp * tiles_x), &P->div_b, &P->div_b_mul, &P->div_b_shr);
  } else {
    P->tiles = (P->M_total + 127) / 128;
    magic((uint32_t)out_w, &P->div_a, &P->div_a_mul, &P->div_a_shr);
    magic((uint32_t)(out_h * out_w), &P->div_b, &P->div_b_mul, &P->div_b_shr);
  }
  P->cin_pad = d.cin_pad; P->cout_pad = d.cout_pad; P->cout = d.cout;
}

// SEG_E_UNSUPPORTED (nothing launched): not a first-layer shape
int fconv_fwd(const seg_conv_desc& d, const seg_view& x, const seg_view* x2, const void* w,
              const float* bias, const seg_view& y, cudaStream_t st) {
  if (!fconv_shape_ok(d, x, x2, y) || (d.flags & SEG_EPI_OUT_F32)) return SEG_E_UNSUPPORTED;
  FconvParams P;
  fill_params(&P, d, x, y.h, y.w, false);
  P.w = reinterpret_cast<const bf16*>(w);
  P.bias = bias;
  P.flags = d.flags;
  if (d.kh == 5)
    return d.cout_pad == 32 ? launch_fconv_t<32, false, false, 5, 2>(P, y.ptr, st)
                            : SEG_E_UNSUPPORTED;
  return d.cout_pad == 32 ? launch_fconv_t<32, false, false>(P, y.ptr, st)
                          : launch_fconv_t<64, false, false>(P, y.ptr, st);
}

int fconv_wgrad(const seg_conv_desc& d, const seg_view& x, const seg_view* x2,
                const seg_view& dz, float* dw, float* db, cudaStream_t st) {
  if (!fconv_shape_ok(d, x, x2, dz) || d.kh != 3) return SEG_E_UNSUPPORTED;
  FconvParams P;
  fill_params(&P, d, x, dz.h, dz.w, false);
  P.dw = dw;
  P.db = db;
  return d.cout_pad == 32 ? launch_fconv_t<32, true, false>(P, dz.ptr, st)
                          : launch_fconv_t<64, true, false>(P, dz.ptr, st);
}

// ---- fused with the 2x2 / stride-2 max-pool that follows the layer -------------------------
static bool pool_geom_ok(const seg_conv_desc& d, const seg_view& x, int oh, int ow,
                         const seg_view& pooled) {
  return fconv_geom_ok(d, x, nullptr, x.n, oh, ow) && d.cout_pad == 32 && oh % 2 == 0 &&
         ow % 2 == 0 && view_dense(pooled) && pooled.c == 32 && pooled.n == x.n &&
         pooled.h == oh / 2 && pooled.w == ow / 2 &&
         (reinterpret_cast<uintptr_t>(pooled.ptr) & 15) == 0;
}

static bool window_view_ok(const seg_view& v) {     // 16-byte chunks of 32-channel pixels
  return v.c == 32 && (reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0 && v.sw % 8 == 0 &&
         v.sh % 8 == 0 && v.sn % 8 == 0;
}

int fconv_pool_fwd(const seg_conv_desc& d, const seg_view& x, const void* w, const float* bias,
                   const seg_view* y_win, int win_y0, int win_x0, const seg_view& pooled,
                   uint8_t* argmax, const float* bn_mean, const float* bn_var, float bn_eps,
                   const float* bn_beta, int w_rows_per_tap, cudaStream_t st) {
  const int oh = (x.h + d.pad_t + d.pad_b - d.kh) / d.stride + 1;
  const int ow = (x.w + d.pad_l + d.pad_r - d.kw) / d.stride + 1;
  if (!pool_geom_ok(d, x, oh, ow, pooled) || (d.flags & SEG_EPI_OUT_F32) ||
      (reinterpret_cast<uintptr_t>(argmax) & 7) != 0)
    return SEG_E_UNSUPPORTED;
  FconvParams P;
  fill_params(&P, d, x, oh, ow, true);
  P.w = reinterpret_cast<const bf16*>(w);
  P.bias = bias;
  P.flags = d.flags;
  P.pooled = reinterpret_cast<bf16*>(pooled.ptr);
  P.amax = argmax;
  if (y_win && y_win->ptr) {
    if (!window_view_ok(*y_win) || y_win->n != x.n || win_y0 < 0 || win_x0 < 0 ||
        win_y0 + y_win->h > oh || win_x0 + y_win->w > ow)
      return SEG_E_UNSUPPORTED;
    // the kernel addresses the window through full-grid coordinates
    P.y = reinterpret_cast<bf16*>(y_win->ptr) - (int64_t)win_y0 * y_win->sh -
          (int64_t)win_x0 * y_win->sw;
    P.y_sn = y_win->sn; P.y_sh = y_win->sh; P.y_sw = y_win->sw;
    P.win_y0 = win_y0; P.win_x0 = win_x0;
    P.win_y1 = win_y0 + y_win->h; P.win_x1 = win_x0 + y_win->w;
  }
  P.bn_mean = bn_mean; P.bn_var = bn_var; P.bn_beta = bn_beta; P.bn_eps = bn_eps;
  if (w_rows_per_tap > 0) P.cin_pad = w_rows_per_tap;      // weight-matrix rows per filter tap
  if (d.kh == 5) return launch_fconv_t<32, false, true, 5, 2>(P, nullptr, st);
  return launch_fconv_t<32, false, true>(P, nullptr, st);
}

int fconv_pool_wgrad(const seg_conv_desc& d, const seg_view& x, const seg_view& dpool,
                     const uint8_t* argmax, const seg_view& pooled, float* dw, float* db,
                     cudaStream_t st) {
  const int oh = x.h + d.pad_t + d.pad_b - 2, ow = x.w + d.pad_l + d.pad_r - 2;
  if (d.kh != 3 || !pool_geom_ok(d, x, oh, ow, pooled) || !pool_geom_ok(d, x, oh, ow, dpool) || !argmax ||
      (reinterpret_cast<uintptr_t>(argmax) & 7) != 0)
    return SEG_E_UNSUPPORTED;
  FconvParams P;
  fill_params(&P, d, x, oh, ow, true);
  P.dw = dw;
  P.db = db;
  P.dpool = reinterpret_cast<const bf16*>(dpool.ptr);
  P.pooled = reinterpret_cast<bf16*>(pooled.ptr);
  P.amax = const_cast<uint8_t*>(argmax);
  return launch_fconv_t<32, true, true>(P, nullptr, st);
}

}  // namespace segb
