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
// channels onto 32 or 64 (padded) output channels, every tensor dense
static bool fconv_shape_ok(const seg_conv_desc& d, const seg_view& x, const seg_view* x2,
                           const seg_view& out) {
  return g_use_fconv && d.kh == 3 && d.kw == 3 && d.stride == 1 && d.cin == 3 && x.c == 4 &&
         !(x2 && x2->ptr) && view_dense(x) && view_dense(out) && out.c == d.cout_pad &&
         (d.cout_pad == 32 || d.cout_pad == 64) &&
         (reinterpret_cast<uintptr_t>(x.ptr) & 7) == 0 &&
         (int64_t)out.n * out.h * out.w < (int64_t)1 << 30 &&
         out.h == x.h + d.pad_t + d.pad_b - 2 && out.w == x.w + d.pad_l + d.pad_r - 2 &&
         out.w > 1 && out.h * out.w > 1;
}

template <int BN, bool WGRAD>
static int launch_fconv_t(const FconvParams& P, const void* io, cudaStream_t st) {
  using Cfg = FconvCfg<BN, WGRAD>;
  static bool attr_done = false;
  if (!attr_done) {
    SEG_CHECK_CUDA(cudaFuncSetAttribute(fconv_kernel<BN, WGRAD>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::kSmemBytes));
    attr_done = true;
  }
  CUtensorMap tm;
  // fwd: store boxes of 32 pixel rows (one per epilogue warp); wgrad: whole 128-pixel dZ tiles
  const int rc = make_probe_tmap(&tm, io, BN, P.M_total, BN, WGRAD ? 128 : 32, BN * 2);
  if (rc) return rc;
  // two CTAs per SM when they fit: a tile's patch loads have ~1 us of latency to hide
  const int per_sm = 2 * Cfg::kSmemBytes <= 220 * 1024 ? 2 : 1;
  int grid = per_sm * num_sms();
  if (grid > P.tiles) grid = P.tiles;
  SEG_CHECK_CUDA(launch_k(fconv_kernel<BN, WGRAD>, dim3(grid), dim3(kFcThreads),
                          (size_t)Cfg::kSmemBytes, st, tm, P));
  return SEG_OK;
}

static void fill_params(FconvParams* P, const seg_conv_desc& d, const seg_view& x,
                        const seg_view& out) {
  memset(P, 0, sizeof(*P));
  P->x4 = reinterpret_cast<const uint2*>(x.ptr);
  P->H = x.h; P->W = x.w; P->Ho = out.h; P->Wo = out.w;
  P->pad_t = d.pad_t; P->pad_l = d.pad_l;
  P->M_total = out.n * out.h * out.w;
  P->tiles = (P->M_total + 127) / 128;
  // exact unsigned division of values < 2^31 by multiply-high + shift
  auto magic = [](uint32_t d, uint32_t* mul, uint32_t* shr) {
    if (d == 1) { *mul = 0xFFFFFFFFu; *shr = 0; return; }      // handled below: x*1
    uint32_t l = 0;
    while ((1u << l) < d) ++l;
    const uint32_t p = 31 + l;
    *mul = (uint32_t)((((uint64_t)1 << p) + d - 1) / d);
    *shr = p - 32;
  };
  magic((uint32_t)out.w, &P->div_wo_mul, &P->div_wo_shr);
  magic((uint32_t)(out.h * out.w), &P->div_hw_mul, &P->div_hw_shr);
  P->cin_pad = d.cin_pad; P->cout_pad = d.cout_pad; P->cout = d.cout;
}

// SEG_E_UNSUPPORTED (nothing launched): not a first-layer shape
int fconv_fwd(const seg_conv_desc& d, const seg_view& x, const seg_view* x2, const void* w,
              const float* bias, const seg_view& y, cudaStream_t st) {
  if (!fconv_shape_ok(d, x, x2, y) || (d.flags & SEG_EPI_OUT_F32)) return SEG_E_UNSUPPORTED;
  FconvParams P;
  fill_params(&P, d, x, y);
  P.w = reinterpret_cast<const bf16*>(w);
  P.bias = bias;
  P.flags = d.flags;
  return d.cout_pad == 32 ? launch_fconv_t<32, false>(P, y.ptr, st)
                          : launch_fconv_t<64, false>(P, y.ptr, st);
}

int fconv_wgrad(const seg_conv_desc& d, const seg_view& x, const seg_view* x2,
                const seg_view& dz, float* dw, float* db, cudaStream_t st) {
  if (!fconv_shape_ok(d, x, x2, dz)) return SEG_E_UNSUPPORTED;
  FconvParams P;
  fill_params(&P, d, x, dz);
  P.dw = dw;
  P.db = db;
  return d.cout_pad == 32 ? launch_fconv_t<32, true>(P, dz.ptr, st)
                          : launch_fconv_t<64, true>(P, dz.ptr, st);
}

}  // namespace segb
