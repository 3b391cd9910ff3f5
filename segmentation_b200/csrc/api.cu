// C ABI: library/device entry points and the convolution family (dispatch between
// the tcgen05 implicit-GEMM path and the CUDA-core path selected by desc->impl).
#include <stdarg.h>
#include <string.h>

#include "simt_conv.cuh"

namespace segb {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// umma_conv.cu
int umma_conv_fwd(const seg_conv_desc& d, const seg_view& x, const seg_view* x2, const void* w,
                  const float* bias, const seg_view& y, cudaStream_t st,
                  const float* post_scale = nullptr, const float* post_shift = nullptr);
int umma_conv_dgrad(const seg_conv_desc& d, const seg_view& dz, const void* w, const seg_view& dx,
                    const seg_view* dx2, const seg_view* mask, const seg_view* mask2,
                    cudaStream_t st, int cin_lo);
int umma_conv_wgrad(const seg_conv_desc& d, const seg_view& x, const seg_view* x2,
                    const seg_view& dz, float* dw, float* db, cudaStream_t st);
int umma_deconv_fwd(const seg_conv_desc& d, const seg_view& x, const void* w, const float* bias,
                    const seg_view& y, cudaStream_t st, const float* post_scale = nullptr,
                    const float* post_shift = nullptr);
int umma_deconv_dgrad(const seg_conv_desc& d, const seg_view& dz, const void* w,
                      const seg_view& dx, const seg_view* mask, cudaStream_t st);
int umma_deconv_wgrad(const seg_conv_desc& d, const seg_view& x, const seg_view& dz, float* dw,
                      cudaStream_t st);
// fconv.cu
int fconv_pool_fwd(const seg_conv_desc& d, const seg_view& x, const void* w, const float* bias,
                   const seg_view* y_win, int win_y0, int win_x0, const seg_view& pooled,
                   uint8_t* argmax, const float* bn_mean, const float* bn_var, float bn_eps,
                   const float* bn_beta, int w_rows_per_tap, cudaStream_t st);
int fconv_pool_wgrad(const seg_conv_desc& d, const seg_view& x, const seg_view& dpool,
                     const uint8_t* argmax, const seg_view& pooled, float* dw, float* db,
                     cudaStream_t st);


void hconv_set_row_align(int a);
void pool_set_rows(int on);
void tail_set_mma(int on);
void conv_set_deep_b_ring(int on);
void hconv_set_rowstage(int on);
void twgrad_set_tred(int on);
void fconv_enable(int on);
void hconv_set_prof(void* p);
void hconv_enable(int on);
void tconv_enable(int on);
void tconv_set_min_eff(int pct);
void twgrad_enable(int on);
void twgrad_set_min_eff(int pct);
void twgrad_set_min_tiles(int n);

static bool desc_ok(const seg_conv_desc* d) {
  return d && d->kh >= 1 && d->kw >= 1 && d->stride >= 1 && d->cin >= 1 && d->cout >= 1 &&
         d->cin_pad >= d->cin && d->cout_pad >= d->cout && d->cin_pad % 8 == 0 &&
         d->cout_pad % 8 == 0;
}

#ifndef SEGB200_KERNEL_PROF
#define SEGB200_KERNEL_PROF 0
#endif
int g_pdl = 1;              // seg_set_option key 7
int g_sm_limit = 0;         // seg_set_option key 17
thread_local const void* g_last_kernel_fn = nullptr;

}  // namespace segb

using namespace segb;

extern "C" {

SEG_API int32_t seg_version(void) { return 100; }

SEG_API const char* seg_last_error_string(void) { return g_err; }

SEG_API const char* seg_last_kernel_name(void) {
  const char* name = nullptr;
  if (g_last_kernel_fn == nullptr || cudaFuncGetName(&name, g_last_kernel_fn) != cudaSuccess ||
      name == nullptr) {
    cudaGetLastError();
    return "";
  }
  return name;
}

#if SEGB200_KERNEL_PROF
// profiling builds only (-DSEGB200_KERNEL_PROF=1, tools/layer_prof.py); see segb200_probes.h
SEG_API int32_t seg_debug_prof_buffer(void* device_buf) {
  hconv_set_prof(device_buf);
  return SEG_OK;
}
#endif

SEG_API int32_t seg_set_option(int32_t key, int32_t value) {
  switch (key) {
    case 1: hconv_enable(value); return SEG_OK;
    case 2: hconv_set_row_align(value); return SEG_OK;
    case 3: tconv_enable(value); return SEG_OK;
    case 4: tconv_set_min_eff(value); return SEG_OK;
    case 5: twgrad_enable(value); return SEG_OK;
    case 6: twgrad_set_min_eff(value); return SEG_OK;
    case 7: g_pdl = value != 0; return SEG_OK;
    case 9: twgrad_set_min_tiles(value); return SEG_OK;
    case 11: pool_set_rows(value); return SEG_OK;
    case 12: conv_set_deep_b_ring(value); return SEG_OK;
    case 14: hconv_set_rowstage(value); return SEG_OK;
    case 15: twgrad_set_tred(value); return SEG_OK;
    case 16: fconv_enable(value); return SEG_OK;
    case 17: g_sm_limit = value > 0 ? value : 0; return SEG_OK;
    case 18: tail_set_mma(value); return SEG_OK;
  }
  set_error("seg_set_option: unknown key %d", key);
  return SEG_E_BAD_SHAPE;
}

SEG_API int32_t seg_device_check(void) {
  int dev = 0;
  SEG_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  SEG_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  SEG_REQUIRE(prop.major == 10 && prop.minor == 0, SEG_E_ARCH,
              "segb200 needs compute capability 10.0 (sm_100a); device %d is %d.%d (%s)", dev,
              prop.major, prop.minor, prop.name);
  return SEG_OK;
}

SEG_API int32_t seg_conv2d_fwd(const seg_conv_desc* d, const seg_view* x, const seg_view* x2,
                       const void* w_bf16, const float* bias, const seg_view* y, void* stream) {
  SEG_REQUIRE(desc_ok(d) && x && w_bf16 && y, SEG_E_BAD_SHAPE, "conv2d_fwd: bad argument");
  SEG_REQUIRE(y->h == (x->h + d->pad_t + d->pad_b - d->kh) / d->stride + 1 &&
                  y->w == (x->w + d->pad_l + d->pad_r - d->kw) / d->stride + 1 && y->n == x->n,
              SEG_E_BAD_SHAPE, "conv2d_fwd: output geometry mismatch (%dx%d)", y->h, y->w);
  SEG_REQUIRE(!(d->flags & SEG_EPI_BIAS) || bias, SEG_E_BAD_SHAPE, "conv2d_fwd: bias missing");
  cudaStream_t st = (cudaStream_t)stream;
  DirectParams P;
  memset(&P, 0, sizeof(P));
  P.x = *x;
  P.x2 = x2 ? *x2 : null_view();
  P.w = reinterpret_cast<const bf16*>(w_bf16);
  P.bias = bias;
  P.y = *y;
  P.kh = d->kh; P.kw = d->kw; P.stride = d->stride; P.pad_t = d->pad_t; P.pad_l = d->pad_l;
  P.in_pad = d->cin_pad; P.out_pad = d->cout_pad;
  P.flags = d->flags & (SEG_EPI_BIAS | SEG_EPI_RELU | SEG_EPI_OUT_F32);
  // class-map heads (<= 4 channels in and out): a streaming kernel for either impl
  if (simt_tiny_conv_ok(P, d->cin)) return simt_tiny_conv(P, d->cin, st);
  if (d->impl == SEG_IMPL_UMMA) return umma_conv_fwd(*d, *x, x2, w_bf16, bias, *y, st);
  return simt_direct(P, st);
}

SEG_API int32_t seg_conv2d_fwd_affine(const seg_conv_desc* d, const seg_view* x, const void* w_bf16,
                                      const float* bias, const float* post_scale,
                                      const float* post_shift, const seg_view* y, void* stream) {
  SEG_REQUIRE(desc_ok(d) && x && w_bf16 && y && post_scale && post_shift, SEG_E_BAD_SHAPE,
              "conv2d_fwd_affine: bad argument");
  SEG_REQUIRE(y->h == (x->h + d->pad_t + d->pad_b - d->kh) / d->stride + 1 &&
                  y->w == (x->w + d->pad_l + d->pad_r - d->kw) / d->stride + 1 && y->n == x->n,
              SEG_E_BAD_SHAPE, "conv2d_fwd_affine: output geometry mismatch (%dx%d)", y->h, y->w);
  SEG_REQUIRE(!(d->flags & SEG_EPI_BIAS) || bias, SEG_E_BAD_SHAPE,
              "conv2d_fwd_affine: bias missing");
  SEG_REQUIRE(d->impl == SEG_IMPL_UMMA && !(d->flags & SEG_EPI_OUT_F32), SEG_E_UNSUPPORTED,
              "conv2d_fwd_affine: tcgen05 path with a bf16 output only");
  return umma_conv_fwd(*d, *x, nullptr, w_bf16, bias, *y, (cudaStream_t)stream, post_scale,
                       post_shift);
}

SEG_API int32_t seg_deconv2d_fwd_affine(const seg_conv_desc* d, const seg_view* x,
                                        const void* w_bf16, const float* bias,
                                        const float* post_scale, const float* post_shift,
                                        const seg_view* y, void* stream) {
  SEG_REQUIRE(desc_ok(d) && x && w_bf16 && y && post_scale && post_shift, SEG_E_BAD_SHAPE,
              "deconv2d_fwd_affine: bad argument");
  SEG_REQUIRE(!(d->flags & SEG_EPI_BIAS) || bias, SEG_E_BAD_SHAPE,
              "deconv2d_fwd_affine: bias missing");
  SEG_REQUIRE(d->impl == SEG_IMPL_UMMA && !(d->flags & SEG_EPI_OUT_F32), SEG_E_UNSUPPORTED,
              "deconv2d_fwd_affine: tcgen05 path with a bf16 output only");
  return umma_deconv_fwd(*d, *x, w_bf16, bias, *y, (cudaStream_t)stream, post_scale, post_shift);
}

SEG_API int32_t seg_conv2d_dgrad(const seg_conv_desc* d, const seg_view* dz, const void* w_bf16,
                         const seg_view* dx, const seg_view* dx2, const seg_view* mask_src,
                         const seg_view* mask_src2, void* stream) {
  SEG_REQUIRE(desc_ok(d) && dz && w_bf16 && dx, SEG_E_BAD_SHAPE, "conv2d_dgrad: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  seg_conv_desc dd = *d;
  dd.flags = (mask_src || mask_src2) ? SEG_EPI_RELU_MASK : 0;
  if (d->impl == SEG_IMPL_UMMA) {
    SEG_REQUIRE(dx->c + ((dx2 && dx2->ptr) ? dx2->c : 0) == d->cin_pad, SEG_E_BAD_SHAPE,
                "conv2d_dgrad: dx channels mismatch");
    return umma_conv_dgrad(dd, *dz, w_bf16, *dx, dx2, mask_src, mask_src2, st, 0);
  }
  TransParams P;
  memset(&P, 0, sizeof(P));
  P.src = *dz;
  P.src.c = d->cout;          // padded dz channels hold zeros; skip them
  P.w = reinterpret_cast<const bf16*>(w_bf16);
  P.out = *dx;
  P.out2 = dx2 ? *dx2 : null_view();
  P.mask = mask_src ? *mask_src : null_view();
  P.mask2 = mask_src2 ? *mask_src2 : null_view();
  P.kh = d->kh; P.kw = d->kw; P.stride = d->stride; P.pad_t = d->pad_t; P.pad_l = d->pad_l;
  P.oc_pad = d->cin_pad; P.ic_pad = d->cout_pad;
  P.flags = dd.flags;
  return simt_transposed(P, st);
}

SEG_API int32_t seg_conv2d_dgrad_slice(const seg_conv_desc* d, const seg_view* dz,
                                       const void* w_bf16, int32_t cin_lo, const seg_view* dx,
                                       const seg_view* mask_src, void* stream) {
  SEG_REQUIRE(desc_ok(d) && dz && w_bf16 && dx, SEG_E_BAD_SHAPE, "conv2d_dgrad_slice: bad argument");
  SEG_REQUIRE(d->impl == SEG_IMPL_UMMA, SEG_E_UNSUPPORTED, "conv2d_dgrad_slice: tcgen05 path only");
  seg_conv_desc dd = *d;
  dd.flags = mask_src ? SEG_EPI_RELU_MASK : 0;
  return umma_conv_dgrad(dd, *dz, w_bf16, *dx, nullptr, mask_src, nullptr, (cudaStream_t)stream,
                         cin_lo);
}

SEG_API int32_t seg_conv2d_wgrad(const seg_conv_desc* d, const seg_view* x, const seg_view* x2,
                         const seg_view* dz, float* dw, float* db, void* stream) {
  SEG_REQUIRE(desc_ok(d) && x && dz && dw, SEG_E_BAD_SHAPE, "conv2d_wgrad: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->impl == SEG_IMPL_UMMA) return umma_conv_wgrad(*d, *x, x2, *dz, dw, db, st);
  if (db) {
    seg_view dzb = *dz;
    dzb.c = d->cout;
    int rc = seg_bias_grad(&dzb, db, stream);
    if (rc) return rc;
  }
  WgradParams P;
  memset(&P, 0, sizeof(P));
  P.big = *x;
  P.big2 = x2 ? *x2 : null_view();
  P.small_ = *dz;
  P.dw = dw;
  P.kh = d->kh; P.kw = d->kw; P.stride = d->stride; P.pad_t = d->pad_t; P.pad_l = d->pad_l;
  P.BC = d->cin; P.SC = d->cout;
  return simt_wgrad(P, st);
}

SEG_API int32_t seg_conv2d_pool_fwd(const seg_conv_desc* d, const seg_view* x4, const void* w_bf16,
                                    const float* bias, const seg_view* y_win, int32_t win_y0,
                                    int32_t win_x0, const seg_view* pooled, uint8_t* argmax,
                                    void* stream) {
  SEG_REQUIRE(desc_ok(d) && x4 && w_bf16 && pooled, SEG_E_BAD_SHAPE,
              "conv2d_pool_fwd: bad argument");
  SEG_REQUIRE(!(d->flags & SEG_EPI_BIAS) || bias, SEG_E_BAD_SHAPE, "conv2d_pool_fwd: bias missing");
  const int rc = fconv_pool_fwd(*d, *x4, w_bf16, bias, y_win, win_y0, win_x0, *pooled, argmax,
                                nullptr, nullptr, 0.f, nullptr, 0, (cudaStream_t)stream);
  SEG_REQUIRE(rc != SEG_E_UNSUPPORTED, SEG_E_UNSUPPORTED,
              "conv2d_pool_fwd: needs the first-layer shape (3x3 stride 1 or 5x5 stride 2 on the "
              "(R,G,B,1) input, 32 padded output channels), an even output grid and dense 16-byte "
              "aligned pooled / argmax tensors; use seg_conv2d_fwd + seg_maxpool_fwd otherwise");
  return rc;
}

SEG_API int32_t seg_conv2d_bn_pool_infer(const seg_conv_desc* d, const seg_view* x4,
                                         const void* w_bf16, int32_t w_rows_per_tap,
                                         const float* bias, const float* bn_mean, const float* bn_var, float bn_eps,
                                         const float* bn_beta, const seg_view* pooled,
                                         uint8_t* argmax, void* stream) {
  SEG_REQUIRE(desc_ok(d) && x4 && w_bf16 && pooled && bn_mean && bn_var && bn_beta, SEG_E_BAD_SHAPE,
              "conv2d_bn_pool_infer: bad argument");
  SEG_REQUIRE(!(d->flags & SEG_EPI_BIAS) || bias, SEG_E_BAD_SHAPE,
              "conv2d_bn_pool_infer: bias missing");
  const int rc = fconv_pool_fwd(*d, *x4, w_bf16, bias, nullptr, 0, 0, *pooled, argmax, bn_mean,
                                bn_var, bn_eps, bn_beta, w_rows_per_tap, (cudaStream_t)stream);
  SEG_REQUIRE(rc != SEG_E_UNSUPPORTED, SEG_E_UNSUPPORTED,
              "conv2d_bn_pool_infer: needs the shape of seg_conv2d_pool_fwd");
  return rc;
}

SEG_API int32_t seg_conv2d_pool_wgrad(const seg_conv_desc* d, const seg_view* x4,
                                      const seg_view* dpool, const uint8_t* argmax,
                                      const seg_view* pooled, float* dw, float* db,
                                      void* stream) {
  SEG_REQUIRE(desc_ok(d) && x4 && dpool && argmax && pooled && dw, SEG_E_BAD_SHAPE,
              "conv2d_pool_wgrad: bad argument");
  const int rc = fconv_pool_wgrad(*d, *x4, *dpool, argmax, *pooled, dw, db, (cudaStream_t)stream);
  SEG_REQUIRE(rc != SEG_E_UNSUPPORTED, SEG_E_UNSUPPORTED,
              "conv2d_pool_wgrad: needs the shape of seg_conv2d_pool_fwd; use seg_maxpool_bwd_y + "
              "seg_conv2d_wgrad otherwise");
  return rc;
}

SEG_API int32_t seg_deconv2d_fwd(const seg_conv_desc* d, const seg_view* x, const void* w_bf16,
                         const float* bias, const seg_view* y, void* stream) {
  SEG_REQUIRE(desc_ok(d) && x && w_bf16 && y, SEG_E_BAD_SHAPE, "deconv2d_fwd: bad argument");
  SEG_REQUIRE(!(d->flags & SEG_EPI_BIAS) || bias, SEG_E_BAD_SHAPE, "deconv2d_fwd: bias missing");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->impl == SEG_IMPL_UMMA) {
    const int rc = umma_deconv_fwd(*d, *x, w_bf16, bias, *y, st);
    // k > stride runs as stride^2 halo-tile launches; a geometry those cannot stage (padded
    // row wider than 256 pixels) is handed to the CUDA-core gather kernel below, which
    // rewrites the whole output
    if (!(rc == SEG_E_UNSUPPORTED && d->kh > d->stride)) return rc;
  }
  TransParams P;
  memset(&P, 0, sizeof(P));
  P.src = *x;
  P.src.c = d->cin;
  P.w = reinterpret_cast<const bf16*>(w_bf16);
  P.bias = bias;
  P.out = *y;
  P.out2 = null_view();
  P.mask = null_view();
  P.mask2 = null_view();
  P.kh = d->kh; P.kw = d->kw; P.stride = d->stride; P.pad_t = d->pad_t; P.pad_l = d->pad_l;
  P.oc_pad = d->cout_pad; P.ic_pad = d->cin_pad;
  P.flags = d->flags & (SEG_EPI_BIAS | SEG_EPI_RELU | SEG_EPI_OUT_F32);
  return simt_transposed(P, st);
}

SEG_API int32_t seg_deconv2d_dgrad(const seg_conv_desc* d, const seg_view* dz, const void* w_bf16,
                           const seg_view* dx, const seg_view* mask_src, void* stream) {
  SEG_REQUIRE(desc_ok(d) && dz && w_bf16 && dx, SEG_E_BAD_SHAPE, "deconv2d_dgrad: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  seg_conv_desc dd = *d;
  dd.flags = mask_src ? SEG_EPI_RELU_MASK : 0;
  if (d->impl == SEG_IMPL_UMMA) return umma_deconv_dgrad(dd, *dz, w_bf16, *dx, mask_src, st);
  DirectParams P;
  memset(&P, 0, sizeof(P));
  P.x = *dz;
  P.x.c = d->cout;
  P.x2 = null_view();
  P.w = reinterpret_cast<const bf16*>(w_bf16);
  P.y = *dx;
  P.mask = mask_src ? *mask_src : null_view();
  P.kh = d->kh; P.kw = d->kw; P.stride = d->stride; P.pad_t = d->pad_t; P.pad_l = d->pad_l;
  P.in_pad = d->cout_pad; P.out_pad = d->cin_pad;
  P.flags = dd.flags;
  return simt_direct(P, st);
}

SEG_API int32_t seg_deconv2d_wgrad(const seg_conv_desc* d, const seg_view* x, const seg_view* dz,
                           float* dw, void* stream) {
  SEG_REQUIRE(desc_ok(d) && x && dz && dw, SEG_E_BAD_SHAPE, "deconv2d_wgrad: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->impl == SEG_IMPL_UMMA) return umma_deconv_wgrad(*d, *x, *dz, dw, st);
  WgradParams P;
  memset(&P, 0, sizeof(P));
  P.big = *dz;
  P.big2 = null_view();
  P.small_ = *x;
  P.dw = dw;
  P.kh = d->kh; P.kw = d->kw; P.stride = d->stride; P.pad_t = d->pad_t; P.pad_l = d->pad_l;
  P.BC = d->cout; P.SC = d->cin;
  return simt_wgrad(P, st);
}

}  // extern "C"
