/* segb200.h — C ABI of the B200-native segmentation hot path.
 *
 * The reference (nathanin/segmentation) has no FFI of its own: its hot path is
 * the set of TensorFlow-1.x ops that `sess.run(self.train_op_list)`
 * (/root/reference/models/basemodel.py:484) and `sess.run(self.inference_ops)`
 * (/root/reference/models/basemodel.py:529) execute for the graphs built by
 * models/unet.py:109-175, models/fcn.py:93-220, models/deconvolution.py:101-178.
 * Each entry point below replaces one of those TF ops (cited per function) with a
 * hand-written sm_100a kernel.  See INTEGRATION.md for the binding a maintainer
 * of the reference would add.
 *
 * Conventions
 *   - plain C, no C++ types, no exceptions; every call returns int32 status
 *     (0 = ok, <0 = seg_status); seg_last_error_string() is thread-local.
 *   - every pointer is a DEVICE pointer unless named host_*; the caller owns all
 *     buffers; the library allocates nothing on the hot path.
 *   - every entry takes `stream` (a cudaStream_t passed as void*), enqueues and
 *     returns: stream-ordered, no hidden synchronisation, CUDA-graph capturable.
 *   - activations are NHWC bf16 described by seg_view (channel stride 1),
 *     logits / loss / gradients of parameters are fp32, weights are fp32 masters
 *     in TF layouts (HWIO conv, HWOI transposed conv) plus bf16 shadows in the
 *     same layout with channels padded (see seg_adam_multi).
 */
#ifndef SEGB200_H_
#define SEGB200_H_

#include <stdint.h>

#if defined(__GNUC__)
#define SEG_API __attribute__((visibility("default")))
#else
#define SEG_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum seg_status {
  SEG_OK = 0,
  SEG_E_BAD_SHAPE = -1,
  SEG_E_ALIGN = -2,
  SEG_E_WORKSPACE = -3,
  SEG_E_ARCH = -4,
  SEG_E_CUDA = -5,
  SEG_E_UNSUPPORTED = -6
} seg_status;

/* NHWC view: element (n,y,x,c) lives at ptr[n*sn + y*sh + x*sw + c].  Crops and
 * channel slices of a larger tensor are expressed by ptr offset + strides, which
 * is how tf.image.resize_image_with_crop_or_pad (models/unet.py:140) and
 * tf.concat (models/unet.py:141) cost no kernel here. */
typedef struct seg_view {
  void* ptr;
  int32_t n, h, w, c;
  int64_t sn, sh, sw;
} seg_view;

enum { SEG_IMPL_UMMA = 0, SEG_IMPL_SIMT = 1 };

enum {
  SEG_EPI_BIAS = 1,       /* + bias[co]                                  */
  SEG_EPI_RELU = 2,       /* max(.,0)  (slim default activation_fn)      */
  SEG_EPI_OUT_F32 = 4,    /* y is fp32 (logits) instead of bf16          */
  SEG_EPI_RELU_MASK = 8   /* dgrad only: zero where mask_src <= 0        */
};

/* Convolution geometry shared by fwd / dgrad / wgrad.  x may be the virtual
 * concatenation [x | x2] along channels (x2 nullable): Cin = x.c + x2.c. */
typedef struct seg_conv_desc {
  int32_t kh, kw, stride;
  int32_t pad_t, pad_l, pad_b, pad_r; /* explicit TF-SAME / VALID padding */
  int32_t cin, cout;                  /* logical (unpadded) channel counts */
  int32_t cin_pad, cout_pad;          /* channel counts of the bf16 shadow */
  int32_t flags;                      /* SEG_EPI_*                         */
  int32_t impl;                       /* SEG_IMPL_*                        */
} seg_conv_desc;

/* ---- library / device -------------------------------------------------- */
SEG_API int32_t seg_version(void);
/* 0 iff the current device is compute capability 10.0 (sm_100a kernels). */
SEG_API int32_t seg_device_check(void);
SEG_API const char* seg_last_error_string(void);
/* Mangled name of the tile kernel the calling thread's most recent conv-family / launch_k
 * call was planned onto ("" if none): the plan (tconv / hconv / igemm / twgrad / wgrad
 * instantiation) is chosen from the shape, and profiles attribute time per kernel family. */
SEG_API const char* seg_last_kernel_name(void);
/* Plan-selection switches for tests and A/B measurements (process-wide; set them before
 * the first launch, not concurrently with launches - the hot path itself never calls this).
 * Every default is the measured-best plan; production code does not need this entry point.
 *   key 1: halo-tile tcgen05 conv kernel where it applies (default 1; 0 forces TMA-im2col).
 *   key 2: smem row alignment of the halo kernel's row staging in pixels (0 = natural).
 *   key 3: spatial-tile tcgen05 conv kernel for 3x3 stride-1 fwd / dgrad (default 1).
 *   key 4: minimum useful-pixel percentage of its 16 x 8/16 tiles (default 65).
 *   key 5 / key 6: the same two switches for the spatial-tile weight-gradient kernel
 *          (defaults 1 and 40).
 *   key 7: programmatic dependent launch of the hot-path kernels (default 1); every such
 *          kernel executes griddepcontrol.wait before its first global-memory access.
 *   key 9: minimum number of 8x16 pixel tiles per CTA of the spatial-tile weight-gradient
 *          kernel (default 8): smaller layers use fewer CTAs.
 *   key 11: row-mapped max-pool kernels (default 1).
 *   key 12: streamed-weight rings as deep as shared memory allows (default 1).
 *   key 14: halo kernel with one filter row (three taps) per streamed weight stage
 *          (default 1; plain k x 3 convolutions).
 *   key 15: spatial-tile weight-gradient epilogue as TMA tensor reduce-adds (one [ci x 32 co]
 *          box per tap and column block) instead of one bulk reduce-add per accumulator row
 *          (default 1: 1.056 -> 1.042 ms per U-Net step).
 *   key 16: first-layer kernel (3x3 stride-1 convolution of the 4-channel (R,G,B,1) input of
 *          seg_stage_input, forward and weight gradient; default 1).
 *   key 17: number of SMs the persistent tile kernels size their grids for (default 0 = all).
 *          Data-parallel training sets it to (SM count - all-reduce CTAs): see parallel.py.
 *   key 18: seg_classmap_tail_infer with the transposed conv on warp-level tensor-core MMAs
 *          (default 1; 0 = the CUDA-core form, also taken for unaligned class maps).
 * (Keys 8, 10 and 13 of round 1 - cluster/DSMEM weight-gradient reduction, wave-quantised
 * halo tiles, two-CTA multicast halo clusters - were measured slower and are gone.) */
SEG_API int32_t seg_set_option(int32_t key, int32_t value);
/* ---- convolution: replaces Conv2D(+BiasAdd+Relu), Conv2DBackpropInput,
 * Conv2DBackpropFilter emitted for slim.convolution2d
 * (models/unet.py:111-167, models/fcn.py:110-128, models/deconvolution.py:109-174) */
SEG_API int32_t seg_conv2d_fwd(const seg_conv_desc* d, const seg_view* x, const seg_view* x2,
                       const void* w_bf16, const float* bias, const seg_view* y, void* stream);
/* dx (and dx2 for a virtual concat) = conv_input_grad(dz).  mask_src/mask_src2
 * (nullable, same geometry as dx/dx2) apply the ReluGrad of the layer that
 * produced x.  dz is the gradient w.r.t. this conv's PRE-activation output. */
SEG_API int32_t seg_conv2d_dgrad(const seg_conv_desc* d, const seg_view* dz, const void* w_bf16,
                         const seg_view* dx, const seg_view* dx2, const seg_view* mask_src,
                         const seg_view* mask_src2, void* stream);
/* The same gradient for the input-channel slice [cin_lo, cin_lo + dx->c) only (cin_lo a
 * multiple of 16): a convolution over a virtual concat (models/unet.py:141-142) can send the
 * skip-connection half and the decoder half of its input gradient to different streams. */
SEG_API int32_t seg_conv2d_dgrad_slice(const seg_conv_desc* d, const seg_view* dz,
                                       const void* w_bf16, int32_t cin_lo, const seg_view* dx,
                                       const seg_view* mask_src, void* stream);
/* dw (fp32, master layout [kh][kw][cin][cout]) += x^T * dz;  caller zeroes dw.
 * db (nullable, fp32 [cout]) += sum over pixels of dz (BiasAddGrad), fused: on the
 * tcgen05 path it is one extra all-ones A-atom of the same GEMM. */
SEG_API int32_t seg_conv2d_wgrad(const seg_conv_desc* d, const seg_view* x, const seg_view* x2,
                         const seg_view* dz, float* dw, float* db, void* stream);

/* ---- first layer fused with the max-pool that follows it (models/unet.py:111-120:
 * conv1_1 -> pool1; models/fcn.py:110-117: conv1 -> pool1; models/deconvolution.py:109-118:
 * conv1_0 -> bn1 -> pool1).  x4 is the (R,G,B,1) input of seg_stage_input; 3x3 stride 1 or
 * (forward entries) 5x5 stride 2, cout_pad == 32, even output grid; argmax nullable.
 * fwd: pooled = maxpool2x2/2(relu(conv(x4) + bias)) with the uint8 window slots of
 *   seg_maxpool_fwd in `argmax`, ONE launch; the full-resolution activation is written only
 *   into y_win (nullable), a view of its window whose top-left output pixel is
 *   (win_y0, win_x0) - U-Net's conv1_1 is read again only by conv1_2 on the crop that feeds
 *   the last skip connection (models/unet.py:118-120,159-161), FCN's conv1 never.
 * wgrad: dw, db += gradients for dz = relu_mask(pool_grad(dpool)), i.e. seg_maxpool_bwd_y
 *   followed by seg_conv2d_wgrad without the full-resolution gradient ever existing: the pool
 *   backward is evaluated inside the operand producer (the ReLU mask comes from `pooled`, the
 *   forward pool output: the input at the argmax IS the pooled value).  A gradient arriving
 *   from a second consumer of the activation (U-Net: conv1_2's input gradient on the window)
 *   is a separate linear term: mask it with the activation (seg_conv2d_dgrad's mask_src) and
 *   run seg_conv2d_wgrad on the window - x4 may be a crop view of the staged input.
 * Both return SEG_E_UNSUPPORTED for other shapes (use the unfused entries). */
SEG_API int32_t seg_conv2d_pool_fwd(const seg_conv_desc* d, const seg_view* x4, const void* w_bf16,
                                    const float* bias, const seg_view* y_win, int32_t win_y0,
                                    int32_t win_x0, const seg_view* pooled, uint8_t* argmax,
                                    void* stream);
/* inference form with slim.batch_norm (moving statistics) between the ReLU and the pool:
 * pooled = maxpool2x2/2(bn(relu(conv(x4) + bias))), each stage rounded to bf16 where the
 * unfused entries store bf16 (seg_conv2d_fwd, seg_batchnorm_infer, seg_maxpool_fwd).
 * w_rows_per_tap: rows of the [taps * rows][cout_pad] bf16 weight matrix per filter tap -
 * 0 = d->cin_pad (the padded HWIO shadow), 3 = the dense [kh*kw*3][cout_pad] matrix of a
 * patch-packed first layer. */
SEG_API int32_t seg_conv2d_bn_pool_infer(const seg_conv_desc* d, const seg_view* x4,
                                         const void* w_bf16, int32_t w_rows_per_tap,
                                         const float* bias, const float* bn_mean, const float* bn_var, float bn_eps,
                                         const float* bn_beta, const seg_view* pooled,
                                         uint8_t* argmax, void* stream);
SEG_API int32_t seg_conv2d_pool_wgrad(const seg_conv_desc* d, const seg_view* x4,
                                      const seg_view* dpool, const uint8_t* argmax,
                                      const seg_view* pooled, float* dw, float* db,
                                      void* stream);

/* ---- transposed convolution: replaces Conv2DBackpropInput (+BiasAdd+Relu) and
 * its gradients for slim.convolution2d_transpose (models/unet.py:138,145,152,159;
 * models/deconvolution.py:150,156,159,166).  Weights HWOI [kh][kw][cout][cin].
 * Output geometry comes from y: VALID => y.h = x.h*s + max(k-s,0). */
SEG_API int32_t seg_deconv2d_fwd(const seg_conv_desc* d, const seg_view* x, const void* w_bf16,
                         const float* bias, const seg_view* y, void* stream);
SEG_API int32_t seg_deconv2d_dgrad(const seg_conv_desc* d, const seg_view* dz, const void* w_bf16,
                           const seg_view* dx, const seg_view* mask_src, void* stream);
SEG_API int32_t seg_deconv2d_wgrad(const seg_conv_desc* d, const seg_view* x, const seg_view* dz,
                           float* dw, void* stream);

/* ---- inference forms with the slim.batch_norm that follows the layer folded into the
 * layer (models/deconvolution.py:120-170: conv/deconv -> ReLU -> batch_norm, is_training
 * False): y = relu(op(x) + bias) * post_scale + post_shift with post_scale / post_shift
 * (cout_pad floats each) from seg_batchnorm_fold.  The activation is normalised from the
 * fp32 accumulator, i.e. without the bf16 rounding the unfused pair (seg_conv2d_fwd,
 * seg_batchnorm_infer) has in between.  tcgen05 path only; SEG_E_UNSUPPORTED where the
 * halo-tile kernel cannot take the geometry (use the unfused pair). */
SEG_API int32_t seg_conv2d_fwd_affine(const seg_conv_desc* d, const seg_view* x, const void* w_bf16,
                                      const float* bias, const float* post_scale,
                                      const float* post_shift, const seg_view* y, void* stream);
SEG_API int32_t seg_deconv2d_fwd_affine(const seg_conv_desc* d, const seg_view* x,
                                        const void* w_bf16, const float* bias,
                                        const float* post_scale, const float* post_shift,
                                        const seg_view* y, void* stream);

/* db[c] += sum over pixels of dz[...,c]  (BiasAddGrad).  Caller zeroes db. */
SEG_API int32_t seg_bias_grad(const seg_view* dz, float* db, void* stream);

/* ---- max-pool with argmax: replaces MaxPool / MaxPoolGrad of slim.max_pool2d
 * (models/unet.py:120,124,128,132; models/fcn.py:117-125;
 * models/deconvolution.py:118,130,138).  VALID, k in {2,3}, stride s.
 * argmax: uint8 window slot dy*k+dx of the FIRST max in row-major window order
 * (dense [n][ho][wo][c]).  Backward is a gather:
 *   dx = relu_mask(pool_grad(dy) + add), add (nullable) = gradient arriving from
 * a second consumer of x, given as a view positioned at (add_y0, add_x0) in x
 * (zero outside) — covers the U-Net skip crop (models/unet.py:140-141). */
SEG_API int32_t seg_maxpool_fwd(const seg_view* x, int32_t k, int32_t s, const seg_view* y,
                        uint8_t* argmax, void* stream);
/* inference: y = batch_norm(maxpool(x)) with the moving statistics of the batch-norm that
 * PRECEDES the pool in the model (models/deconvolution.py:126-138: conv -> bn -> pool).  The
 * normalisation has a positive slope and every rounding is monotonic, so the result equals
 * seg_maxpool_fwd(seg_batchnorm_infer(x)) bit for bit without the normalised full-resolution
 * tensor.  k == stride, channels [0, bn_c) are normalised, the rest copied; no argmax. */
SEG_API int32_t seg_maxpool_bn_infer(const seg_view* x, int32_t k, const float* moving_mean,
                                     const float* moving_var, float eps, const float* beta,
                                     int32_t bn_c, const seg_view* y, void* stream);
SEG_API int32_t seg_maxpool_bwd(const seg_view* dy, const uint8_t* argmax, int32_t k, int32_t s,
                        const seg_view* add, int32_t add_y0, int32_t add_x0,
                        const seg_view* mask_src, const seg_view* dx, void* stream);
/* same, given the forward pool output as well: where no `add` gradient arrives the ReLU
 * mask is taken from pooled_y (x at the argmax is the pooled value; elsewhere the routed
 * gradient is zero), so mask_src is read only inside the `add` window. */
SEG_API int32_t seg_maxpool_bwd_y(const seg_view* dy, const uint8_t* argmax, int32_t k, int32_t s,
                          const seg_view* add, int32_t add_y0, int32_t add_x0,
                          const seg_view* mask_src, const seg_view* pooled_y, const seg_view* dx,
                          void* stream);
/* same with two incoming pool-output gradients (dy + dy2): the pooled tensor has
 * two consumers (FCN pool3/pool4 feed the next conv AND a score conv,
 * models/fcn.py:192-195). */
SEG_API int32_t seg_maxpool_bwd2(const seg_view* dy, const seg_view* dy2, const uint8_t* argmax,
                         int32_t k, int32_t s, const seg_view* mask_src, const seg_view* dx,
                         void* stream);
/* ... and given the forward pool output: the ReLU mask comes from pooled_y, the pool input
 * (mask_src) is not read. */
SEG_API int32_t seg_maxpool_bwd2_y(const seg_view* dy, const seg_view* dy2, const uint8_t* argmax,
                                   int32_t k, int32_t s, const seg_view* mask_src,
                                   const seg_view* pooled_y, const seg_view* dx, void* stream);
/* ReluGrad as a copy: dz = (y > 0) ? dy : 0 (dy, y, dz same geometry). */
SEG_API int32_t seg_relu_grad(const seg_view* dy, const seg_view* y, const seg_view* dz, void* stream);

/* ---- depthwise bilinear x f upsample: replaces tf.nn.conv2d_transpose with the
 * constant diagonal filter bank of utils/upsampling.py:27-46 (models/fcn.py:142,
 * 163,171,199,207,215), SAME, k = 2f - f%2, fused with the skip add
 * (models/fcn.py:204,212).  out = up(x) + add (add nullable).  fp32 or bf16 out. */
SEG_API int32_t seg_bilinear_upsample_fwd(const seg_view* x, int32_t factor, const seg_view* add,
                                  const seg_view* y, int32_t y_is_f32, void* stream);
/* mask_src (nullable): ReluGrad of the layer that produced x is applied to dx. */
SEG_API int32_t seg_bilinear_upsample_bwd(const seg_view* dy, int32_t dy_is_f32, int32_t factor,
                                  const seg_view* mask_src, const seg_view* dx, void* stream);

/* ---- tf.image.resize_bilinear legacy (models/deconvolution.py:163) */
SEG_API int32_t seg_resize_bilinear_fwd(const seg_view* x, const seg_view* y, void* stream);
SEG_API int32_t seg_resize_bilinear_bwd(const seg_view* dy, const seg_view* dx, void* stream);

/* ---- slim.batch_norm (center only, after ReLU; models/deconvolution.py:116...)
 * stats: sum[c], sumsq[c] (fp32, caller zeroes).  apply: y=(x-mean)*rstd+beta.
 * finalize turns sums into mean / rstd and updates the moving statistics. */
SEG_API int32_t seg_batchnorm_stats(const seg_view* x, float* sum, float* sumsq, void* stream);
SEG_API int32_t seg_batchnorm_finalize(const float* sum, const float* sumsq, int64_t count, int32_t c,
                               float eps, float decay, float* mean, float* rstd,
                               float* moving_mean, float* moving_var, void* stream);
SEG_API int32_t seg_batchnorm_apply(const seg_view* x, const float* mean, const float* rstd,
                            const float* beta, const seg_view* y, void* stream);
/* inference form: mean / var are the moving statistics */
SEG_API int32_t seg_batchnorm_infer(const seg_view* x, const float* moving_mean,
                            const float* moving_var, float eps, const float* beta,
                            const seg_view* y, void* stream);
/* inference form folded for a conv epilogue (seg_conv2d_fwd_affine): scale[i] =
 * rsqrt(var[i] + eps), shift[i] = beta[i] - mean[i] * scale[i] for i < c, both 0 for
 * c <= i < c_pad */
SEG_API int32_t seg_batchnorm_fold(const float* moving_mean, const float* moving_var, float eps,
                                   const float* beta, int32_t c, int32_t c_pad, float* scale,
                                   float* shift, void* stream);
/* bwd pass 1: dbeta[c] += sum dy, dxhat[c] += sum dy*xhat;  pass 2 writes dx and
 * applies the ReluGrad of the producing conv (x itself is the mask source). */
SEG_API int32_t seg_batchnorm_bwd_reduce(const seg_view* dy, const seg_view* x, const float* mean,
                                 const float* rstd, float* dbeta, float* dxhat, void* stream);
SEG_API int32_t seg_batchnorm_bwd_apply(const seg_view* dy, const seg_view* x, const float* mean,
                                const float* rstd, const float* dbeta, const float* dxhat,
                                int64_t count, int32_t relu_mask, const seg_view* dx,
                                void* stream);

/* ---- slim.dropout keep 0.5 (models/deconvolution.py:129,144,154): Philox-4x32-10
 * keyed by (seed, stream_id), element e -> counter e/4, word e%4; the same call
 * applied to a gradient is the backward.  y = x * keep / keep_prob. */
SEG_API int32_t seg_dropout(const seg_view* x, uint64_t seed, uint32_t stream_id, float keep_prob,
                    const seg_view* y, void* stream);

/* Batched, CUDA-graph-replayable form of the same op: image n of x uses the Philox stream
 *   stream_id0 + n * per_image_step + (step_dev ? *step_dev * step_mul : 0),
 * step_dev a device uint32 the host bumps between graph replays (fresh masks every training
 * step, models/deconvolution.py:128-129).  per_image_step != 0: the element index restarts
 * at every image, so T MC-dropout passes of one tile (BASELINE config 5) are ONE launch with
 * one stream per pass; per_image_step == 0: one stream over the whole batch (== seg_dropout).
 * h*w*c must be a multiple of 8. */
SEG_API int32_t seg_dropout_ex(const seg_view* x, uint64_t seed, uint32_t stream_id0,
                               uint32_t per_image_step, const uint32_t* step_dev,
                               uint32_t step_mul, float keep_prob, const seg_view* y,
                               void* stream);

/* ---- loss: one_hot + softmax_cross_entropy_with_logits + reduce_mean and its
 * gradient (models/basemodel.py:59-70,194,360).  logits fp32 [n,h,w,c]; labels
 * uint8 view (c == 1; may be a centre crop of the full mask, models/unet.py:71-72).
 * loss_sum += sum_pixels xent (caller zeroes; mean = loss_sum / pixels);
 * dlogits (nullable) = (softmax - onehot) / pixels, bf16 with channels padded to
 * dlogits.c (zero filled). */
SEG_API int32_t seg_softmax_xent_fwd_bwd(const seg_view* logits, const seg_view* labels,
                                 float* loss_sum, const seg_view* dlogits, void* stream);
/* FCN-8s training head in ONE launch (models/fcn.py:207-220 upscore x8 -> basemodel.py:59-70
 * loss -> the gradient of both): loss_sum += sum xent(bilinear_up8(x), labels), dx = the x8
 * transposed conv's input gradient of (softmax - onehot) / pixels, optionally masked by the
 * ReluGrad of mask_src; the full-resolution logits / dlogits tensors are not materialised
 * (logits: nullable dense fp32 [n,8h,8w,c], written when given).  x, dx bf16 [n,h,w,c],
 * c <= 32; labels uint8 [n,8h,8w,1].  dx (and logits) are bit-identical to
 * seg_bilinear_upsample_fwd(factor 8, fp32 out) -> seg_softmax_xent_fwd_bwd (bf16 dlogits)
 * -> seg_bilinear_upsample_bwd; the loss differs by the order of its final sum. */
SEG_API int32_t seg_upscore8_xent_fwd_bwd(const seg_view* x, const seg_view* labels,
                                          float* loss_sum, const seg_view* dx,
                                          const seg_view* mask_src, float* logits, void* stream);

/* Fused classification head for training: a 1x1 convolution to n_classes <= 4 (x bf16
 * [n,h,w,cin], cin 16 or 32; w_bf16 = its [cin_pad][cout_pad] shadow; models/unet.py:166-167)
 * + the loss above + the whole backward of that layer in one pass over x:
 * logits (nullable, fp32 [n,h,w,n_classes]) = x.W + b; loss_sum += sum xent;
 * dx = relu_mask_x(dlogits.W^T) (bf16, same geometry as x); dw [cin][n_classes] and
 * db [n_classes] (fp32) are accumulated (+=).  dlogits is rounded to bf16 before it is
 * used, as in the unfused path. */
SEG_API int32_t seg_head1x1_xent(const seg_view* x, const void* w_bf16, int32_t cout_pad,
                         const float* bias, const seg_view* labels, int32_t n_classes,
                         const seg_view* logits, float* loss_sum, const seg_view* dx, float* dw,
                         float* db, void* stream);
/* ---- inference head (models/unet.py:76-79): probs = sigmoid(logits) (fp32),
 * labelmap = float32(argmax(sigmoid(logits), 3)), first index on ties. */
SEG_API int32_t seg_sigmoid_argmax(const seg_view* logits, float* probs, float* labelmap,
                           void* stream);

/* ---- class-map tail of the generic conv/deconvolution model, inference, ONE launch:
 * tf.image.resize_bilinear(x, [rh, rw]) -> 2x2/stride-2 transposed conv (+bias, ReLU) to
 * n_classes -> slim.batch_norm with the moving statistics -> 3x3 SAME conv (+bias, no
 * activation) to n_classes -> sigmoid + argmax (models/deconvolution.py:163-174, :79-82).
 * x bf16 [n,hs,ws,32]; w_up / w_out are the bf16 shadows of the two layers ([2][2][cout_pad]
 * [cin_pad] and [3][3][cin_pad][cout_pad]); 2 <= n_classes <= 4.  Outputs at [n,2rh,2rw]:
 * logits (nullable, fp32 [.,n_classes]), probs (fp32), labelmap (fp32, first index on ties).
 * Every intermediate is rounded to bf16 exactly where the unfused entries store bf16, so
 * the results agree with seg_resize_bilinear_fwd + seg_deconv2d_fwd + seg_batchnorm_infer +
 * seg_conv2d_fwd + seg_sigmoid_argmax up to fp32 summation order. */
SEG_API int32_t seg_classmap_tail_infer(const seg_view* x, int32_t rh, int32_t rw,
                                        const void* w_up_bf16, int32_t up_cout_pad,
                                        int32_t up_cin_pad, const float* b_up,
                                        const float* bn_mean, const float* bn_var, float bn_eps,
                                        const float* bn_beta, const void* w_out_bf16,
                                        int32_t out_cin_pad, int32_t out_cout_pad,
                                        const float* b_out, int32_t n_classes, float* logits,
                                        float* probs, float* labelmap, void* stream);

/* ---- MC-dropout statistics: probs [t][count] -> mean[count], var[count]
 * (population variance over the t passes, Welford). */
SEG_API int32_t seg_mc_mean_var(const float* probs, int32_t t, int64_t count, float* mean, float* var,
                        void* stream);

/* ---- optimizer: tf.train.AdamOptimizer (models/basemodel.py:321,366), epsilon
 * outside the bias correction.  One launch over a flat fp32 parameter buffer.
 * `segments` (device, int32[6*nseg]) maps master elements to the padded bf16
 * shadow: {master_off, numel, inner(=last dim), inner_pad, mid(=dim -2), mid_pad}.
 * grad is scaled by grad_scale (1/world_size for data-parallel) and zeroed.
 * lr_t = lr*sqrt(1-beta2^t)/(1-beta1^t) is read from the device scalar lr_t_dev
 * when non-null (so a captured CUDA graph can be replayed with a new value),
 * else taken from the host argument lr_t.
 * `chunks` (device, int32[2*nchunks]) is the launch plan the host builds once: one
 * {segment index, first element inside the segment} pair per block, each block
 * covering at most seg_adam_chunk_elems() elements of ONE segment. */
SEG_API int32_t seg_adam_chunk_elems(void);
SEG_API int32_t seg_adam_multi(float* param, float* grad, float* m, float* v, void* shadow_bf16,
                       const int32_t* segments, const int64_t* shadow_offsets,
                       const int32_t* chunks, int32_t nchunks, float lr_t,
                       const float* lr_t_dev, float beta1, float beta2, float eps,
                       float grad_scale, void* stream);

/* ---- input staging (BaseModel._init_input, models/basemodel.py:145-177, fed by
 * utils/datasets.py:176-190): source pixels -> y4, a dense bf16 [n,h,w,4] tensor holding
 * (R, G, B, 1) per pixel - the layout the first-layer convolution kernel reads (its fourth
 * weight row is zero; in the weight gradient the constant channel yields the bias gradient).
 *   x_kind 0: x is fp32 [n,src_h,src_w,c] in [0,1];  1: uint8, divided by 255 in fp32
 *             (utils/datasets.py:178 / :41) before the bf16 rounding.
 *   crop_yx (device int32 [n][2], nullable): top-left corner of image n's h x w window in
 *             the source (the joint image+mask random crop of utils/datasets.py:184-185; the
 *             caller draws the offsets).  Null: src_h == h and src_w == w.
 *   mask_src (nullable, uint8 [n,src_h,src_w]) -> mask_dst (uint8 [n,h,w], same window):
 *             mask_kind 0 copies class labels, 1 maps 255 -> 1 and everything else -> 0
 *             (uint8(mask / 255), utils/datasets.py:179,172).
 *   ctl (host struct, nullable): per-step scalars written by the same launch, so that a
 *             training step is [this kernel, one CUDA-graph replay]: *loss_sum is published
 *             to host_ring[2*(publish_step&3)] = {loss_sum, bits of publish_step} (pinned
 *             host memory, device-accessible) when publish_step >= 0, then zeroed;
 *             *lr_t_dev = lr_t; *step_dev = step.  Null members are skipped. */
typedef struct seg_stage_ctl {
  float* loss_sum;
  float* host_ring;
  int32_t publish_step;
  float lr_t;
  float* lr_t_dev;
  int32_t step;
  int32_t* step_dev;
} seg_stage_ctl;
SEG_API int32_t seg_stage_input(const void* x, int32_t x_kind, int32_t c, int32_t src_h,
                                int32_t src_w, const int32_t* crop_yx, const seg_view* y4,
                                const uint8_t* mask_src, int32_t mask_kind, uint8_t* mask_dst,
                                const seg_stage_ctl* ctl, void* stream);
/* ---- layout helpers */
/* fp32 NHWC [n,h,w,c] -> bf16 NHWC with channels zero-padded to y.c */
SEG_API int32_t seg_pack_input(const float* x, int32_t c, const seg_view* y, void* stream);
/* fp32 NHWC [n,h,w,c] -> bf16 y [n,ho,wo,CP] with the kh x kw x c input patch of every output
 * pixel packed into the channel axis: y[.., (r*kw+s)*c+ci] = x[.., oy*stride+r-pad_t,
 * ox*stride+s-pad_l, ci], zero outside the image and for channels >= kh*kw*c.  Turns the
 * first convolution of a model (3 input channels: `slim.convolution2d(input, ...)`,
 * models/unet.py:111, models/fcn.py:110) into a 1x1 convolution over CP channels whose
 * weight matrix is the TF HWIO tensor read as [kh*kw*c][cout]. */
SEG_API int32_t seg_pack_patches(const float* x, int32_t c, int32_t h, int32_t w, int32_t kh,
                         int32_t kw, int32_t stride, int32_t pad_t, int32_t pad_l,
                         const seg_view* y, void* stream);
SEG_API int32_t seg_fill_zero(void* ptr, int64_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEGB200_H_ */
