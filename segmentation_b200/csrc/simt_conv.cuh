// Parameter blocks + launchers of the CUDA-core convolution kernels (simt_conv.cu).
#pragma once
#include "common.cuh"

namespace segb {

struct DirectParams {
  seg_view x, x2;          // input (virtual concat [x | x2]); x2.ptr may be null
  const bf16* w;           // [taps][in_pad][out_pad]
  const float* bias;
  seg_view y;              // output view (bf16 or fp32)
  seg_view mask;           // optional relu-mask source (geometry of y)
  int kh, kw, stride, pad_t, pad_l;
  int in_pad, out_pad, flags;
};

struct TransParams {
  seg_view src;            // small tensor (dz for conv dgrad, x for deconv fwd)
  const bf16* w;           // [taps][oc_pad][ic_pad]  (oc = channel of `out`)
  const float* bias;
  seg_view out, out2;      // output(s): oc in [0,out.c) -> out, rest -> out2
  seg_view mask, mask2;
  int kh, kw, stride, pad_t, pad_l;
  int oc_pad, ic_pad, flags;
};

struct WgradParams {
  seg_view big, big2;      // tensor indexed with stride/tap (virtual concat allowed)
  seg_view small_;         // tensor indexed by m
  float* dw;               // [taps][BC][SC] fp32, accumulated with atomics
  int kh, kw, stride, pad_t, pad_l;
  int BC, SC;              // logical channel counts written to dw
  int pix_per_split;
};

int simt_direct(const DirectParams& P, cudaStream_t st);
// <= 4 real input and output channels, stride 1: streaming one-thread-per-pixel kernel
bool simt_tiny_conv_ok(const DirectParams& P, int cin);
int simt_tiny_conv(const DirectParams& P, int cin, cudaStream_t st);
int simt_transposed(const TransParams& P, cudaStream_t st);
int simt_wgrad(WgradParams P, cudaStream_t st);

}  // namespace segb
