// CUDA-core (SIMT) convolution kernels: geometry-general direct, transposed
// (gather form) and correlation-wgrad kernels.  They cover every kernel size /
// stride / padding the three reference graphs use and serve as
//   (a) the on-device cross-check for the tcgen05 implicit-GEMM kernels, and
//   (b) the execution path for layer shapes the tcgen05 path does not take yet.
// bf16 operands, fp32 accumulation, same storage points as the tcgen05 path.
#include "simt_conv.cuh"

namespace segb {

static int grid_for(int64_t total, int block) {
  int64_t g = ceil_div64(total, block);
  const int64_t cap = (int64_t)num_sms() * 32;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

// out[n,p,q,oc] = sum_{r,s,ic} in[n,p*st+r-pt,q*st+s-pl,ic] * w[r,s,ic,oc]
__global__ void direct_conv_kernel(DirectParams P) {
  const int ncg = P.out_pad / 8;
  const int64_t total = (int64_t)P.y.n * P.y.h * P.y.w * ncg;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int cg = idx % ncg;
    int64_t m = idx / ncg;
    const int q = m % P.y.w;
    m /= P.y.w;
    const int p = m % P.y.h;
    const int n = m / P.y.h;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int r = 0; r < P.kh; ++r) {
      const int iy = p * P.stride + r - P.pad_t;
      if (iy < 0 || iy >= P.x.h) continue;
      for (int s = 0; s < P.kw; ++s) {
        const int ix = q * P.stride + s - P.pad_l;
        if (ix < 0 || ix >= P.x.w) continue;
        const bf16* wt = P.w + ((int64_t)(r * P.kw + s) * P.in_pad) * P.out_pad + cg * 8;
        const bf16* xp = view_at(P.x, n, iy, ix);
        for (int ic = 0; ic < P.x.c; ++ic) {
          const float xv = __bfloat162float(xp[ic]);
          const uint4 wv = *reinterpret_cast<const uint4*>(wt + (int64_t)ic * P.out_pad);
          const uint32_t wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[2 * j] += xv * __uint_as_float(wr[j] << 16);
            acc[2 * j + 1] += xv * __uint_as_float(wr[j] & 0xFFFF0000u);
          }
        }
        if (P.x2.ptr) {
          const bf16* xp2 = view_at(P.x2, n, iy, ix);
          for (int ic = 0; ic < P.x2.c; ++ic) {
            const float xv = __bfloat162float(xp2[ic]);
            const uint4 wv =
                *reinterpret_cast<const uint4*>(wt + (int64_t)(ic + P.x.c) * P.out_pad);
            const uint32_t wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[2 * j] += xv * __uint_as_float(wr[j] << 16);
              acc[2 * j + 1] += xv * __uint_as_float(wr[j] & 0xFFFF0000u);
            }
          }
        }
      }
    }
    const int64_t off = n * P.y.sn + p * P.y.sh + q * P.y.sw;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int oc = cg * 8 + j;
      if (oc >= P.y.c) break;
      float v = acc[j];
      if (P.flags & SEG_EPI_BIAS) v += P.bias[oc];
      if (P.flags & SEG_EPI_RELU) v = fmaxf(v, 0.f);
      if (P.flags & SEG_EPI_RELU_MASK) {
        const float mv = __bfloat162float(view_at(P.mask, n, p, q)[oc]);
        if (!(mv > 0.f)) v = 0.f;
      }
      if (P.flags & SEG_EPI_OUT_F32)
        reinterpret_cast<float*>(P.y.ptr)[off + oc] = v;
      else
        reinterpret_cast<bf16*>(P.y.ptr)[off + oc] = __float2bfloat16(v);
    }
  }
}

// Convolution with <= 4 real input and <= 4 real output channels, stride 1 - a head that maps
// class scores to class scores (DeconvModel conv_out 2 -> 2 at full resolution, reference
// models/deconvolution.py:174).  One thread per output pixel, the whole filter bank in
// shared memory, one 8-byte load per tap.  The tensor-core path would multiply 16 x 16
// padded channels and is epilogue bound (4.6 ms at 32 x 1024 x 1024); this one streams.
__global__ void tiny_conv_kernel(DirectParams P, int cin) {
  __shared__ float sw[25 * 16];
  __shared__ float sb[4];
  const int taps = P.kh * P.kw;
  for (int i = threadIdx.x; i < taps * 16; i += blockDim.x) {
    const int t = i >> 4, ic = (i >> 2) & 3, oc = i & 3;
    sw[i] = (ic < cin && oc < P.y.c)
                ? __bfloat162float(P.w[((int64_t)t * P.in_pad + ic) * P.out_pad + oc]) : 0.f;
  }
  if (threadIdx.x < 4)
    sb[threadIdx.x] = ((P.flags & SEG_EPI_BIAS) && (int)threadIdx.x < P.y.c) ? P.bias[threadIdx.x] : 0.f;
  __syncthreads();
  const int64_t total = (int64_t)P.y.n * P.y.h * P.y.w;
  for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < total;
       m += (int64_t)gridDim.x * blockDim.x) {
    const int q = m % P.y.w;
    const int64_t t2 = m / P.y.w;
    const int p = t2 % P.y.h;
    const int n = t2 / P.y.h;
    float acc[4] = {sb[0], sb[1], sb[2], sb[3]};
    for (int r = 0; r < P.kh; ++r) {
      const int iy = p + r - P.pad_t;
      if (iy < 0 || iy >= P.x.h) continue;
      for (int s = 0; s < P.kw; ++s) {
        const int ix = q + s - P.pad_l;
        if (ix < 0 || ix >= P.x.w) continue;
        const uint2 u = *reinterpret_cast<const uint2*>(view_at(P.x, n, iy, ix));
        const float xv[4] = {__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u),
                             __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u)};
        const float* wt = sw + (r * P.kw + s) * 16;
#pragma unroll
        for (int ic = 0; ic < 4; ++ic)
#pragma unroll
          for (int oc = 0; oc < 4; ++oc) acc[oc] += xv[ic] * wt[ic * 4 + oc];
      }
    }
    const int64_t off = n * P.y.sn + p * P.y.sh + q * P.y.sw;
#pragma unroll
    for (int oc = 0; oc < 4; ++oc) {
      if (oc >= P.y.c) break;
      float v = acc[oc];
      if (P.flags & SEG_EPI_RELU) v = fmaxf(v, 0.f);
      if (P.flags & SEG_EPI_OUT_F32)
        reinterpret_cast<float*>(P.y.ptr)[off + oc] = v;
      else
        reinterpret_cast<bf16*>(P.y.ptr)[off + oc] = __float2bfloat16(v);
    }
  }
}

bool simt_tiny_conv_ok(const DirectParams& P, int cin) {
  return cin <= 4 && P.y.c <= 4 && P.stride == 1 && P.x2.ptr == nullptr && P.kh * P.kw <= 25 &&
         !(P.flags & SEG_EPI_RELU_MASK) && P.x.sw >= 4 && P.x.sw % 4 == 0 && P.x.sh % 4 == 0 &&
         P.x.sn % 4 == 0 && (reinterpret_cast<uintptr_t>(P.x.ptr) & 7) == 0;
}

int simt_tiny_conv(const DirectParams& P, int cin, cudaStream_t st) {
  const int64_t total = (int64_t)P.y.n * P.y.h * P.y.w;
  tiny_conv_kernel<<<grid_for(total, 256), 256, 0, st>>>(P, cin);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

// out[n,h,w,oc] = sum_{r,s | (h+pt-r)%st==0} sum_ic src[n,(h+pt-r)/st,(w+pl-s)/st,ic]*w[r,s,oc,ic]
__global__ void transposed_conv_kernel(TransParams P) {
  const int OC = P.out.c + (P.out2.ptr ? P.out2.c : 0);
  const int64_t total = (int64_t)P.out.n * P.out.h * P.out.w * OC;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int oc = idx % OC;
    int64_t m = idx / OC;
    const int x = m % P.out.w;
    m /= P.out.w;
    const int y = m % P.out.h;
    const int n = m / P.out.h;
    float acc = 0.f;
    for (int r = 0; r < P.kh; ++r) {
      const int ty = y + P.pad_t - r;
      if (ty < 0 || ty % P.stride) continue;
      const int py = ty / P.stride;
      if (py >= P.src.h) continue;
      for (int s = 0; s < P.kw; ++s) {
        const int tx = x + P.pad_l - s;
        if (tx < 0 || tx % P.stride) continue;
        const int px = tx / P.stride;
        if (px >= P.src.w) continue;
        const bf16* sp = view_at(P.src, n, py, px);
        const bf16* wt = P.w + ((int64_t)(r * P.kw + s) * P.oc_pad + oc) * P.ic_pad;
        for (int ic = 0; ic < P.src.c; ++ic)
          acc += __bfloat162float(sp[ic]) * __bfloat162float(wt[ic]);
      }
    }
    float v = acc;
    if (P.flags & SEG_EPI_BIAS) v += P.bias[oc];
    if (P.flags & SEG_EPI_RELU) v = fmaxf(v, 0.f);
    const bool second = oc >= P.out.c;
    const seg_view& O = second ? P.out2 : P.out;
    const int ocl = second ? oc - P.out.c : oc;
    if (P.flags & SEG_EPI_RELU_MASK) {
      const seg_view& Mv = second ? P.mask2 : P.mask;
      if (Mv.ptr) {
        const float mv = __bfloat162float(view_at(Mv, n, y, x)[ocl]);
        if (!(mv > 0.f)) v = 0.f;
      }
    }
    const int64_t off = n * O.sn + y * O.sh + x * O.sw + ocl;
    if (P.flags & SEG_EPI_OUT_F32)
      reinterpret_cast<float*>(O.ptr)[off] = v;
    else
      reinterpret_cast<bf16*>(O.ptr)[off] = __float2bfloat16(v);
  }
}

// dw[r,s,bc,sc] += sum_m big[n,p*st+r-pt,q*st+s-pl,bc] * small[n,p,q,sc]
__global__ void corr_wgrad_kernel(WgradParams P) {
  const int bc_tiles = (P.BC + 7) / 8;
  const int tap = blockIdx.x / bc_tiles;
  const int bc0 = (blockIdx.x % bc_tiles) * 8;
  const int r = tap / P.kw, s = tap % P.kw;
  int sc_threads = 1;
  while (sc_threads < P.SC && sc_threads < (int)blockDim.x) sc_threads <<= 1;
  const int m_par = blockDim.x / sc_threads;
  const int sc_lane = threadIdx.x % sc_threads;
  const int m_lane = threadIdx.x / sc_threads;
  const int64_t M = (int64_t)P.small_.n * P.small_.h * P.small_.w;
  const int64_t m_begin = (int64_t)blockIdx.y * P.pix_per_split;
  const int64_t m_end = min(M, m_begin + P.pix_per_split);
  for (int sc = sc_lane; sc < P.SC; sc += sc_threads) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int64_t m = m_begin + m_lane; m < m_end; m += m_par) {
      const int q = m % P.small_.w;
      const int64_t t = m / P.small_.w;
      const int p = t % P.small_.h;
      const int n = t / P.small_.h;
      const int iy = p * P.stride + r - P.pad_t;
      const int ix = q * P.stride + s - P.pad_l;
      if (iy < 0 || iy >= P.big.h || ix < 0 || ix >= P.big.w) continue;
      const float sv = __bfloat162float(view_at(P.small_, n, p, q)[sc]);
      const bf16* bp = view_at(P.big, n, iy, ix);
      const bf16* bp2 = P.big2.ptr ? view_at(P.big2, n, iy, ix) : nullptr;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int bc = bc0 + j;
        float bv = 0.f;
        if (bc < P.big.c)
          bv = __bfloat162float(bp[bc]);
        else if (bp2 && bc - P.big.c < P.big2.c)
          bv = __bfloat162float(bp2[bc - P.big.c]);
        acc[j] += bv * sv;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int bc = bc0 + j;
      if (bc < P.BC && acc[j] != 0.f)
        atomicAdd(P.dw + ((int64_t)tap * P.BC + bc) * P.SC + sc, acc[j]);
    }
  }
}


int simt_direct(const DirectParams& P, cudaStream_t st) {
  const int64_t total = (int64_t)P.y.n * P.y.h * P.y.w * (P.out_pad / 8);
  direct_conv_kernel<<<grid_for(total, 128), 128, 0, st>>>(P);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

int simt_transposed(const TransParams& P, cudaStream_t st) {
  const int OC = P.out.c + (P.out2.ptr ? P.out2.c : 0);
  const int64_t total = (int64_t)P.out.n * P.out.h * P.out.w * OC;
  transposed_conv_kernel<<<grid_for(total, 128), 128, 0, st>>>(P);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

int simt_wgrad(WgradParams P, cudaStream_t st) {
  const int taps = P.kh * P.kw;
  const int bc_tiles = (P.BC + 7) / 8;
  const int64_t M = (int64_t)P.small_.n * P.small_.h * P.small_.w;
  // enough m-splits for ~8 blocks per SM, at least 512 pixels per block
  int64_t want = ceil_div64((int64_t)num_sms() * 8, (int64_t)taps * bc_tiles);
  int64_t pps = ceil_div64(M, want > 0 ? want : 1);
  if (pps < 512) pps = 512;
  P.pix_per_split = (int)pps;
  dim3 grid(taps * bc_tiles, (unsigned)ceil_div64(M, pps));
  corr_wgrad_kernel<<<grid, 128, 0, st>>>(P);
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

}  // namespace segb
