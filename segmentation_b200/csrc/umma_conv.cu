// Host side of the tcgen05 implicit-GEMM kernels: tensor-map construction (tiled
// and im2col), tile-shape selection and launches.
#include <cstdio>
#include <cstdlib>
#include "umma_conv.cuh"
#include "hconv.cuh"
#include "tconv.cuh"
#include "twgrad.cuh"

#include <mutex>

namespace segb {

// ---------------------------------------------------------------------------
// driver entry points (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                   cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode_tiled = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;
static int g_driver_version = 0;

static int load_encoders() {
  static std::once_flag once;
  static int status = SEG_OK;
  std::call_once(once, []() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) !=
            cudaSuccess || !fn) {
      status = SEG_E_CUDA;
      return;
    }
    g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
    fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q) !=
            cudaSuccess || !fn) {
      status = SEG_E_CUDA;
      return;
    }
    g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
    cudaDriverGetVersion(&g_driver_version);
  });
  if (status != SEG_OK) set_error("cuTensorMapEncode* driver entry points unavailable");
  return status;
}

static CUtensorMapSwizzle swizzle_enum(int bytes) {
  return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
         : bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
         : bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                       : CU_TENSOR_MAP_SWIZZLE_NONE;
}

// 2-D row-major bf16 matrix [rows][cols] with row pitch `ld` elements.
static int make_tmap_2d(CUtensorMap* tm, const void* ptr, int64_t cols, int64_t rows, int64_t ld,
                        int box_cols, int box_rows, int swizzle_bytes) {
  SEG_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * 2) % 16 == 0, SEG_E_ALIGN,
              "tensor map: base/pitch must be 16-byte aligned (ptr=%p ld=%lld)", ptr,
              (long long)ld);
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t est[2] = {1, 1};
  CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr),
                              gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              swizzle_enum(swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SEG_REQUIRE(r == CUDA_SUCCESS, SEG_E_CUDA,
              "cuTensorMapEncodeTiled failed (%d) cols=%lld rows=%lld ld=%lld box=%dx%d sw=%d",
              (int)r, (long long)cols, (long long)rows, (long long)ld, box_cols, box_rows,
              swizzle_bytes);
  return SEG_OK;
}

// im2col map over an NHWC view seen as (C, W, H, N).
static int make_tmap_im2col(CUtensorMap* tm, const seg_view& v, int low_w, int low_h, int up_w,
                            int up_h, int stride, int channels, int pixels, int swizzle_bytes) {
  SEG_REQUIRE((reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0 && (v.sw * 2) % 16 == 0 &&
                  (v.sh * 2) % 16 == 0 && (v.sn * 2) % 16 == 0,
              SEG_E_ALIGN, "im2col tensor map: view must be 16-byte aligned in every stride");
  SEG_REQUIRE(low_w >= -128 && low_w <= 127 && low_h >= -128 && low_h <= 127 && up_w >= -128 &&
                  up_w <= 127 && up_h >= -128 && up_h <= 127,
              SEG_E_UNSUPPORTED, "im2col corners out of range");
  cuuint64_t gdim[4] = {(cuuint64_t)v.c, (cuuint64_t)v.w, (cuuint64_t)v.h, (cuuint64_t)v.n};
  cuuint64_t gstr[3] = {(cuuint64_t)v.sw * 2, (cuuint64_t)v.sh * 2, (cuuint64_t)v.sn * 2};
  int lower[2] = {low_w, low_h};
  int upper[2] = {up_w, up_h};
  cuuint32_t est[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = g_encode_im2col(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, v.ptr, gdim, gstr, lower,
                               upper, (cuuint32_t)channels, (cuuint32_t)pixels, est,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(swizzle_bytes),
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SEG_REQUIRE(r == CUDA_SUCCESS, SEG_E_CUDA,
              "cuTensorMapEncodeIm2col failed (%d) c=%d w=%d h=%d n=%d low=(%d,%d) up=(%d,%d) "
              "st=%d ch=%d px=%d sw=%d",
              (int)r, v.c, v.w, v.h, v.n, low_w, low_h, up_w, up_h, stride, channels, pixels,
              swizzle_bytes);
  // Drivers up to CUDA 13.1 mis-encode im2col maps of tensors smaller than 128 KiB
  // (bit 21 of the second descriptor word must be cleared); same fix-up as the one
  // NVIDIA's own conv templates apply after cuTensorMapEncodeIm2col.
  if (g_driver_version <= 13010) {
    const int64_t span = (int64_t)(v.n - 1) * v.sn + (int64_t)(v.h - 1) * v.sh +
                         (int64_t)(v.w - 1) * v.sw + v.c;
    if (span * 2 < 131072) reinterpret_cast<uint64_t*>(tm)[1] &= ~(1ull << 21);
  }
  return SEG_OK;
}

int make_probe_tmap(CUtensorMap* tm, const void* ptr, int64_t cols, int64_t rows, int box_cols,
                    int box_rows, int swizzle_bytes) {
  int rc = load_encoders();
  if (rc) return rc;
  return make_tmap_2d(tm, ptr, cols, rows, cols, box_cols, box_rows, swizzle_bytes);
}

// tiled 4-D map over an NHWC view seen as (C, W, H, N); box = {channels, width, 1, 1}
static int make_tmap_rows(CUtensorMap* tm, const seg_view& v, int channels, int width,
                          int swizzle_bytes) {
  SEG_REQUIRE((reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0 && (v.sw * 2) % 16 == 0 &&
                  (v.sh * 2) % 16 == 0 && (v.sn * 2) % 16 == 0,
              SEG_E_ALIGN, "row tensor map: view must be 16-byte aligned in every stride");
  cuuint64_t gdim[4] = {(cuuint64_t)v.c, (cuuint64_t)v.w, (cuuint64_t)v.h, (cuuint64_t)v.n};
  cuuint64_t gstr[3] = {(cuuint64_t)v.sw * 2, (cuuint64_t)v.sh * 2, (cuuint64_t)v.sn * 2};
  cuuint32_t box[4] = {(cuuint32_t)channels, (cuuint32_t)width, 1, 1};
  cuuint32_t est[4] = {1, 1, 1, 1};
  CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, v.ptr, gdim, gstr, box, est,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(swizzle_bytes),
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SEG_REQUIRE(r == CUDA_SUCCESS, SEG_E_CUDA,
              "cuTensorMapEncodeTiled(4d rows) failed (%d) c=%d w=%d h=%d n=%d box=%dx%d sw=%d",
              (int)r, v.c, v.w, v.h, v.n, channels, width, swizzle_bytes);
  return SEG_OK;
}

static int pick_chunk(int c1, int c2) {
  const int cands[3] = {64, 32, 16};
  for (int k : cands)
    if (c1 % k == 0 && c2 % k == 0 && c1 >= k) return k;
  return 0;
}

static bool pixel_dense(const seg_view& v) {
  return v.sh == (int64_t)v.w * v.sw && v.sn == (int64_t)v.h * v.sh;
}

static EpiDest make_dest(const seg_view* v, const seg_view* mask) {
  EpiDest d;
  memset(&d, 0, sizeof(d));
  if (v && v->ptr) {
    d.ptr = v->ptr;
    d.sn = v->sn; d.sh = v->sh; d.sw = v->sw;
    d.cols = v->c;
  }
  if (mask && mask->ptr) {
    d.mask = reinterpret_cast<const bf16*>(mask->ptr);
    d.msn = mask->sn; d.msh = mask->sh; d.msw = mask->sw;
  }
  return d;
}

// ---------------------------------------------------------------------------
// igemm launch
// ---------------------------------------------------------------------------
struct IgemmJob {
  seg_view a1, a2;             // activation sources (a2.ptr == null: none)
  int kh, kw, stride;
  int low_h, low_w, up_h, up_w;
  int Ho, Wo, batch;           // base-pixel grid
  const void* w;               // 2-D bf16 weight matrix
  int w_rows, w_cols;
  bool b_mn;
  int b_rows_per_tap;
  bool tap_flip;
  int N_total;
  int max_bn;                  // BN must divide this (tile may not straddle taps / dests)
  EpiDest d0, d1;
  int split_n;
  const float* bias;
  int flags;
  int ps_k, ps_cout;
  bool a_tiled2d;
};

template <int KC, int BN, bool B_MN>
static int launch_igemm_t(const IgemmJob& J, cudaStream_t st) {
  using Cfg = IgemmCfg<KC, BN>;
  static bool attr_done = false;
  if (!attr_done) {
    SEG_CHECK_CUDA(cudaFuncSetAttribute(igemm_kernel<KC, BN, B_MN>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::kSmemBytes));
    attr_done = true;
  }
  CUtensorMap tmA1, tmA2, tmB;
  int rc;
  if (J.a_tiled2d) {
    const int64_t rows = (int64_t)J.batch * J.Ho * J.Wo;
    rc = make_tmap_2d(&tmA1, J.a1.ptr, J.a1.c, rows, J.a1.sw, KC, kBlockM, KC * 2);
    if (rc) return rc;
    tmA2 = tmA1;
  } else {
    rc = make_tmap_im2col(&tmA1, J.a1, J.low_w, J.low_h, J.up_w, J.up_h, J.stride, KC, kBlockM,
                          KC * 2);
    if (rc) return rc;
    if (J.a2.ptr) {
      rc = make_tmap_im2col(&tmA2, J.a2, J.low_w, J.low_h, J.up_w, J.up_h, J.stride, KC, kBlockM,
                            KC * 2);
      if (rc) return rc;
    } else {
      tmA2 = tmA1;
    }
  }
  if (B_MN)
    rc = make_tmap_2d(&tmB, J.w, J.w_cols, J.w_rows, J.w_cols, Cfg::kAtomN, KC, Cfg::kAtomN * 2);
  else
    rc = make_tmap_2d(&tmB, J.w, J.w_cols, J.w_rows, J.w_cols, KC, BN, KC * 2);
  if (rc) return rc;

  IgemmParams P;
  memset(&P, 0, sizeof(P));
  P.a_tiled2d = J.a_tiled2d ? 1 : 0;
  P.M_total = J.batch * J.Ho * J.Wo;
  P.Ho = J.Ho; P.Wo = J.Wo;
  P.stride = J.stride;
  P.base_h = J.low_h; P.base_w = J.low_w;
  P.kh = J.kh; P.kw = J.kw;
  P.chunks1 = J.a1.c / KC;
  P.chunks2 = J.a2.ptr ? J.a2.c / KC : 0;
  P.tap_flip = J.tap_flip ? 1 : 0;
  P.b_rows_per_tap = J.b_rows_per_tap;
  P.N_total = J.N_total;
  P.d0 = J.d0; P.d1 = J.d1; P.split_n = J.split_n;
  P.bias = J.bias;
  P.flags = J.flags;
  P.ps_k = J.ps_k; P.ps_cout = J.ps_cout;
  const int m_tiles = (P.M_total + kBlockM - 1) / kBlockM;
  const int tiles = m_tiles * (J.N_total / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  SEG_CHECK_CUDA(launch_k(igemm_kernel<KC, BN, B_MN>, dim3(grid), dim3(kConvThreads), (size_t)(Cfg::kSmemBytes), st, tmA1, tmA2, tmB, P));
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

template <int KC, bool B_MN>
static int launch_igemm_bn(const IgemmJob& J, int BN, cudaStream_t st) {
  switch (BN) {
    case 256: return launch_igemm_t<KC, 256, B_MN>(J, st);
    case 128: return launch_igemm_t<KC, 128, B_MN>(J, st);
    case 64: return launch_igemm_t<KC, 64, B_MN>(J, st);
    case 32: return launch_igemm_t<KC, 32, B_MN>(J, st);
    case 16: return launch_igemm_t<KC, 16, B_MN>(J, st);
  }
  set_error("igemm: unsupported BN %d", BN);
  return SEG_E_UNSUPPORTED;
}

static int launch_igemm(const IgemmJob& J, cudaStream_t st) {
  int rc = load_encoders();
  if (rc) return rc;
  const int KC = pick_chunk(J.a1.c, J.a2.ptr ? J.a2.c : 0);
  SEG_REQUIRE(KC != 0, SEG_E_UNSUPPORTED, "igemm: channel counts (%d,%d) not multiples of 16",
              J.a1.c, J.a2.ptr ? J.a2.c : 0);
  // BN: largest of {256..16} dividing max_bn; shrink while the grid underfills the SMs.
  int BN = 0;
  for (int c = 256; c >= 16; c >>= 1)
    if (J.max_bn % c == 0) { BN = c; break; }
  SEG_REQUIRE(BN != 0 && J.N_total % BN == 0, SEG_E_UNSUPPORTED,
              "igemm: N=%d (tile bound %d) not a multiple of 16", J.N_total, J.max_bn);
  const int64_t m_tiles = ceil_div64((int64_t)J.batch * J.Ho * J.Wo, kBlockM);
  while (BN > 64 && m_tiles * (J.N_total / BN) < num_sms()) BN >>= 1;
  if (J.b_mn) {
    switch (KC) {
      case 64: return launch_igemm_bn<64, true>(J, BN, st);
      case 32: return launch_igemm_bn<32, true>(J, BN, st);
      default: return launch_igemm_bn<16, true>(J, BN, st);
    }
  }
  switch (KC) {
    case 64: return launch_igemm_bn<64, false>(J, BN, st);
    case 32: return launch_igemm_bn<32, false>(J, BN, st);
    default: return launch_igemm_bn<16, false>(J, BN, st);
  }
}

// ---------------------------------------------------------------------------
// wgrad launch
// ---------------------------------------------------------------------------
struct WgradJob {
  seg_view big, big2, small_;
  int kh, kw, stride, low_h, low_w, up_h, up_w;
  int BC, SC;
  float* dw;
  float* db;
};

template <int AW, int BN>
static int launch_wgrad_t(const WgradJob& J, cudaStream_t st) {
  using Cfg = WgradCfg<AW, BN>;
  static bool attr_done = false;
  if (!attr_done) {
    SEG_CHECK_CUDA(cudaFuncSetAttribute(wgrad_kernel<AW, BN>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::kSmemBytes));
    attr_done = true;
  }
  CUtensorMap tmA1, tmA2, tmB;
  int rc = make_tmap_im2col(&tmA1, J.big, J.low_w, J.low_h, J.up_w, J.up_h, J.stride, AW,
                            kWgradPK, AW * 2);
  if (rc) return rc;
  if (J.big2.ptr) {
    rc = make_tmap_im2col(&tmA2, J.big2, J.low_w, J.low_h, J.up_w, J.up_h, J.stride, AW, kWgradPK,
                          AW * 2);
    if (rc) return rc;
  } else {
    tmA2 = tmA1;
  }
  const int64_t M = (int64_t)J.small_.n * J.small_.h * J.small_.w;
  rc = make_tmap_2d(&tmB, J.small_.ptr, J.small_.c, M, J.small_.sw, Cfg::kAtomN, kWgradPK,
                    Cfg::kAtomN * 2);
  if (rc) return rc;

  WgradUmmaParams P;
  memset(&P, 0, sizeof(P));
  P.M_total = (int)M;
  P.Ho = J.small_.h; P.Wo = J.small_.w;
  P.stride = J.stride;
  P.base_h = J.low_h; P.base_w = J.low_w;
  P.kh = J.kh; P.kw = J.kw;
  P.chunks1 = J.big.c / AW;
  P.chunks2 = J.big2.ptr ? J.big2.c / AW : 0;
  P.total_atoms = J.kh * J.kw * (P.chunks1 + P.chunks2);
  P.n_tiles = J.small_.c / BN;
  P.BC = J.BC; P.SC = J.SC;
  P.dw = J.dw;
  P.db = J.db;
  const int groups = (P.total_atoms + (J.db ? 1 : 0) + Cfg::kNA - 1) / Cfg::kNA;
  const int base_ctas = groups * P.n_tiles;
  const int total_kb = (int)ceil_div64(M, kWgradPK);
  int splits = (2 * num_sms() + base_ctas - 1) / base_ctas;
  if (splits > total_kb) splits = total_kb;
  if (splits < 1) splits = 1;
  P.kb_per_split = (total_kb + splits - 1) / splits;
  splits = (total_kb + P.kb_per_split - 1) / P.kb_per_split;
  dim3 grid(base_ctas, splits);
  SEG_CHECK_CUDA(launch_k(wgrad_kernel<AW, BN>, dim3(grid), dim3(kIgemmThreads), (size_t)(Cfg::kSmemBytes), st, tmA1, tmA2, tmB, P));
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

template <int AW>
static int launch_wgrad_bn(const WgradJob& J, int BN, cudaStream_t st) {
  switch (BN) {
    case 256: return launch_wgrad_t<AW, 256>(J, st);
    case 128: return launch_wgrad_t<AW, 128>(J, st);
    case 64: return launch_wgrad_t<AW, 64>(J, st);
    case 32: return launch_wgrad_t<AW, 32>(J, st);
    case 16: return launch_wgrad_t<AW, 16>(J, st);
  }
  set_error("wgrad: unsupported BN %d", BN);
  return SEG_E_UNSUPPORTED;
}

static int launch_wgrad(const WgradJob& J, cudaStream_t st) {
  int rc = load_encoders();
  if (rc) return rc;
  const int AW = pick_chunk(J.big.c, J.big2.ptr ? J.big2.c : 0);
  SEG_REQUIRE(AW != 0, SEG_E_UNSUPPORTED, "wgrad: channel counts not multiples of 16");
  SEG_REQUIRE(pixel_dense(J.small_) && J.small_.sw == J.small_.c, SEG_E_UNSUPPORTED,
              "wgrad: dz must be a dense NHWC tensor");
  int BN = 0;
  for (int c = 256; c >= 16; c >>= 1)
    if (J.small_.c % c == 0) { BN = c; break; }
  SEG_REQUIRE(BN != 0, SEG_E_UNSUPPORTED, "wgrad: N=%d not a multiple of 16", J.small_.c);
  switch (AW) {
    case 64: return launch_wgrad_bn<64>(J, BN, st);
    case 32: return launch_wgrad_bn<32>(J, BN, st);
    default: return launch_wgrad_bn<16>(J, BN, st);
  }
}


// ---------------------------------------------------------------------------
// halo-tile convolution (hconv.cuh): launch
// ---------------------------------------------------------------------------
struct HconvJob {
  seg_view a1, a2;
  int kh, kw;
  int pad_t, pad_l;            // padded-position (0,0) = tensor coordinate (-pad_t, -pad_l)
  int Hp, Wp_logical;          // padded grid per image
  int Ho, Wo, batch;
  const void* w;
  int w_rows, w_cols;
  bool b_mn;
  int b_rows_per_tap;
  bool tap_flip;
  int N_total, max_bn;
  EpiDest d0, d1;
  int split_n;
  const float* bias;
  int flags;
  const int* tap_rows;         // optional: weight-matrix row of tap t = r*kw+s (see HconvParams)
  const float* post_scale;     // optional affine map after bias / ReLU (see HconvParams)
  const float* post_shift;
};

static long long* g_prof_buf = nullptr;     // in-kernel timeline buffer (test hook)
void hconv_set_prof(void* p) { g_prof_buf = reinterpret_cast<long long*>(p); }
static int g_deep_b_ring = 1;       // seg_set_option key 12: streamed-B rings as deep as smem allows
void conv_set_deep_b_ring(int on) { g_deep_b_ring = on != 0; }
static int g_hconv_row_align = 0;   // 0: natural (128-byte) row alignment, 8: pad rows to 8 px

template <int KC, int BN, bool B_MN, int TPS = 1>
static int launch_hconv_t(const HconvJob& J, const HconvParams& P0, int smem_bytes,
                          cudaStream_t st) {
  static int attr_smem = 0;
  if (attr_smem < smem_bytes) {
    SEG_CHECK_CUDA(cudaFuncSetAttribute(hconv_kernel<KC, BN, B_MN, TPS>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_smem = 227 * 1024;
  }
  constexpr int kAtomN = BN < 64 ? BN : 64;
  HconvParams P = P0;
  CUtensorMap tmA1, tmA2, tmB;
  int rc;
  if (P.flat) {
    rc = make_tmap_2d(&tmA1, J.a1.ptr, J.a1.c, (int64_t)J.batch * P.Hp * P.Wp, J.a1.sw, KC,
                      P.box_rows, KC * 2);
    if (rc) return rc;
    if (J.a2.ptr) {
      rc = make_tmap_2d(&tmA2, J.a2.ptr, J.a2.c, (int64_t)J.batch * P.Hp * P.Wp, J.a2.sw, KC,
                        P.box_rows, KC * 2);
      if (rc) return rc;
    } else {
      tmA2 = tmA1;
    }
  } else {
    rc = make_tmap_rows(&tmA1, J.a1, KC, P.row_px, KC * 2);
    if (rc) return rc;
    if (J.a2.ptr) {
      rc = make_tmap_rows(&tmA2, J.a2, KC, P.row_px, KC * 2);
      if (rc) return rc;
    } else {
      tmA2 = tmA1;
    }
  }
  if (B_MN)
    rc = make_tmap_2d(&tmB, J.w, J.w_cols, J.w_rows, J.w_cols, kAtomN, KC, kAtomN * 2);
  else
    rc = make_tmap_2d(&tmB, J.w, J.w_cols, J.w_rows, J.w_cols, KC, BN, KC * 2);
  if (rc) return rc;
  const int m_tiles = (P.P_total + kBlockM - 1) / kBlockM;
  const int n_tiles = J.N_total / BN;
  const int tiles = m_tiles * n_tiles;
  int grid = tiles < num_sms() ? tiles : num_sms();
  if (P.b_resident && n_tiles > 1) {
    grid -= grid % n_tiles;           // every CTA must keep one N-slice for its lifetime
    if (grid < n_tiles) P.b_resident = 0, grid = tiles < num_sms() ? tiles : num_sms();
  }
  SEG_CHECK_CUDA(launch_k(hconv_kernel<KC, BN, B_MN, TPS>, dim3(grid), dim3(kConvThreads), (size_t)(smem_bytes), st, tmA1, tmA2, tmB, P));
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

static int g_hconv_rowstage = 1;       // seg_set_option key 14: one filter row per weight stage
void hconv_set_rowstage(int on) { g_hconv_rowstage = on != 0; }

template <int KC, bool B_MN>
static int launch_hconv_bn(const HconvJob& J, const HconvParams& P, int BN, int smem,
                           cudaStream_t st, int tps = 1) {
  if (tps == 3) {                      // planned only for KC = 64, BN = 64 / 128
    if (KC != 64) return SEG_E_UNSUPPORTED;
    switch (BN) {
      case 128: return launch_hconv_t<64, 128, B_MN, 3>(J, P, smem, st);
      case 64: return launch_hconv_t<64, 64, B_MN, 3>(J, P, smem, st);
    }
    return SEG_E_UNSUPPORTED;
  }
  switch (BN) {
    case 256: return launch_hconv_t<KC, 256, B_MN>(J, P, smem, st);
    case 128: return launch_hconv_t<KC, 128, B_MN>(J, P, smem, st);
    case 64: return launch_hconv_t<KC, 64, B_MN>(J, P, smem, st);
    case 32: return launch_hconv_t<KC, 32, B_MN>(J, P, smem, st);
    case 16: return launch_hconv_t<KC, 16, B_MN>(J, P, smem, st);
  }
  return SEG_E_UNSUPPORTED;
}

// Returns SEG_E_UNSUPPORTED (without setting up anything) when the shape does not fit the
// halo kernel; the caller then uses the im2col kernel.
static int launch_hconv(const HconvJob& J, cudaStream_t st) {
  int rc = load_encoders();
  if (rc) return rc;
  const int c2 = J.a2.ptr ? J.a2.c : 0;
  const int KC = pick_chunk(J.a1.c, c2);
  if (KC == 0) return SEG_E_UNSUPPORTED;
  const int SWZ = KC * 2;
  int BN = 0;
  for (int c = 256; c >= 16; c >>= 1)
    if (J.max_bn % c == 0) { BN = c; break; }
  if (BN == 0 || J.N_total % BN) return SEG_E_UNSUPPORTED;

  HconvParams P;
  memset(&P, 0, sizeof(P));
  const bool dense = view_dense(J.a1) && (!J.a2.ptr || view_dense(J.a2));
  const bool nopad = J.pad_t == 0 && J.pad_l == 0 && J.Hp == J.a1.h && J.Wp_logical == J.a1.w;
  P.flat = (dense && nopad) ? 1 : 0;
  P.row_px = J.Wp_logical;
  if (!P.flat) {
    if (J.Wp_logical > 256) return SEG_E_UNSUPPORTED;
    int align = 128 / SWZ;                       // TMA smem destinations: 128-byte aligned
    if (g_hconv_row_align > align) align = g_hconv_row_align;
    P.Wp = ((J.Wp_logical + align - 1) / align) * align;
  } else {
    P.Wp = J.Wp_logical;
  }
  // staging geometry of one tile of mt x 128 positions (+ halo); false: does not fit
  // (the kernel runs mt = 1: a two-accumulator variant sharing one B stream was built and
  // measured neutral at best, and its loop overhead cost every halo launch ~10 %)
  auto geom = [&](int mt, int* box_rows, int* nboxes, int* stage_bytes, int* live_bytes) {
    const int tile_m = kBlockM * mt;
    int bytes;
    if (P.flat) {
      const int need = tile_m + (J.kh - 1) * P.Wp + J.kw - 1;
      *box_rows = need <= 256 ? ((need + 7) / 8) * 8 : 128;
      *nboxes = (need + *box_rows - 1) / *box_rows;
      bytes = *nboxes * *box_rows * SWZ;
      *live_bytes = bytes;
    } else {
      *box_rows = 0; *nboxes = 0;
      const int nrows = (P.Wp - 1 + tile_m - 1 + (J.kh - 1) * P.Wp + J.kw - 1) / P.Wp + 1;
      bytes = nrows * P.Wp * SWZ;
      *live_bytes = nrows * J.Wp_logical * SWZ;
    }
    *stage_bytes = ((bytes + 1023) / 1024) * 1024;
    return *stage_bytes <= 72 * 1024;
  };
  int live1 = 0;
  if (!geom(1, &P.box_rows, &P.nboxes, &P.a_stage_bytes, &live1)) return SEG_E_UNSUPPORTED;
  P.Hp = J.Hp; P.batch = J.batch; P.Ho = J.Ho; P.Wo = J.Wo;
  P.P_total = J.batch * J.Hp * P.Wp;
  P.kh = J.kh; P.kw = J.kw;
  P.pad_t = J.pad_t; P.pad_l = J.pad_l;
  P.chunks1 = J.a1.c / KC;
  P.chunks2 = c2 / KC;
  P.tap_flip = J.tap_flip ? 1 : 0;
  P.b_rows_per_tap = J.b_rows_per_tap;
  if (J.tap_rows) {
    if (J.kh * J.kw > 25) return SEG_E_UNSUPPORTED;
    P.use_tap_rows = 1;
    for (int t = 0; t < J.kh * J.kw; ++t) P.tap_rows[t] = J.tap_rows[t];
  }
  P.N_total = J.N_total;
  P.d0 = J.d0; P.d1 = J.d1; P.split_n = J.split_n;
  P.bias = J.bias; P.flags = J.flags;
  P.post_scale = J.post_scale; P.post_shift = J.post_shift;
  P.prof = g_prof_buf;

  // shared-memory budget: B resident if every (chunk, tap) tile of one N-slice fits next
  // to >= 2 A stages, else a B ring of a few stages.
  const int budget = 222 * 1024 - 2048;
  const int taps = J.kh * J.kw;
  const int chunks = P.chunks1 + P.chunks2;
  auto plan = [&](int bn, int* sa, int* sb, int* resident) {
    const int bbytes = bn * KC * 2;
    const int all_b = taps * chunks * bbytes;
    if (taps * chunks <= kHconvMaxSB && all_b + 2 * P.a_stage_bytes <= budget) {
      *resident = 1;
      *sb = taps * chunks;
    } else {
      *resident = 0;
      int s = (64 * 1024) / bbytes;
      *sb = s < 2 ? 2 : (s > 12 ? 12 : s);
      if (g_deep_b_ring) {
        // B is the bulk of what a tile ingests (taps x more than A) and every B stage is
        // re-armed only after the MMA that read it retires: the ring depth, not bandwidth,
        // sets the ingest rate (measured ~25 B/clk at 64 KB in flight).  Give A what one
        // tile keeps busy (<= 3 stages) and B the rest.
        const int a_keep = chunks < 3 ? (chunks < 2 ? 2 : chunks) : 3;
        int deep = (budget - a_keep * P.a_stage_bytes) / bbytes;
        if (deep > kHconvMaxSB) deep = kHconvMaxSB;
        if (deep > taps * chunks) deep = taps * chunks;
        if (deep > *sb) *sb = deep;
      }
    }
    int a = (budget - *sb * bbytes) / P.a_stage_bytes;
    *sa = a > kHconvMaxSA ? kHconvMaxSA : a;
    return *sa >= 1;
  };
  int SA = 0, SB = 0, res = 0;
  while (!plan(BN, &SA, &SB, &res) || (SA < 2 && BN > 32)) {
    if (BN <= 16) return SEG_E_UNSUPPORTED;
    BN >>= 1;
  }
  // fill the machine: halve BN while there are fewer tiles than SMs (not below 64)
  const int64_t m_tiles = ceil_div64(P.P_total, kBlockM);
  while (BN > 64 && m_tiles * (J.N_total / BN) < num_sms()) {
    BN >>= 1;
    plan(BN, &SA, &SB, &res);
  }
  P.SA = SA; P.SB = SB; P.b_resident = res;
  const int smem = SA * P.a_stage_bytes + SB * BN * KC * 2 + 2048;
  static const bool dbg = getenv("SEGB200_DEBUG_PLAN") != nullptr;
  if (dbg)
    fprintf(stderr, "hconv plan: P=%d N=%d k=%dx%d chunks=%d KC=%d flat=%d Wp=%d -> BN=%d SA=%d "
            "SB=%d res=%d stage=%d tiles=%lld\n", P.P_total, J.N_total, J.kh, J.kw, chunks, KC, P.flat,
            P.Wp, BN, SA, SB, res, P.a_stage_bytes,
            (long long)(ceil_div64(P.P_total, (int64_t)kBlockM) * (J.N_total / BN)));
  // one filter row (3 taps) per weight stage: a third of the barrier round trips in the
  // issue loop; the ring is re-planned in row units
  if (g_hconv_rowstage && !res && KC == 64 && (BN == 64 || BN == 128) && J.kw == 3 &&
      !J.tap_rows) {                   // (sub-kernels with a tap table: not yet validated)
    const int stage = 3 * BN * KC * 2;
    int sb = (budget - 3 * P.a_stage_bytes) / stage;
    if (sb > kHconvMaxSB) sb = kHconvMaxSB;
    if (sb > J.kh * chunks) sb = J.kh * chunks;
    int sa = sb >= 2 ? (budget - sb * stage) / P.a_stage_bytes : 0;
    if (sa > kHconvMaxSA) sa = kHconvMaxSA;
    if (sb >= 2 && sa >= 2) {
      P.SA = sa; P.SB = sb;
      const int smem3 = sa * P.a_stage_bytes + sb * stage + 2048;
      return J.b_mn ? launch_hconv_bn<64, true>(J, P, BN, smem3, st, 3)
                    : launch_hconv_bn<64, false>(J, P, BN, smem3, st, 3);
    }
  }
  if (J.b_mn) {
    switch (KC) {
      case 64: return launch_hconv_bn<64, true>(J, P, BN, smem, st);
      case 32: return launch_hconv_bn<32, true>(J, P, BN, smem, st);
      default: return launch_hconv_bn<16, true>(J, P, BN, smem, st);
    }
  }
  switch (KC) {
    case 64: return launch_hconv_bn<64, false>(J, P, BN, smem, st);
    case 32: return launch_hconv_bn<32, false>(J, P, BN, smem, st);
    default: return launch_hconv_bn<16, false>(J, P, BN, smem, st);
  }
}


// ---------------------------------------------------------------------------
// spatial-tile convolution (tconv.cuh): launch
// ---------------------------------------------------------------------------
// tiled 4-D map over an NHWC view seen as (C, W, H, N) with an arbitrary box
static int make_tmap_box(CUtensorMap* tm, const seg_view& v, int box_c, int box_w, int box_h,
                         int swizzle_bytes) {
  SEG_REQUIRE((reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0 && (v.sw * 2) % 16 == 0 &&
                  (v.sh * 2) % 16 == 0 && (v.sn * 2) % 16 == 0,
              SEG_E_ALIGN, "box tensor map: view must be 16-byte aligned in every stride");
  cuuint64_t gdim[4] = {(cuuint64_t)v.c, (cuuint64_t)v.w, (cuuint64_t)v.h, (cuuint64_t)v.n};
  cuuint64_t gstr[3] = {(cuuint64_t)v.sw * 2, (cuuint64_t)v.sh * 2, (cuuint64_t)v.sn * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t est[4] = {1, 1, 1, 1};
  CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, v.ptr, gdim, gstr, box, est,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(swizzle_bytes),
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SEG_REQUIRE(r == CUDA_SUCCESS, SEG_E_CUDA,
              "cuTensorMapEncodeTiled(4d box) failed (%d) c=%d w=%d h=%d n=%d box=%dx%dx%d sw=%d",
              (int)r, v.c, v.w, v.h, v.n, box_c, box_w, box_h, swizzle_bytes);
  return SEG_OK;
}

struct TconvJob {
  seg_view a1, a2;             // input sources (a2.ptr == null: none)
  int pad_t, pad_l;
  seg_view d0, d1;             // destinations (d1.ptr == null: none); d0.h x d0.w = output grid
  seg_view m0, m1;             // ReLU-grad mask sources (same geometry as d0 / d1), nullable
  const void* w;
  int w_rows, w_cols;
  bool b_mn;
  int b_rows_per_tap;
  bool tap_flip;
  int N_total, max_bn;
  int split_n;
  const float* bias;
  int flags;
};

static int g_tconv_min_eff = 65;     // percent of computed output pixels that must be useful (swept: 55 / 60 / 65 / 70 / 85 -> 0.951 / 0.940 / 0.937 / 0.945 / 0.956 ms per U-Net step)
static bool g_use_tconv = true;
void tconv_enable(int on) { g_use_tconv = on != 0; }
void tconv_set_min_eff(int pct) { g_tconv_min_eff = pct; }

struct TconvPlan {
  int KC, BN, MT, SA, SB, resident, smem;
  TconvParams P;
};

template <int KC, int BN, bool B_MN, int MT>
static int launch_tconv_t(const TconvJob& J, const TconvPlan& L, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    SEG_CHECK_CUDA(cudaFuncSetAttribute(tconv_kernel<KC, BN, B_MN, MT>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  constexpr int kAtomN = BN < 64 ? BN : 64;
  constexpr int BNH = 32;               // channels per epilogue store / mask box
  constexpr int PW = 8 * MT + 2, PH = kTconvTH + 2;
  CUtensorMap tmA1, tmA2, tmB, tmD0, tmD1, tmM0, tmM1;
  int rc = make_tmap_box(&tmA1, J.a1, KC, PW, PH, KC * 2);
  if (rc) return rc;
  if (J.a2.ptr) {
    rc = make_tmap_box(&tmA2, J.a2, KC, PW, PH, KC * 2);
    if (rc) return rc;
  } else {
    tmA2 = tmA1;
  }
  if (B_MN)
    rc = make_tmap_2d(&tmB, J.w, J.w_cols, J.w_rows, J.w_cols, kAtomN, KC, kAtomN * 2);
  else
    rc = make_tmap_2d(&tmB, J.w, J.w_cols, J.w_rows, J.w_cols, KC, BN, KC * 2);
  if (rc) return rc;
  rc = make_tmap_box(&tmD0, J.d0, BNH, 8, 4, BNH * 2);
  if (rc) return rc;
  if (J.d1.ptr) {
    rc = make_tmap_box(&tmD1, J.d1, BNH, 8, 4, BNH * 2);
    if (rc) return rc;
  } else {
    tmD1 = tmD0;
  }
  tmM0 = tmD0;
  tmM1 = tmD1;
  if (J.flags & SEG_EPI_RELU_MASK) {
    rc = make_tmap_box(&tmM0, J.m0, BNH, 8, 4, BNH * 2);
    if (rc) return rc;
    if (J.d1.ptr) {
      rc = make_tmap_box(&tmM1, J.m1, BNH, 8, 4, BNH * 2);
      if (rc) return rc;
    } else {
      tmM1 = tmM0;
    }
  }
  const TconvParams& P = L.P;
  const int tiles = P.batch * P.tiles_y * P.tiles_x * P.n_tiles;
  int grid = tiles < num_sms() ? tiles : num_sms();
  if (P.b_resident && P.n_tiles > 1) grid -= grid % P.n_tiles;   // fixed N-slice per CTA
  SEG_CHECK_CUDA(launch_k(tconv_kernel<KC, BN, B_MN, MT>, dim3(grid), dim3(kTconvThreads), (size_t)(L.smem), st, tmA1, tmA2, tmB, tmD0, tmD1, tmM0, tmM1, P));
  SEG_LAUNCH_CHECK();
  return SEG_OK;
}

template <int KC, bool B_MN>
static int launch_tconv_bn(const TconvJob& J, const TconvPlan& L, cudaStream_t st) {
  if (L.MT == 2) {
    switch (L.BN) {
      case 128: return launch_tconv_t<KC, 128, B_MN, 2>(J, L, st);
      case 64: return launch_tconv_t<KC, 64, B_MN, 2>(J, L, st);
      case 32: return launch_tconv_t<KC, 32, B_MN, 2>(J, L, st);
    }
  } else {
    switch (L.BN) {
      case 128: return launch_tconv_t<KC, 128, B_MN, 1>(J, L, st);
      case 64: return launch_tconv_t<KC, 64, B_MN, 1>(J, L, st);
      case 32: return launch_tconv_t<KC, 32, B_MN, 1>(J, L, st);
    }
  }
  return SEG_E_UNSUPPORTED;
}

// smem plan for (KC, BN, MT); false if it does not fit
static bool tconv_plan(const TconvJob& J, int KC, int BN, int MT, int chunks, TconvPlan* L) {
  const int budget = 227 * 1024 - 1024;                 // minus base-alignment slack
  const int a_stage = (((8 * MT + 2) * (kTconvTH + 2) * KC * 2) + 1023) / 1024 * 1024;
  const int bbytes = BN * KC * 2;
  const int box = 32 * 32 * 2;                          // one [4][8][32] staging / mask box
  const int nld = MT * (BN / 32);
  const int ew = 4 * kTconvEW;                          // epilogue warps
  const int my_max = (nld + kTconvEW - 1) / kTconvEW;   // chunks per warp per tile
  const int msk = (J.flags & SEG_EPI_RELU_MASK) ? ew * my_max * box : 0;
  // staging boxes per epilogue warp: a box is reused only after the TMA store that read it
  // has left shared memory (~1000 cycles behind its commit): at least one per chunk the
  // warp drains per tile, two if that still leaves room for a resident B next to 3 A stages
  int nstg = my_max > 2 ? my_max : 2;
  const int nstg_min = my_max > 1 ? my_max : 1;
  while (nstg > nstg_min && 9 * chunks * bbytes + 3 * a_stage + ew * nstg * box + msk + 3072 > budget) --nstg;
  const int stg = ew * nstg * box;
  const int fixed = stg + msk + 2048 /*bias*/ + 1024 /*barriers*/;
  int SB, resident;
  if (9 * chunks <= kTconvMaxSB && 9 * chunks * bbytes + 2 * a_stage + fixed <= budget) {
    resident = 1;
    SB = 9 * chunks;
  } else {
    resident = 0;
    SB = (48 * 1024) / bbytes;
    SB = SB < 3 ? 3 : (SB > 9 ? 9 : SB);
    if (g_deep_b_ring) {                                  // see launch_hconv's plan()
      int deep = (budget - fixed - 3 * a_stage) / bbytes;
      if (deep > kTconvMaxSB) deep = kTconvMaxSB;
      if (deep > 9 * chunks) deep = 9 * chunks;
      if (deep > SB) SB = deep;
    }
  }
  int SA = (budget - fixed - SB * bbytes) / a_stage;
  if (SA > kTconvMaxSA) SA = kTconvMaxSA;
  if (SA > 2 * chunks + 2) SA = 2 * chunks + 2;         // more stages than two tiles is waste
  if (SA < 2) return false;
  L->KC = KC; L->BN = BN; L->MT = MT; L->SA = SA; L->SB = SB; L->resident = resident;
  TconvParams& P = L->P;
  P.SA = SA; P.SB = SB; P.b_resident = resident;
  P.nstg = nstg;
  P.a_stage_bytes = a_stage;
  P.off_b = SA * a_stage;
  P.off_stage = P.off_b + ((SB * bbytes + 1023) / 1024) * 1024;
  P.off_mask = P.off_stage + stg;
  P.off_bias = P.off_mask + msk;
  P.off_bars = P.off_bias + 2048;
  L->smem = P.off_bars + 1024 + 1024;
  return L->smem <= 227 * 1024;
}

// Returns SEG_E_UNSUPPORTED (nothing launched) when the shape does not suit this kernel.
static int launch_tconv(const TconvJob& J, cudaStream_t st) {
  int rc = load_encoders();
  if (rc) return rc;
  if (J.flags & SEG_EPI_OUT_F32) return SEG_E_UNSUPPORTED;
  const int c2 = J.a2.ptr ? J.a2.c : 0;
  const int KC = pick_chunk(J.a1.c, c2);
  if (KC == 0 || J.N_total > 512) return SEG_E_UNSUPPORTED;
  int BN = 0;
  for (int c = 128; c >= 32; c >>= 1)
    if (J.max_bn % c == 0) { BN = c; break; }
  if (BN == 0 || J.N_total % BN) return SEG_E_UNSUPPORTED;
  const int Ho = J.d0.h, Wo = J.d0.w, batch = J.d0.n;
  const int tiles_y = (Ho + kTconvTH - 1) / kTconvTH;
  // column-block count: the one wasting fewer columns, wider on ties
  const int w1 = (Wo + 7) / 8 * 8, w2 = (Wo + 15) / 16 * 16;
  int MT = w2 <= w1 ? 2 : 1;
  const int wcomp = MT == 2 ? w2 : w1;
  if ((int64_t)Ho * Wo * 100 < (int64_t)g_tconv_min_eff * tiles_y * kTconvTH * wcomp)
    return SEG_E_UNSUPPORTED;
  const int chunks = (J.a1.c + c2) / KC;
  TconvPlan L;
  memset(&L, 0, sizeof(L));
  // The widest N-slice the layer allows, even when its weights then stream through a ring
  // instead of staying resident: a 128 x 128 x 16 MMA runs at the tensor pipe's rate, two
  // 128 x 64 x 16 ones take 1.5x as long (operand-read bound), and halving BN also re-reads
  // the activation tile once more.  Measured against "shrink BN until the weights are
  // resident" (round 1's rule): U-Net step 0.944 -> 0.937 ms, FCN-8s config 2 1.270 ->
  // 1.196 ms (its 128- and 256-channel layers), profiles/r02_experiments.md #26.
  bool ok = tconv_plan(J, KC, BN, MT, chunks, &L);
  while (!ok) {
    if (MT == 2) MT = 1;
    else if (BN > 32) BN >>= 1;
    else return SEG_E_UNSUPPORTED;
    ok = tconv_plan(J, KC, BN, MT, chunks, &L);
  }
  MT = L.MT; BN = L.BN;
  // Fill the machine: a layer with few pixels and many channels gives fewer tiles than
  // SMs.  Narrower column blocks / N-slices multiply the tile count; estimate each
  // candidate's time as waves x MMA cycles per tile (measured: 64 / 48 / 40 cycles per
  // 128 x BN x 16 MMA for BN = 128 / 64 / 32) and keep the cheapest.
  {
    auto tiles_of = [&](int mt, int bn) {
      return (int64_t)batch * tiles_y * ((Wo + 8 * mt - 1) / (8 * mt)) * (J.N_total / bn);
    };
    auto cost_of = [&](const TconvPlan& C) {
      const int64_t t = tiles_of(C.MT, C.BN);
      const int64_t waves = (t + num_sms() - 1) / num_sms();
      const int mma = C.BN >= 128 ? 64 : (C.BN == 64 ? 48 : 40);
      double c = (double)waves * (C.MT * mma * 9.0 * chunks * (KC / 16) + C.MT * (C.BN / 32) * 150.0);
      return C.resident ? c : c * 1.25;
    };
    if (tiles_of(MT, BN) < num_sms()) {
      double best = cost_of(L);
      for (int mt = MT; mt >= 1; --mt)
        for (int bn = BN; bn >= 32; bn >>= 1) {
          if (mt == MT && bn == BN) continue;
          TconvPlan C;
          memset(&C, 0, sizeof(C));
          if (J.N_total % bn || !tconv_plan(J, KC, bn, mt, chunks, &C)) continue;
          const double c = cost_of(C);
          if (c < best * 0.9) { best = c; L = C; }
        }
      MT = L.MT; BN = L.BN;
    }
  }
  TconvParams& P = L.P;
  const int TW = 8 * MT;
  P.tiles_x = (Wo + TW - 1) / TW;
  P.tiles_y = tiles_y;
  P.batch = batch;
  P.n_tiles = J.N_total / BN;
  P.chunks1 = J.a1.c / KC;
  P.chunks2 = c2 / KC;
  P.pad_t = J.pad_t; P.pad_l = J.pad_l;
  P.tap_flip = J.tap_flip ? 1 : 0;
  P.b_rows_per_tap = J.b_rows_per_tap;
  P.split_n = J.d1.ptr ? J.split_n : 0;
  P.bias = J.bias;
  P.bias_cols = J.d0.c;
  P.n_total = J.N_total;
  P.flags = J.flags;
  P.prof = g_prof_buf;
  if (g_prof_buf)
    fprintf(stderr, "tconv plan: KC=%d BN=%d MT=%d SA=%d SB=%d resident=%d nstg=%d smem=%d tiles=%dx%dx%dx%d\n",
            KC, BN, MT, L.SA, L.SB, L.resident, P.nstg, L.smem, batch, P.tiles_y, P.tiles_x, P.n_tiles);
  if (J.b_mn) {
    switch (KC) {
      case 64: return launch_tconv_bn<64, true>(J, L, st);
      case 32: return launch_tconv_bn<32, true>(J, L, st);
      default: return launch_tconv_bn<16, true>(J, L, st);
    }
  }
  switch (KC) {
    case 64: return launch_tconv_bn<64, false>(J, L, st);
    case 32: return launch_tconv_bn<32, false>(J, L, st);
    default: return launch_tconv_bn<16, false>(J, L, st);
  }
}


// ---------------------------------------------------------------------------
// spatial-tile weight gradient (twgrad.cuh): launch
// ---------------------------------------------------------------------------
static bool g_use_twgrad = true;
static int g_twgrad_min_eff = 40;
static int g_twgrad_min_tiles = 8;   // pixel tiles per CTA below which the grid is narrowed
void twgrad_set_min_tiles(int n) { g_twgrad_min_tiles = n < 1 ? 1 : n; }
void twgrad_enable(int on) { g_use_twgrad = on != 0; }
void twgrad_set_min_eff(int pct) { g_twgrad_min_eff = pct; }

static int g_twgrad_tred = 1;          // seg_set_option key 15: TMA tensor reduce-add epilogue
void twgrad_set_tred(int on) { g_twgrad_tred = on != 0; }

// dW [9][BC][SC] fp32 as a 3-D tensor {SC, BC, 9} with 128-byte-swizzled boxes {32, rows, 1}
static int make_tmap_dw(CUtensorMap* tm, float* dw, int SC, int BC, int box_rows) {
  SEG_REQUIRE((reinterpret_cast<uintptr_t>(dw) & 15) == 0 && SC % 4 == 0, SEG_E_ALIGN,
              "dW tensor map: base and row pitch must be 16-byte aligned");
  cuuint64_t gdim[3] = {(cuuint64_t)SC, (cuuint64_t)BC, 9};
  cuuint64_t gstr[2] = {(cuuint64_t)SC * 4, (cuuint64_t)BC * SC * 4};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
  cuuint32_t est[3] = {1, 1, 1};
  CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dw, gdim, gstr, box, est,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SEG_REQUIRE(r == CUDA_SUCCESS, SEG_E_CUDA, "cuTensorMapEncodeTiled(dW) failed (%d) SC=%d BC=%d",
              (int)r, SC, BC);
  return SEG_OK;
}

template <int AW, int BN, bool TRED = false>
static int launch_twgrad_t(const WgradJob& J, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    SEG_CHECK_CUDA(cudaFuncSetAttribute(twgrad_kernel<AW, BN, TRED>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  constexpr int kAtomN = BN < 64 ? BN : 64;
  CUtensorMap tmX1, tmX2, tmZ;
  int rc = make_tmap_box(&tmX1, J.big, AW, kTwPW, kTwPH, AW * 2);
  if (rc) return rc;
  if (J.big2.ptr) {
    rc = make_tmap_box(&tmX2, J.big2, AW, kTwPW, kTwPH, AW * 2);
    if (rc) return rc;
  } else {
    tmX2 = tmX1;
  }
  rc = make_tmap_box(&tmZ, J.small_, kAtomN, kTwTW, kTwTH, kAtomN * 2);
  if (rc) return rc;
  TwgradParams P;
  memset(&P, 0, sizeof(P));
  P.tiles_x = (J.small_.w + kTwTW - 1) / kTwTW;
  P.tiles_y = (J.small_.h + kTwTH - 1) / kTwTH;
  P.batch = J.small_.n;
  P.chunks1 = J.big.c / AW;
  P.chunks2 = J.big2.ptr ? J.big2.c / AW : 0;
  P.n_slices = J.small_.c / BN;
  P.pad_t = -J.low_h; P.pad_l = -J.low_w;
  P.BC = J.BC; P.SC = J.SC;
  P.dw = J.dw; P.db = J.db;
  const int combos = (P.chunks1 + P.chunks2) * P.n_slices;
  const int tiles = P.batch * P.tiles_y * P.tiles_x;
  int per = num_sms() / combos;
  if (per < 1) per = 1;
  if (per > tiles) per = tiles;
  // Every CTA ends with a full-size partial sum for the global reduction (~25k cycles of
  // epilogue and L2 reduction traffic that does not shrink with its share of the pixels),
  // and these kernels run beside the input-gradient chain on a second stream: give a CTA
  // at least g_twgrad_min_tiles pixel tiles, leaving the other SMs to the main stream.
  if (per > 1 && tiles / per < g_twgrad_min_tiles) {
    per = tiles / g_twgrad_min_tiles;
    if (per < 1) per = 1;
  }
  P.x_bytes = ((kTwPW * kTwPH * AW * 2) + 1023) / 1024 * 1024;
  P.stage_bytes = P.x_bytes + kTwTH * kTwTW * BN * 2;
  int stages = (227 * 1024 - 2048 - 1024) / P.stage_bytes;
  if (stages > kTwMaxStages) stages = kTwMaxStages;
  P.stages = stages;
  P.off_bars = stages * P.stage_bytes;
  const int smem = P.off_bars + 1024 + 1024;
  constexpr int kCols = (AW == 64 ? 5 : 3) * BN;
  // the idle pipeline stages hold the staged accumulators of the epilogue
  P.staged_ok = 128 * (kCols + 4) * 4 <= stages * P.stage_bytes ? 1 : 0;
  P.ctas_per_combo = per;
  CUtensorMap tmDW = tmZ;              // unused unless TRED
  if (TRED) {
    if (!P.staged_ok) return SEG_E_UNSUPPORTED;
    rc = make_tmap_dw(&tmDW, J.dw, J.SC, J.BC, AW);
    if (rc) return rc;
  }
  SEG_CHECK_CUDA(launch_kc(twgrad_kernel<AW, BN, TRED>, dim3(combos * per), dim3(kConvThreads), (size_t)(smem), st, 1, tmX1, tmX2, tmZ, tmDW, P));
  return SEG_OK;
}

// Returns SEG_E_UNSUPPORTED (nothing launched) when the shape does not suit this kernel.
static int launch_twgrad(const WgradJob& J, cudaStream_t st) {
  if (!(J.kh == 3 && J.kw == 3 && J.stride == 1)) return SEG_E_UNSUPPORTED;
  int rc = load_encoders();
  if (rc) return rc;
  const int AW = pick_chunk(J.big.c, J.big2.ptr ? J.big2.c : 0);
  if (AW == 0) return SEG_E_UNSUPPORTED;
  const int co = J.small_.c;
  const int BN = co % 64 == 0 ? 64 : (co % 32 == 0 ? 32 : (co % 16 == 0 ? 16 : 0));
  if (BN == 0) return SEG_E_UNSUPPORTED;
  const int Ho = J.small_.h, Wo = J.small_.w;
  const int64_t comp = (int64_t)((Ho + kTwTH - 1) / kTwTH * kTwTH) * ((Wo + kTwTW - 1) / kTwTW * kTwTW);
  if ((int64_t)Ho * Wo * 100 < (int64_t)g_twgrad_min_eff * comp) return SEG_E_UNSUPPORTED;
  // TMA tensor reduce-add epilogue (option 15): whole BN slices, 16-byte aligned rows
  if (g_twgrad_tred && BN >= 32 && J.SC % BN == 0 &&
      (reinterpret_cast<uintptr_t>(J.dw) & 15) == 0) {
    int rc2 = SEG_E_UNSUPPORTED;
    if (AW == 64 && BN == 64) rc2 = launch_twgrad_t<64, 64, true>(J, st);
    else if (AW == 64 && BN == 32) rc2 = launch_twgrad_t<64, 32, true>(J, st);
    else if (AW == 32 && BN == 64) rc2 = launch_twgrad_t<32, 64, true>(J, st);
    else if (AW == 32 && BN == 32) rc2 = launch_twgrad_t<32, 32, true>(J, st);
    else if (AW == 16 && BN == 32) rc2 = launch_twgrad_t<16, 32, true>(J, st);
    if (rc2 != SEG_E_UNSUPPORTED) return rc2;
  }
  switch (AW) {
    case 64:
      switch (BN) {
        case 64: return launch_twgrad_t<64, 64>(J, st);
        case 32: return launch_twgrad_t<64, 32>(J, st);
        default: return launch_twgrad_t<64, 16>(J, st);
      }
    case 32:
      switch (BN) {
        case 64: return launch_twgrad_t<32, 64>(J, st);
        case 32: return launch_twgrad_t<32, 32>(J, st);
        default: return launch_twgrad_t<32, 16>(J, st);
      }
    default:
      switch (BN) {
        case 64: return launch_twgrad_t<16, 64>(J, st);
        case 32: return launch_twgrad_t<16, 32>(J, st);
        default: return launch_twgrad_t<16, 16>(J, st);
      }
  }
}

void hconv_set_row_align(int a) { g_hconv_row_align = a; }

static bool g_use_hconv = true;
void hconv_enable(int on) { g_use_hconv = on != 0; }

// ---------------------------------------------------------------------------
// op-level entry points used by api.cu
// ---------------------------------------------------------------------------
// fconv.cu: first-layer kernel; SEG_E_UNSUPPORTED = not its shape, nothing launched
int fconv_fwd(const seg_conv_desc& d, const seg_view& x, const seg_view* x2, const void* w,
              const float* bias, const seg_view& y, cudaStream_t st);
int fconv_wgrad(const seg_conv_desc& d, const seg_view& x, const seg_view* x2,
                const seg_view& dz, float* dw, float* db, cudaStream_t st);

int umma_conv_fwd(const seg_conv_desc& d, const seg_view& x, const seg_view* x2, const void* w,
                  const float* bias, const seg_view& y, cudaStream_t st, const float* post_scale,
                  const float* post_shift) {
  if (x.c == 4 && d.cin == 3) {
    const int rc = fconv_fwd(d, x, x2, w, bias, y, st);
    if (rc != SEG_E_UNSUPPORTED) return rc;
    SEG_REQUIRE(false, SEG_E_UNSUPPORTED,
                "conv_fwd: a 4-channel (R,G,B,1) input needs the first-layer kernel's shape "
                "(3x3, stride 1, dense tensors, 32 or 64 padded output channels)");
  }
  IgemmJob J;
  memset(&J, 0, sizeof(J));
  J.a1 = x;
  J.a2 = x2 ? *x2 : null_view();
  SEG_REQUIRE(x.c + J.a2.c == d.cin_pad, SEG_E_BAD_SHAPE,
              "conv_fwd: input channels %d+%d != cin_pad %d", x.c, J.a2.c, d.cin_pad);
  J.kh = d.kh; J.kw = d.kw; J.stride = d.stride;
  J.low_h = -d.pad_t; J.low_w = -d.pad_l;
  J.up_h = d.pad_b - (d.kh - 1); J.up_w = d.pad_r - (d.kw - 1);
  J.Ho = y.h; J.Wo = y.w; J.batch = y.n;
  J.w = w; J.w_rows = d.kh * d.kw * d.cin_pad; J.w_cols = d.cout_pad;
  J.b_mn = true; J.b_rows_per_tap = d.cin_pad; J.tap_flip = false;
  J.N_total = d.cout_pad; J.max_bn = d.cout_pad;
  J.d0 = make_dest(&y, nullptr);
  J.bias = bias; J.flags = d.flags;
  // the folded batch-norm epilogue exists in the halo-tile kernel only
  if (g_use_tconv && d.stride == 1 && d.kh == 3 && d.kw == 3 && !post_scale) {
    TconvJob T;
    memset(&T, 0, sizeof(T));
    T.a1 = J.a1; T.a2 = J.a2;
    T.pad_t = d.pad_t; T.pad_l = d.pad_l;
    T.d0 = y; T.d1 = null_view(); T.m0 = null_view(); T.m1 = null_view();
    T.w = J.w; T.w_rows = J.w_rows; T.w_cols = J.w_cols;
    T.b_mn = true; T.b_rows_per_tap = J.b_rows_per_tap; T.tap_flip = false;
    T.N_total = J.N_total; T.max_bn = J.max_bn;
    T.bias = bias; T.flags = d.flags;
    const int rc = launch_tconv(T, st);
    if (rc != SEG_E_UNSUPPORTED) return rc;
  }
  if (g_use_hconv && d.stride == 1 && d.kh * d.kw > 1) {
    HconvJob H;
    memset(&H, 0, sizeof(H));
    H.a1 = J.a1; H.a2 = J.a2;
    H.kh = d.kh; H.kw = d.kw;
    H.pad_t = d.pad_t; H.pad_l = d.pad_l;
    H.Hp = x.h + d.pad_t + d.pad_b; H.Wp_logical = x.w + d.pad_l + d.pad_r;
    H.Ho = y.h; H.Wo = y.w; H.batch = y.n;
    H.w = J.w; H.w_rows = J.w_rows; H.w_cols = J.w_cols;
    H.b_mn = true; H.b_rows_per_tap = J.b_rows_per_tap; H.tap_flip = false;
    H.N_total = J.N_total; H.max_bn = J.max_bn;
    H.d0 = J.d0; H.bias = bias; H.flags = d.flags;
    H.post_scale = post_scale; H.post_shift = post_shift;
    const int rc = launch_hconv(H, st);
    if (rc != SEG_E_UNSUPPORTED) return rc;
  }
  SEG_REQUIRE(!post_scale, SEG_E_UNSUPPORTED,
              "conv_fwd: the folded batch-norm epilogue needs the halo-tile kernel's geometry");
  return launch_igemm(J, st);
}

// cin_lo: first input channel of the slice [cin_lo, cin_lo + dx.c (+ dx2.c)) this call
// computes (0 and all of cin_pad for the whole gradient): the weight rows of tap t start at
// t * cin_pad + cin_lo, i.e. the same matrix seen from a row offset.
int umma_conv_dgrad(const seg_conv_desc& d, const seg_view& dz, const void* w, const seg_view& dx,
                    const seg_view* dx2, const seg_view* mask, const seg_view* mask2,
                    cudaStream_t st, int cin_lo) {
  SEG_REQUIRE(d.stride == 1, SEG_E_UNSUPPORTED, "umma conv_dgrad: stride 1 only");
  SEG_REQUIRE(dz.c == d.cout_pad, SEG_E_BAD_SHAPE, "conv_dgrad: dz.c %d != cout_pad %d", dz.c,
              d.cout_pad);
  const int n_slice = dx.c + ((dx2 && dx2->ptr) ? dx2->c : 0);
  SEG_REQUIRE(cin_lo >= 0 && cin_lo % 16 == 0 && cin_lo + n_slice <= d.cin_pad, SEG_E_BAD_SHAPE, "conv_dgrad: channel slice [%d, %d) outside cin_pad %d", cin_lo,
              cin_lo + n_slice, d.cin_pad);
  IgemmJob J;
  memset(&J, 0, sizeof(J));
  J.a1 = dz;
  J.a2 = null_view();
  J.kh = d.kh; J.kw = d.kw; J.stride = 1;
  J.low_h = -(d.kh - 1 - d.pad_t); J.low_w = -(d.kw - 1 - d.pad_l);
  J.up_h = dx.h - dz.h + J.low_h; J.up_w = dx.w - dz.w + J.low_w;
  J.Ho = dx.h; J.Wo = dx.w; J.batch = dx.n;
  J.w = reinterpret_cast<const uint8_t*>(w) + (size_t)cin_lo * d.cout_pad * 2;
  J.w_rows = d.kh * d.kw * d.cin_pad - cin_lo; J.w_cols = d.cout_pad;
  J.b_mn = false; J.b_rows_per_tap = d.cin_pad; J.tap_flip = true;
  J.N_total = n_slice;
  J.d0 = make_dest(&dx, mask);
  if (dx2 && dx2->ptr) {
    J.d1 = make_dest(dx2, mask2);
    J.split_n = dx.c;
    J.max_bn = dx.c;
    for (int c = 256; c >= 16; c >>= 1)
      if (dx.c % c == 0 && dx2->c % c == 0) { J.max_bn = c; break; }
  } else {
    J.max_bn = n_slice;
  }

  J.flags = d.flags & (SEG_EPI_RELU_MASK);
  if (g_use_tconv && d.kh == 3 && d.kw == 3) {
    TconvJob T;
    memset(&T, 0, sizeof(T));
    T.a1 = dz; T.a2 = null_view();
    T.pad_t = d.kh - 1 - d.pad_t; T.pad_l = d.kw - 1 - d.pad_l;
    T.d0 = dx;
    T.d1 = (dx2 && dx2->ptr) ? *dx2 : null_view();
    T.m0 = (mask && mask->ptr) ? *mask : null_view();
    T.m1 = (mask2 && mask2->ptr) ? *mask2 : null_view();
    T.w = J.w; T.w_rows = J.w_rows; T.w_cols = J.w_cols;
    T.b_mn = false; T.b_rows_per_tap = J.b_rows_per_tap; T.tap_flip = true;
    T.N_total = J.N_total; T.max_bn = J.max_bn; T.split_n = J.split_n;
    T.flags = J.flags;
    // a mask must cover every destination it is requested for
    const bool mask_ok = !(J.flags & SEG_EPI_RELU_MASK) ||
                         (T.m0.ptr && (!T.d1.ptr || T.m1.ptr));
    if (mask_ok) {
      const int rc = launch_tconv(T, st);
      if (rc != SEG_E_UNSUPPORTED) return rc;
    }
  }
  if (g_use_hconv && d.kh * d.kw > 1) {
    HconvJob H;
    memset(&H, 0, sizeof(H));
    H.a1 = dz; H.a2 = null_view();
    H.kh = d.kh; H.kw = d.kw;
    H.pad_t = d.kh - 1 - d.pad_t; H.pad_l = d.kw - 1 - d.pad_l;
    H.Hp = dx.h + d.kh - 1; H.Wp_logical = dx.w + d.kw - 1;
    H.Ho = dx.h; H.Wo = dx.w; H.batch = dx.n;
    H.w = J.w; H.w_rows = J.w_rows; H.w_cols = J.w_cols;
    H.b_mn = false; H.b_rows_per_tap = J.b_rows_per_tap; H.tap_flip = true;
    H.N_total = J.N_total; H.max_bn = J.max_bn;
    H.d0 = J.d0; H.d1 = J.d1; H.split_n = J.split_n; H.flags = J.flags;
    const int rc = launch_hconv(H, st);
    if (rc != SEG_E_UNSUPPORTED) return rc;
  }
  return launch_igemm(J, st);
}

int umma_conv_wgrad(const seg_conv_desc& d, const seg_view& x, const seg_view* x2,
                    const seg_view& dz, float* dw, float* db, cudaStream_t st) {
  if (x.c == 4 && d.cin == 3) {
    const int rc = fconv_wgrad(d, x, x2, dz, dw, db, st);
    if (rc != SEG_E_UNSUPPORTED) return rc;
    SEG_REQUIRE(false, SEG_E_UNSUPPORTED,
                "conv_wgrad: a 4-channel (R,G,B,1) input needs the first-layer kernel's shape");
  }
  WgradJob J;
  memset(&J, 0, sizeof(J));
  J.big = x;
  J.big2 = x2 ? *x2 : null_view();
  J.small_ = dz;
  J.kh = d.kh; J.kw = d.kw; J.stride = d.stride;
  J.low_h = -d.pad_t; J.low_w = -d.pad_l;
  J.up_h = d.pad_b - (d.kh - 1); J.up_w = d.pad_r - (d.kw - 1);
  J.BC = d.cin; J.SC = d.cout;
  J.dw = dw;
  J.db = db;
  if (g_use_twgrad) {
    const int rc = launch_twgrad(J, st);
    if (rc != SEG_E_UNSUPPORTED) return rc;
  }
  return launch_wgrad(J, st);
}

// Transposed conv with k > stride, VALID (DeconvModel's 5x5 / stride 2, reference
// models/deconvolution.py:150-160): output pixel o = stride*j + p receives only the taps
// k = p + stride*m, so each of the stride^2 output parity classes (py, px) is a stride-1
// correlation of x with an My x Mx sub-kernel (M_p = ceil((k - p) / stride)) over the fully
// padded input:  out_p[j] = sum_r x_pad[j + r] * w[p + stride*(M_p - 1 - r)].  Each class
// runs on the halo-tile kernel with a strided destination view and a tap -> weight-row
// table.  Returns SEG_E_UNSUPPORTED if a class does not fit that kernel.
static int deconv_fwd_parity(const seg_conv_desc& d, const seg_view& x, const void* w,
                             const float* bias, const seg_view& y, cudaStream_t st,
                             const float* post_scale, const float* post_shift) {
  const int s = d.stride, k = d.kh;
  const size_t esz = (d.flags & SEG_EPI_OUT_F32) ? 4 : 2;
  for (int py = 0; py < s; ++py) {
    const int My = (k - py + s - 1) / s;
    for (int px = 0; px < s; ++px) {
      const int Mx = (k - px + s - 1) / s;
      if (My < 1 || Mx < 1) continue;
      HconvJob H;
      memset(&H, 0, sizeof(H));
      H.a1 = x; H.a2 = null_view();
      H.kh = My; H.kw = Mx;
      H.pad_t = My - 1; H.pad_l = Mx - 1;
      H.Hp = x.h + 2 * (My - 1); H.Wp_logical = x.w + 2 * (Mx - 1);
      H.Ho = x.h + My - 1; H.Wo = x.w + Mx - 1; H.batch = x.n;
      SEG_REQUIRE(H.Ho == (y.h - py + s - 1) / s && H.Wo == (y.w - px + s - 1) / s,
                  SEG_E_BAD_SHAPE, "deconv_fwd: output %dx%d is not in*stride + k - stride", y.h,
                  y.w);
      seg_view yv = y;
      yv.ptr = reinterpret_cast<uint8_t*>(y.ptr) + ((int64_t)py * y.sh + (int64_t)px * y.sw) * esz;
      yv.h = H.Ho; yv.w = H.Wo;
      yv.sh = y.sh * s; yv.sw = y.sw * s;
      H.w = w; H.w_rows = k * k * d.cout_pad; H.w_cols = d.cin_pad;
      H.b_mn = false; H.b_rows_per_tap = d.cout_pad; H.tap_flip = false;
      H.N_total = d.cout_pad; H.max_bn = d.cout_pad;
      H.d0 = make_dest(&yv, nullptr);
      H.bias = bias; H.flags = d.flags;
      H.post_scale = post_scale; H.post_shift = post_shift;
      int rows[25];
      if (My * Mx > 25) return SEG_E_UNSUPPORTED;
      for (int r = 0; r < My; ++r)
        for (int c = 0; c < Mx; ++c)
          rows[r * Mx + c] = ((py + s * (My - 1 - r)) * k + (px + s * (Mx - 1 - c))) * d.cout_pad;
      H.tap_rows = rows;
      const int rc = launch_hconv(H, st);
      if (rc) return rc;
    }
  }
  return SEG_OK;
}

// transposed conv, VALID: k == stride as GEMM + pixel shuffle, k > stride by output parity
int umma_deconv_fwd(const seg_conv_desc& d, const seg_view& x, const void* w, const float* bias,
                    const seg_view& y, cudaStream_t st, const float* post_scale,
                    const float* post_shift) {
  SEG_REQUIRE(d.kh == d.kw && d.kh >= d.stride && d.pad_t == 0 && d.pad_l == 0 && d.pad_b == 0 &&
                  d.pad_r == 0,
              SEG_E_UNSUPPORTED, "umma deconv_fwd: square k >= stride, VALID only");
  SEG_REQUIRE(x.c == d.cin_pad, SEG_E_BAD_SHAPE, "deconv_fwd: x.c != cin_pad");
  if (d.kh > d.stride) {
    int rc = load_encoders();
    if (rc) return rc;
    SEG_REQUIRE(y.c <= d.cout_pad, SEG_E_BAD_SHAPE, "deconv_fwd: y.c > cout_pad");
    return deconv_fwd_parity(d, x, w, bias, y, st, post_scale, post_shift);
  }
  SEG_REQUIRE(!post_scale, SEG_E_UNSUPPORTED,
              "deconv_fwd: the folded batch-norm epilogue needs k > stride (halo-tile kernel)");
  IgemmJob J;
  memset(&J, 0, sizeof(J));
  J.a1 = x;
  J.a2 = null_view();
  J.kh = 1; J.kw = 1; J.stride = 1;
  J.Ho = x.h; J.Wo = x.w; J.batch = x.n;
  J.w = w; J.w_rows = d.kh * d.kw * d.cout_pad; J.w_cols = d.cin_pad;
  J.b_mn = false; J.b_rows_per_tap = 0; J.tap_flip = false;
  J.N_total = d.kh * d.kw * d.cout_pad; J.max_bn = d.cout_pad;
  J.d0 = make_dest(&y, nullptr);
  J.bias = bias; J.flags = d.flags;
  J.ps_k = d.stride; J.ps_cout = d.cout_pad;
  return launch_igemm(J, st);
}

int umma_deconv_dgrad(const seg_conv_desc& d, const seg_view& dz, const void* w,
                      const seg_view& dx, const seg_view* mask, cudaStream_t st) {
  // dx[j] = sum_k dz[stride*j + k] * w[k]: a strided correlation over dz, any k >= stride
  SEG_REQUIRE(d.kh == d.kw && d.kh >= d.stride && d.pad_t == 0 && d.pad_l == 0 && d.pad_b == 0 &&
                  d.pad_r == 0,
              SEG_E_UNSUPPORTED, "umma deconv_dgrad: square k >= stride, VALID only");
  SEG_REQUIRE(dz.c == d.cout_pad, SEG_E_BAD_SHAPE, "deconv_dgrad: dz.c != cout_pad");
  IgemmJob J;
  memset(&J, 0, sizeof(J));
  J.a1 = dz;
  J.a2 = null_view();
  J.kh = d.kh; J.kw = d.kw; J.stride = d.stride;
  J.low_h = 0; J.low_w = 0;
  J.up_h = -(d.kh - 1); J.up_w = -(d.kw - 1);
  J.Ho = dx.h; J.Wo = dx.w; J.batch = dx.n;
  J.w = w; J.w_rows = d.kh * d.kw * d.cout_pad; J.w_cols = d.cin_pad;
  J.b_mn = true; J.b_rows_per_tap = d.cout_pad; J.tap_flip = false;
  J.N_total = d.cin_pad; J.max_bn = d.cin_pad;
  J.d0 = make_dest(&dx, mask);
  J.flags = d.flags & (SEG_EPI_RELU_MASK);
  return launch_igemm(J, st);
}

int umma_deconv_wgrad(const seg_conv_desc& d, const seg_view& x, const seg_view& dz, float* dw,
                      cudaStream_t st) {
  SEG_REQUIRE(d.kh == d.kw && d.kh >= d.stride && d.pad_t == 0 && d.pad_l == 0 && d.pad_b == 0 &&
                  d.pad_r == 0,
              SEG_E_UNSUPPORTED, "umma deconv_wgrad: square k >= stride, VALID only");
  WgradJob J;
  memset(&J, 0, sizeof(J));
  J.big = dz;
  J.big2 = null_view();
  J.small_ = x;
  J.kh = d.kh; J.kw = d.kw; J.stride = d.stride;
  J.low_h = 0; J.low_w = 0;
  J.up_h = -(d.kh - 1); J.up_w = -(d.kw - 1);
  J.BC = d.cout; J.SC = d.cin;
  J.dw = dw;
  return launch_wgrad(J, st);
}

// probe: D[M][N] (fp32) = A[M][K] * B, A K-major bf16; mode bit0: B is [K][N] (MN-major)
// instead of [N][K] (K-major).  Exercises descriptors without im2col.
int umma_probe(int mode, int M, int N, int K, const void* a, const void* b, float* dptr,
               cudaStream_t st) {
  IgemmJob J;
  memset(&J, 0, sizeof(J));
  J.a_tiled2d = true;
  J.a1.ptr = const_cast<void*>(a);
  J.a1.n = 1; J.a1.h = 1; J.a1.w = M; J.a1.c = K;
  J.a1.sw = K; J.a1.sh = (int64_t)M * K; J.a1.sn = (int64_t)M * K;
  J.a2 = null_view();
  J.kh = 1; J.kw = 1; J.stride = 1;
  J.Ho = 1; J.Wo = M; J.batch = 1;
  J.w = b;
  J.b_mn = (mode & 1) != 0;
  if (J.b_mn) { J.w_rows = K; J.w_cols = N; } else { J.w_rows = N; J.w_cols = K; }
  J.b_rows_per_tap = 0;
  J.N_total = N; J.max_bn = N;
  seg_view dv;
  dv.ptr = dptr; dv.n = 1; dv.h = 1; dv.w = M; dv.c = N;
  dv.sw = N; dv.sh = (int64_t)M * N; dv.sn = (int64_t)M * N;
  J.d0 = make_dest(&dv, nullptr);
  J.flags = SEG_EPI_OUT_F32;
  return launch_igemm(J, st);
}

}  // namespace segb
