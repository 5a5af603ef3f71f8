// Fused distance + score head kernel (north-star kernel (a)).
//
// One thread owns VEC consecutive pixels of one image and holds their D-channel embedding in
// registers (NCHW input => D independent, fully coalesced VEC*4-byte loads per thread, all in
// flight together).  Everything the reference derives from the distance logits of a pixel is
// produced from those registers in the same pass; nothing of size N*K*D is ever materialised.
//
// Replaces anomaly/models/models.py:636-657, DeepLabV3Plus-Pytorch/network/utils.py:89-118,
// anomaly/eval_ood_traditional.py:218,276-278,288-290,301-304,434 and
// DeepLabV3Plus-Pytorch/test_embedding.py:339-350,428-433,445 (see include/dml_b200.h).
#pragma once
#include "dml_common.cuh"
#include "dml_bilinear.cuh"

namespace dml {

constexpr int HEAD_THREADS = 256;
// prototype modes: dense [K,D] table; m*I fast path; input already holds the logits z (scores only)
// HEAD_MS: the logits are gathered on the fly from up to HEAD_MAX_SCALES low-resolution logit maps
// (bilinear upsample, align_corners=False, averaged over the scales), then scored like HEAD_LOGITS
// HEAD_MSS: same, but every CTA owns a 2-D output tile and first stages the low-resolution footprint of
// all scales in shared memory (class-innermost, 16-byte vector reads, immediate class offsets)
constexpr int HEAD_DENSE = 0, HEAD_IDENT = 1, HEAD_LOGITS = 2, HEAD_MS = 3, HEAD_MSS = 4;
constexpr int MS_TILE_THREADS_X = 64;                      // threads per output row of a CTA tile
constexpr int MS_TILE_ROWS = 16;                           // output rows per CTA tile (4 per tile-loop iteration)
// shared-memory stride (floats) of one staged low-resolution pixel: classes padded to a multiple of 4 and the
// stride kept == 4 (mod 8) so that 8 neighbouring pixels fall into 8 disjoint 4-bank groups (conflict-free LDS.128)
__host__ __device__ constexpr int ms_stride(int D) { return (((D + 3) / 4) & 1) ? ((D + 3) & ~3) : ((D + 3) & ~3) + 4; }
constexpr int HEAD_MAX_NOVEL = 8;
constexpr int HEAD_MAX_SCALES = DML_MAX_SCALES;
constexpr int HEAD_MAX_CONF_BINS = 32 * 32;

struct HeadArgs {
  const float* x;
  const float* mu;
  float diag_m;
  float msp_scale;  // 2 * diag_m * log2(e): softmax over z_k == softmax over 2 m x_k when mu = m I
  int first;
  int ref_order;    // IDENT + logits only: evaluate d_k in the op order of the reference's torch-CPU code (bit-exact parity mode)
  float clamp;
  const double* mu_novel;
  int n_novel, novel_base;
  double novel_thr;
  float* logits;
  uint8_t* label_u8;
  long long* label_i64;
  float* maxlogit;
  float* eds;
  float* msp;
  float* feat;
  double* novel_dist;
  int* minmax;  // [B,4] fp32 bit patterns (all values are >= 0 so int order == float order)
  int want_eds_mm, want_msp_mm;
  const uint8_t* gt_u8;
  const long long* gt_i64;
  unsigned long long* conf;
  int crow, ccol;
  int B, K;
  long long HW;
  unsigned out_mask;  // OUT_* bits: which outputs are wanted (hoists the pointer tests out of the kernel)
  // HEAD_MS only: low-resolution maps [B,K,ms_h[s],ms_w[s]], fp32 source-index scales in/out
  // (torch area_pixel_compute_scale), divisor = number of scales, output width
  const float* ms_z[HEAD_MAX_SCALES];
  int ms_h[HEAD_MAX_SCALES], ms_w[HEAD_MAX_SCALES];
  float ms_rh[HEAD_MAX_SCALES], ms_rw[HEAD_MAX_SCALES];
  int ms_n, ms_W, ms_recip;
  float ms_div, ms_inv;
  // HEAD_MSS only: output height, per-scale staged footprint (rows, cols, float offset in shared memory)
  int ms_H, ms_smem_floats;
  int ms_fh[HEAD_MAX_SCALES], ms_fw[HEAD_MAX_SCALES], ms_soff[HEAD_MAX_SCALES];
};

enum : unsigned { OUT_LABEL_U8 = 1u, OUT_LABEL_I64 = 2u, OUT_MAXLOGIT = 4u, OUT_EDS = 8u, OUT_MSP = 16u, OUT_MINMAX = 32u,
                  OUT_CONF = 64u, OUT_GT_U8 = 128u, OUT_LOGITS = 256u, OUT_FEAT = 512u, OUT_NOVEL_DIST = 1024u };
constexpr int HEAD_TILES_PER_BLOCK = 4;  // consecutive tiles of one image per CTA: block-level work is amortised

// -(sum_d (x_d - mu_d)^2) in float64, in NumPy's pairwise-summation order for a contiguous
// row of D elements (DeepLabV3Plus-Pytorch/test_embedding.py:430 is np.sum(..., axis=1) on a
// C-contiguous [HW, D] float64 array): D < 8 sequential; otherwise 8 interleaved accumulators
// combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the D % 8 tail added sequentially.
// __dmul_rn/__dadd_rn keep nvcc from contracting to FMA (NumPy rounds the square separately).
template <int D, int VEC>
__device__ __forceinline__ double novel_neg_dist(const float (&x)[D][VEC], int v, const double* __restrict__ mu) {
  double res;
  if constexpr (D < 8) {
    res = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      double t = __dsub_rn((double)x[d][v], mu[d]);
      res = __dadd_rn(res, __dmul_rn(t, t));
    }
  } else {
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      double t = __dsub_rn((double)x[j][v], mu[j]);
      r[j] = __dmul_rn(t, t);
    }
    constexpr int BLK_END = D - (D % 8);
#pragma unroll
    for (int i = 8; i < BLK_END; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        double t = __dsub_rn((double)x[i + j][v], mu[i + j]);
        r[j] = __dadd_rn(r[j], __dmul_rn(t, t));
      }
    }
    res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                    __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
#pragma unroll
    for (int d = BLK_END; d < D; ++d) {
      double t = __dsub_rn((double)x[d][v], mu[d]);
      res = __dadd_rn(res, __dmul_rn(t, t));
    }
  }
  return -res;
}

// Multi-scale gather (anomaly/models/models.py:659-661 + anomaly/eval_ood_traditional.py:198-208):
//   scores[k] = sum_s  bilinear_s(z_s[k]) / n_scales        (accumulated in scale order, fp32)
// for VEC horizontally adjacent output pixels.  The arithmetic is torch's upsample_bilinear2d (source
// index scale*(dst+0.5)-0.5 clamped at 0, lambda1 = src - floor, value = h0*(w0*v00 + w1*v01) +
// h1*(w0*v10 + w1*v11)) with the FMA contraction pinned to the one the torch build evaluates --
// src = fma(scale, dst+0.5, -0.5), t = fma(w0, v_0, w1*v_1), value = fma(h0, t0, h1*t1) -- which makes
// the result bit-identical to torch's (checked against the CPU kernel in tests/).  The division by
// the number of scales is correctly rounded (q = v*inv, one FMA residual correction: Markstein) like the
// CPU `scores_tmp / 5`, or the plain multiplication by 1/n that torch's CUDA div-by-scalar kernel
// performs (`ms_recip`).  The low-resolution maps are tiny (<= 0.5 MB per image and scale) and are
// served by L1/L2; nothing of full resolution is read.
template <int D, int VEC>
__device__ __forceinline__ void ms_gather(const HeadArgs& a, int b, long long p0, float (&x)[D][VEC]) {
  const int y = (int)(p0 / a.ms_W);
  const int xo = (int)(p0 - (long long)y * a.ms_W);
#pragma unroll
  for (int d = 0; d < D; ++d)
#pragma unroll
    for (int v = 0; v < VEC; ++v) x[d][v] = 0.f;
#pragma unroll 1
  for (int s = 0; s < a.ms_n; ++s) {
    const int hs = a.ms_h[s], ws = a.ms_w[s];
    const float h1r = bilinear_src(a.ms_rh[s], y);
    const int h1 = (int)h1r;
    const int dy = (h1 < hs - 1) ? ws : 0;
    const float h1l = h1r - h1, h0l = 1.0f - h1l;
    int o00[VEC], dx[VEC];
    float w1l[VEC], w0l[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const float w1r = bilinear_src(a.ms_rw[s], xo + v);
      const int w1 = (int)w1r;
      dx[v] = (w1 < ws - 1) ? 1 : 0;
      w1l[v] = w1r - w1;
      w0l[v] = 1.0f - w1l[v];
      o00[v] = h1 * ws + w1;
    }
    const int plane = hs * ws;
    const float* q = a.ms_z[s] + (long long)b * D * plane;
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float* r = q + o00[v];
        const float v00 = __ldg(r), v01 = __ldg(r + dx[v]), v10 = __ldg(r + dy), v11 = __ldg(r + dy + dx[v]);
        const float val = bilinear_blend(w0l[v], w1l[v], h0l, h1l, v00, v01, v10, v11);
        x[k][v] = bl_add(x[k][v], scale_share(val, a.ms_inv, a.ms_div, a.ms_recip != 0));
      }
      q += plane;
    }
  }
}

// floor of torch's bilinear source index for output coordinate `dst` (same fp32 arithmetic as the taps)
__device__ __forceinline__ int ms_src_floor(float scale, int dst) { return (int)bilinear_src(scale, dst); }

// HEAD_MSS stage 1: copy the low-resolution footprint of the CTA's output tile (origin ty0, tx0) into shared
// memory for every scale, laid out [row][col][ms_stride(D)] (class innermost).  Rows / columns past the map
// edge are clamped duplicates that no tap addresses.
template <int D>
__device__ __forceinline__ void ms_stage(const HeadArgs& a, int b, int ty0, int tx0, float* s_ms) {
  constexpr int ST = ms_stride(D);
  for (int s = 0; s < a.ms_n; ++s) {
    const int hs = a.ms_h[s], ws = a.ms_w[s], fh = a.ms_fh[s], fw = a.ms_fw[s];
    const int r0 = ms_src_floor(a.ms_rh[s], ty0), c0 = ms_src_floor(a.ms_rw[s], tx0);
    const float* zb = a.ms_z[s] + (long long)b * D * hs * ws;
    float* S = s_ms + a.ms_soff[s];
    const int n = fh * fw;
    // one staged pixel per thread and step (a single integer division), its D classes in the inner loop:
    // consecutive threads read consecutive columns of one class plane
    for (int rc = threadIdx.x; rc < n; rc += HEAD_THREADS) {
      const int r = rc / fw, c = rc - r * fw;
      const float* src = zb + min(r0 + r, hs - 1) * ws + min(c0 + c, ws - 1);
      float* dst = S + rc * ST;
#pragma unroll
      for (int k = 0; k < D; ++k) dst[k] = __ldg(src + k * hs * ws);
    }
  }
}

// HEAD_MSS stage 2: the arithmetic of ms_gather with the four taps read from the staged footprint
// (ceil(D/4) 16-byte shared-memory loads per tap, class offsets are immediates).  Classes are processed
// two at a time with sm_100's packed fp32 instructions (FMUL2 / FFMA2 / FADD2: the same IEEE round-to-nearest
// result per lane, half the issue slots); the odd lane of the last pair is padding and is dropped.
template <int D, int VEC>
__device__ __forceinline__ void ms_gather_staged(const HeadArgs& a, const float* s_ms, int ty0, int tx0, int y, int xo,
                                                 float (&x)[D][VEC]) {
  constexpr int ST = ms_stride(D);
  constexpr int NQ = (D + 3) / 4;
  constexpr int NP = (D + 1) / 2;
  f32x2 acc[NP][VEC];
#pragma unroll
  for (int p = 0; p < NP; ++p)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[p][v] = 0ull;
  const f32x2 inv2 = pack2(a.ms_inv, a.ms_inv), ndiv2 = pack2(-a.ms_div, -a.ms_div);
  const bool recip = a.ms_recip != 0;
#pragma unroll 1
  for (int s = 0; s < a.ms_n; ++s) {
    const int hs = a.ms_h[s], ws = a.ms_w[s], fw = a.ms_fw[s];
    const int r0 = ms_src_floor(a.ms_rh[s], ty0), c0 = ms_src_floor(a.ms_rw[s], tx0);
    float h1r = __fmaf_rn(a.ms_rh[s], y + 0.5f, -0.5f);
    h1r = h1r < 0.f ? 0.f : h1r;
    const int h1 = (int)h1r;
    const int dy = (h1 < hs - 1) ? fw * ST : 0;
    const float h1l = h1r - h1, h0l = 1.0f - h1l;
    const f32x2 h0 = pack2(h0l, h0l), h1p = pack2(h1l, h1l);
    const float* S = s_ms + a.ms_soff[s] + (h1 - r0) * fw * ST;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float w1r = __fmaf_rn(a.ms_rw[s], (xo + v) + 0.5f, -0.5f);
      w1r = w1r < 0.f ? 0.f : w1r;
      const int w1 = (int)w1r;
      const int dx = (w1 < ws - 1) ? ST : 0;
      const float w1l = w1r - w1, w0l = 1.0f - w1l;
      const f32x2 w0 = pack2(w0l, w0l), w1p = pack2(w1l, w1l);
      const ulonglong2* p00 = reinterpret_cast<const ulonglong2*>(S + (w1 - c0) * ST);
      const ulonglong2* p01 = reinterpret_cast<const ulonglong2*>(S + (w1 - c0) * ST + dx);
      const ulonglong2* p10 = reinterpret_cast<const ulonglong2*>(S + (w1 - c0) * ST + dy);
      const ulonglong2* p11 = reinterpret_cast<const ulonglong2*>(S + (w1 - c0) * ST + dy + dx);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const ulonglong2 a00 = p00[q], a01 = p01[q], a10 = p10[q], a11 = p11[q];
        const f32x2 v00[2] = {a00.x, a00.y}, v01[2] = {a01.x, a01.y}, v10[2] = {a10.x, a10.y}, v11[2] = {a11.x, a11.y};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int p = 2 * q + j;
          if (p < NP) {
            const f32x2 t0 = fma2_rn(w0, v00[j], mul2_rn(w1p, v01[j]));
            const f32x2 t1 = fma2_rn(w0, v10[j], mul2_rn(w1p, v11[j]));
            const f32x2 val = fma2_rn(h0, t0, mul2_rn(h1p, t1));
            f32x2 t = mul2_rn(val, inv2);
            if (!recip) t = fma2_rn(fma2_rn(ndiv2, t, val), inv2, t);
            acc[p][v] = add2_rn(acc[p][v], t);
          }
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < NP; ++p)
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float lo, hi;
      unpack2(acc[p][v], lo, hi);
      x[2 * p][v] = lo;
      if (2 * p + 1 < D) x[2 * p + 1][v] = hi;
    }
}

// Adds one observation to a block-shared histogram with a single shared-memory update per distinct
// bin in the warp (segmentation labels are spatially coherent: usually 1-3 distinct bins per warp).
__device__ __forceinline__ void warp_histogram_add(unsigned int* s_bins, int bin) {
  unsigned todo = __ballot_sync(0xffffffffu, bin >= 0);
  const int lane = threadIdx.x & 31;
  while (todo) {
    const int leader = __ffs(todo) - 1;
    const int b = __shfl_sync(0xffffffffu, bin, leader);
    const unsigned members = __ballot_sync(0xffffffffu, bin == b);
    if (lane == leader) atomicAdd(&s_bins[b], (unsigned)__popc(members));
    todo &= ~members;
  }
}

// EXTRA = the outputs that keep x[][] live to the end or need fp64 / per-class stores (distance
// logits, NPM novel prototypes, novel_dist, NHWC feature copy); compiled separately so that the lean
// scoring path (labels + EDS + MSP + min/max + confusion) keeps a short instruction stream and a
// small register file.  `skip0` (OOD.exclude_back) only ever removes class 0 from the SCORES, so it
// costs one uniform select on the k = 0 terms instead of a predicate per class.
template <int D, int MODE, int VEC, bool EXTRA>
__global__ void __launch_bounds__(HEAD_THREADS) head_kernel(const HeadArgs a) {
  constexpr bool IDENT = (MODE == HEAD_IDENT);
  constexpr bool MSS = (MODE == HEAD_MSS);
  constexpr bool MS = (MODE == HEAD_MS) || MSS;
  constexpr bool LOGITS = (MODE == HEAD_LOGITS) || MS;
  constexpr bool DENSE = (MODE == HEAD_DENSE);
  constexpr float LOG2E = 1.4426950408889634f;
  constexpr float PINF = __builtin_huge_valf();
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  __shared__ int s_red[4][HEAD_THREADS / 32];

  // dynamic smem carve-up: [novel prototypes (double)] [dense mu, rows padded to DP floats] [confusion bins]
  constexpr int DP = (D + 3) & ~3;
  const int n_novel = EXTRA ? a.n_novel : 0;
  double* s_novel = reinterpret_cast<double*>(smem_dyn);
  float* s_mu = reinterpret_cast<float*>(s_novel + n_novel * D);
  unsigned int* s_conf = reinterpret_cast<unsigned int*>(s_mu + (DENSE ? a.K * DP : 0));
  const int K = DENSE ? a.K : D;
  const int nbins = a.conf ? a.crow * a.ccol : 0;

  if constexpr (EXTRA)
    for (int i = threadIdx.x; i < n_novel * D; i += HEAD_THREADS) s_novel[i] = a.mu_novel[i];
  if constexpr (DENSE) {
    for (int i = threadIdx.x; i < K * DP; i += HEAD_THREADS) {
      int k = i / DP, d = i - k * DP;
      s_mu[i] = d < D ? a.mu[k * D + d] : 0.f;
    }
  }
  for (int i = threadIdx.x; i < nbins; i += HEAD_THREADS) s_conf[i] = 0u;
  const int b = blockIdx.y;
  // HEAD_MSS: 2-D output tile of MS_TILE_ROWS x (MS_TILE_THREADS_X * VEC) pixels per CTA
  int ms_ty0 = 0, ms_tx0 = 0;
  float* s_ms = nullptr;
  if constexpr (MSS) {
    const int tiles_x = (a.ms_W + MS_TILE_THREADS_X * VEC - 1) / (MS_TILE_THREADS_X * VEC);
    const int tile_y = blockIdx.x / tiles_x;
    ms_ty0 = tile_y * MS_TILE_ROWS;
    ms_tx0 = (blockIdx.x - tile_y * tiles_x) * (MS_TILE_THREADS_X * VEC);
    s_ms = reinterpret_cast<float*>(smem_dyn + (((size_t)nbins * sizeof(unsigned) + 15) & ~(size_t)15));
    ms_stage<D>(a, b, ms_ty0, ms_tx0, s_ms);
  }
  if ((EXTRA && n_novel > 0) || DENSE || nbins > 0 || MSS) __syncthreads();

  const uint64_t pol = policy_evict_first();
  const bool skip0 = a.first != 0;
  const unsigned om = a.out_mask;
  const bool want_msp = (om & OUT_MSP) || a.want_msp_mm;
  const float* x_img = MS ? nullptr : a.x + ((long long)b * D) * a.HW;
  // running per-thread min / max of the score maps (int order == float order for values >= 0)
  int emin = 0x7fffffff, emax = (int)0x80000000, mmin = 0x7fffffff, mmax = (int)0x80000000;

#pragma unroll 1
  for (int it = 0; it < HEAD_TILES_PER_BLOCK; ++it) {
  long long p0;
  bool active;
  int ms_y = 0, ms_x = 0;
  if constexpr (MSS) {
    static_assert(HEAD_THREADS / MS_TILE_THREADS_X * HEAD_TILES_PER_BLOCK == MS_TILE_ROWS, "tile shape");
    ms_y = ms_ty0 + it * (HEAD_THREADS / MS_TILE_THREADS_X) + (threadIdx.x / MS_TILE_THREADS_X);
    ms_x = ms_tx0 + (threadIdx.x % MS_TILE_THREADS_X) * VEC;
    active = ms_y < a.ms_H && ms_x < a.ms_W;
    p0 = (long long)ms_y * a.ms_W + ms_x;
  } else {
    p0 = (((long long)blockIdx.x * HEAD_TILES_PER_BLOCK + it) * HEAD_THREADS + threadIdx.x) * VEC;
    active = p0 < a.HW;
  }

  float x[D][VEC];
  if constexpr (MS) {
    if (active) {
      if constexpr (MSS) ms_gather_staged<D, VEC>(a, s_ms, ms_ty0, ms_tx0, ms_y, ms_x, x);
      else ms_gather<D, VEC>(a, b, p0, x);
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d)
#pragma unroll
        for (int v = 0; v < VEC; ++v) x[d][v] = 0.f;
    }
    // optional store of the averaged full-resolution maps (the reference's `scores` / `ft1`)
    if ((om & OUT_LOGITS) && active) {
      float* lgm = a.logits + ((long long)b * D) * a.HW + p0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        FVec<VEC> t;
#pragma unroll
        for (int v = 0; v < VEC; ++v) t.v[v] = x[d][v];
        st_stream<VEC>(lgm + (long long)d * a.HW, t);
      }
    }
  } else {
    const float* q = x_img + p0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      FVec<VEC> t;
      if (active) {
        t = ld_stream<VEC>(q, pol);
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) t.v[v] = 0.f;
      }
      q += a.HW;
#pragma unroll
      for (int v = 0; v < VEC; ++v) x[d][v] = t.v[v];
    }
  }

  int label[VEC];
  float smin[VEC], eds[VEC], mspv[VEC];
  if constexpr (IDENT && !EXTRA) {
    // ---- lean scoring path for mu = m I (no per-class distance is ever formed) --------------------------------
    // With S = sum_k x_k^2 and t_k = (x_k - m)^2 the K = D distances are d_k = S - x_k^2 + t_k, hence
    //   sum_k d_k        = (D-1) S + sum_k t_k                       (EDS over all classes)
    //   sum_{k>=1} d_k   = (D-2) S + x_0^2 + sum_{k>=1} t_k          (EDS with class 0 excluded)
    // -- sums of non-negative terms only, so nothing cancels even when a pixel sits on its prototype -- and
    // d_k - d_j = -2 m (x_k - x_j) exactly: the nearest prototype is the largest (m >= 0) / smallest (m < 0)
    // channel, first index on ties like torch.max.  The max-softmax needs the same extremum (see below), so the
    // label costs one compare + select per class and the distances of the losing classes cost nothing.
    const bool sk = skip0 && D > 1;
    const bool up = a.msp_scale >= 0.f;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float s0 = x[0][v] * x[0][v], s1 = 0.f, t0 = 0.f, t1 = 0.f;   // two interleaved chains each (ILP, shorter error chains)
#pragma unroll
      for (int k = 1; k < D; ++k) {
        const float t = x[k][v] - a.diag_m;
        if (k & 1) { s1 = fmaf(x[k][v], x[k][v], s1); t1 = fmaf(t, t, t1); }
        else       { s0 = fmaf(x[k][v], x[k][v], s0); t0 = fmaf(t, t, t0); }
      }
      const float S = s0 + s1, T1 = t0 + t1;
      const float u0 = x[0][v] - a.diag_m;
      float e = sk ? fmaf((float)(D - 2), S, fmaf(x[0][v], x[0][v], T1)) : fmaf((float)(D - 1), S, fmaf(u0, u0, T1));
      if (a.clamp > 0.f) e = (e >= a.clamp) ? a.clamp : e;
      eds[v] = e;
      // extremum over the score classes (ext_s) and over all classes (ext_all)
      float ext1 = x[D > 1 ? 1 : 0][v];
#pragma unroll
      for (int k = 2; k < D; ++k) ext1 = up ? fmaxf(ext1, x[k][v]) : fminf(ext1, x[k][v]);
      const float ext_all = D > 1 ? (up ? fmaxf(ext1, x[0][v]) : fminf(ext1, x[0][v])) : x[0][v];
      const float ext_s = sk ? ext1 : ext_all;
      int lab = 0;
#pragma unroll
      for (int k = D - 1; k >= 1; --k) lab = (x[k][v] == ext_all) ? k : lab;
      label[v] = (x[0][v] == ext_all) ? 0 : lab;
      // softmax_k(z) == softmax_k(2 m x_k): max-softmax = 1 / sum_k exp2(c x_k - c x_ext), c = 2 m log2(e).
      // The sum lies in [1, D]: MUFU.RCP + one Newton step (<= 1 ulp) replaces the IEEE division sequence.
      float pr = 0.f;
      if (want_msp) {
        const float c = a.msp_scale;
        const float off = -c * ext_s;
        float s = sk ? 0.f : ex2_approx(fmaf(c, x[0][v], off));
#pragma unroll
        for (int k = 1; k < D; ++k) s += ex2_approx(fmaf(c, x[k][v], off));
        const float r = rcp_approx(s);
        pr = fmaf(r, fmaf(-s, r, 1.0f), r);
      }
      mspv[v] = pr;
      smin[v] = 0.f;
    }
    if (om & OUT_MAXLOGIT) {
      // max logit = -(distance to the nearest score class ls): leave-one-out sum of squares + (x_ls - m)^2,
      // where x_ls is the extremum itself
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float ext1 = x[D > 1 ? 1 : 0][v];
#pragma unroll
        for (int k = 2; k < D; ++k) ext1 = up ? fmaxf(ext1, x[k][v]) : fminf(ext1, x[k][v]);
        int ls = label[v];
        float xe = x[0][v];
        if (sk) {
          ls = 1;
#pragma unroll
          for (int k = D - 1; k >= 1; --k) ls = (x[k][v] == ext1) ? k : ls;
          xe = ext1;
        } else if (D > 1) {
          xe = up ? fmaxf(ext1, x[0][v]) : fminf(ext1, x[0][v]);
        }
        float r = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) r = (d == ls) ? r : fmaf(x[d][v], x[d][v], r);
        const float t = xe - a.diag_m;
        smin[v] = fmaf(t, t, r);
      }
    }
  } else {
  // per pixel: d0 = distance to class 0; (dmin1, arg1) = best of classes >= 1; esum1 = sum over classes >= 1
  float d0[VEC], dmin1[VEC], esum1[VEC], ssum[VEC];
  int arg1[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { d0[v] = 0.f; dmin1[v] = PINF; esum1[v] = 0.f; ssum[v] = 0.f; arg1[v] = 0; }
  float* lg = nullptr;
  if constexpr (EXTRA && !LOGITS) lg = (om & OUT_LOGITS) ? a.logits + ((long long)b * K) * a.HW + p0 : nullptr;

  // esum1 accumulates the score classes in index order starting from the first one -- ((d_0 + d_1) + d_2) + ... --
  // the order torch.sum(scores, dim=1) uses on the class planes (anomaly/eval_ood_traditional.py:302), so that the
  // EDS map is bit-identical to the reference's on identical logits
  auto visit = [&](int k, int v, float dk) {
    if (k == 0) {
      d0[v] = dk;
      esum1[v] = skip0 ? 0.f : dk;
    } else {
      if (dk < dmin1[v]) { dmin1[v] = dk; arg1[v] = k; }
      esum1[v] += dk;
    }
  };

  if constexpr (IDENT) {
    // d_k = sum_{d != k} x_d^2 + (x_k - m)^2.  The leave-one-out sum of squares is assembled
    // from positive terms only (groups of 4 channels: in-group partners + all other groups),
    // so there is no cancellation against ||x||^2 when a pixel sits on its prototype.
    constexpr int NG = (D + 3) / 4;
    float outer[NG][VEC];
    {
      float g[NG][VEC];
#pragma unroll
      for (int j = 0; j < NG; ++j)
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          float s = x[4 * j][v] * x[4 * j][v];
#pragma unroll
          for (int m = 1; m < 4; ++m)
            if (4 * j + m < D) s = fmaf(x[4 * j + m][v], x[4 * j + m][v], s);
          g[j][v] = s;
        }
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float pre[NG], suf[NG];
        pre[0] = 0.f;
#pragma unroll
        for (int j = 1; j < NG; ++j) pre[j] = pre[j - 1] + g[j - 1][v];
        suf[NG - 1] = 0.f;
#pragma unroll
        for (int j = NG - 2; j >= 0; --j) suf[j] = suf[j + 1] + g[j + 1][v];
#pragma unroll
        for (int j = 0; j < NG; ++j) outer[j][v] = pre[j] + suf[j];
      }
    }
    bool ref_done = false;
    if constexpr (EXTRA && D < 16) {
      if (a.ref_order) {
        // Parity mode: z_k = -sum_d (x_d - mu_kd)^2 exactly as the reference's torch-CPU ops round it
        // (anomaly/models/models.py:649-651): subtract (x - 0 is exact), square, then torch's CPU sum over a
        // contiguous inner dim of D < 16 floats, one rounded add at a time: for 8 <= D < 16 the tail elements 8 .. D-1
        // first, then 0 .. 7; for D < 8 element 0, then the tail 4 .. D-1, then 1 .. 3 (recovered by probing the build
        // in this image and pinned by tests/test_head_reference_order.py); no FMA contraction.
        ref_done = true;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          FVec<VEC> z;
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < D; ++i) {
              int d;
              if constexpr (D >= 8) d = i < D - 8 ? 8 + i : i - (D - 8);
              else if constexpr (D > 4) d = i == 0 ? 0 : (i <= D - 4 ? 3 + i : i - (D - 4));
              else d = i;
              const float t = (d == k) ? __fsub_rn(x[d][v], a.diag_m) : x[d][v];
              const float sq = __fmul_rn(t, t);
              acc = (i == 0) ? sq : __fadd_rn(acc, sq);
            }
            visit(k, v, acc);
            z.v[v] = -acc;
          }
          if (lg && active) st_stream<VEC>(lg + (long long)k * a.HW, z);
        }
      }
    }
    if (!ref_done) {
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const int j = k >> 2;
      FVec<VEC> z;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float r = outer[j][v];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int kk = 4 * j + m;
          if (kk < D && kk != k) r = fmaf(x[kk][v], x[kk][v], r);
        }
        const float t = x[k][v] - a.diag_m;
        const float dk = fmaf(t, t, r);
        visit(k, v, dk);
        z.v[v] = -dk;
      }
      if constexpr (EXTRA)
        if (lg && active) st_stream<VEC>(lg + (long long)k * a.HW, z);
    }
    }  // !ref_done
    // softmax_k(z) == softmax_k(2 m x_k) when mu = m I (z_k - z_j = 2 m (x_k - x_j) exactly), so the
    // max-softmax is 1 / sum_k exp2(c x_k - c x_ext), c = 2 m log2(e), x_ext the max (m > 0) / min (m < 0).
    if (want_msp) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float c = a.msp_scale;
        float ext = skip0 ? x[D > 1 ? 1 : 0][v] : x[0][v];
#pragma unroll
        for (int k = 1; k < D; ++k) ext = c >= 0.f ? fmaxf(ext, x[k][v]) : fminf(ext, x[k][v]);
        const float off = -c * ext;
        float s = skip0 ? 0.f : ex2_approx(fmaf(c, x[0][v], off));
#pragma unroll
        for (int k = 1; k < D; ++k) s += ex2_approx(fmaf(c, x[k][v], off));
        ssum[v] = s;
      }
    }
  } else if constexpr (LOGITS) {
    // input channels ARE the logits z_k (anomaly path: distances taken at stride 8, then upsampled
    // and averaged over scales by the caller, anomaly/eval_ood_traditional.py:198-210): d_k = -z_k
#pragma unroll
    for (int k = 0; k < D; ++k)
#pragma unroll
      for (int v = 0; v < VEC; ++v) visit(k, v, -x[k][v]);
    if (want_msp) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float zmax = skip0 ? x[D > 1 ? 1 : 0][v] : x[0][v];
#pragma unroll
        for (int k = 1; k < D; ++k) zmax = fmaxf(zmax, x[k][v]);
        float s = skip0 ? 0.f : ex2_approx((x[0][v] - zmax) * LOG2E);
#pragma unroll
        for (int k = 1; k < D; ++k) s += ex2_approx((x[k][v] - zmax) * LOG2E);
        ssum[v] = s;
      }
    }
  } else {
    // dense prototypes: direct form, prototypes broadcast from shared memory; online softmax over the
    // score classes (running minimum distance mrun, running sum of exp(-(d_k - mrun)))
    float mrun[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) mrun[v] = PINF;
    for (int k = 0; k < K; ++k) {
      const float4* mrow = reinterpret_cast<const float4*>(s_mu + k * DP);
      float acc[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
#pragma unroll
      for (int q = 0; q < DP / 4; ++q) {
        const float4 m4 = mrow[q];
        const float mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int d = 4 * q + m;
          if (d < D) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              const float t = x[d][v] - mm[m];
              acc[v] = fmaf(t, t, acc[v]);
            }
          }
        }
      }
      FVec<VEC> z;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float dk = acc[v];
        visit(k, v, dk);
        if (want_msp && (k > 0 || !skip0)) {
          const float nm = fminf(mrun[v], dk);
          ssum[v] = ssum[v] * ex2_approx((nm - mrun[v]) * LOG2E) + ex2_approx((nm - dk) * LOG2E);
          mrun[v] = nm;
        }
        z.v[v] = -dk;
      }
      if constexpr (EXTRA)
        if (lg && active) st_stream<VEC>(lg + (long long)k * a.HW, z);
    }
  }

  // ---- per-pixel results ---------------------------------------------------------------
  // label = argmin over ALL classes (first minimum wins, like torch.max on the logits);
  // scores (eds, maxlogit, msp) over the score classes (all, or all but class 0)
  float dbest[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const bool zero_wins = (K == 1) || !(dmin1[v] < d0[v]);
    label[v] = zero_wins ? 0 : arg1[v];
    dbest[v] = zero_wins ? d0[v] : dmin1[v];
    smin[v] = (skip0 && K > 1) ? dmin1[v] : dbest[v];
    float e = (skip0 && K == 1) ? d0[v] : esum1[v];
    if (a.clamp > 0.f) e = (e >= a.clamp) ? a.clamp : e;
    eds[v] = e;
    mspv[v] = want_msp ? (1.0f / ssum[v]) : 0.f;
  }

  // NPM override (float64 like the reference): label <- novel_base + j where
  // z_nov > thr && z_nov > max_k z_k  (max over ALL classes: test_embedding.py:445)
  if constexpr (EXTRA) {
    for (int j = 0; j < n_novel; ++j) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const double zn = novel_neg_dist<D, VEC>(x, v, s_novel + j * D);
        const double zmax = (double)(-dbest[v]);
        if (zn > a.novel_thr && zn > zmax) label[v] = a.novel_base + j;
        if ((om & OUT_NOVEL_DIST) && active)
          a.novel_dist[((long long)j * a.B + b) * a.HW + p0 + v] = zn;
      }
    }
  }

  }  // generic (per-class distance) path

  const long long pix = (long long)b * a.HW + p0;
  if (active) {
    if (om & OUT_LABEL_U8) {
      if constexpr (VEC == 4) {
        *reinterpret_cast<uchar4*>(a.label_u8 + pix) =
            make_uchar4((unsigned char)label[0], (unsigned char)label[1], (unsigned char)label[2], (unsigned char)label[3]);
      } else if constexpr (VEC == 2) {
        *reinterpret_cast<uchar2*>(a.label_u8 + pix) = make_uchar2((unsigned char)label[0], (unsigned char)label[1]);
      } else {
        a.label_u8[pix] = (unsigned char)label[0];
      }
    }
    if (om & OUT_LABEL_I64) {
#pragma unroll
      for (int v = 0; v < VEC; v += 2) {
        if constexpr (VEC >= 2) {
          *reinterpret_cast<longlong2*>(a.label_i64 + pix + v) = make_longlong2(label[v], label[v + 1]);
        } else {
          a.label_i64[pix] = label[0];
        }
      }
    }
    if (om & OUT_MAXLOGIT) {
      FVec<VEC> t;
#pragma unroll
      for (int v = 0; v < VEC; ++v) t.v[v] = -smin[v];
      st_keep<VEC>(a.maxlogit + pix, t);
    }
    if (om & OUT_EDS) {
      FVec<VEC> t;
#pragma unroll
      for (int v = 0; v < VEC; ++v) t.v[v] = eds[v];
      st_keep<VEC>(a.eds + pix, t);
    }
    if (om & OUT_MSP) {
      FVec<VEC> t;
#pragma unroll
      for (int v = 0; v < VEC; ++v) t.v[v] = mspv[v];
      st_keep<VEC>(a.msp + pix, t);
    }
    if constexpr (EXTRA) if (om & OUT_FEAT) {
      // NHWC copy: the D channels of a pixel are contiguous -- 16-byte stores when the row length allows it
      // (cudaMalloc'ed / torch tensors are 256-byte aligned; D % 4 == 0 keeps every pixel 16-byte aligned)
      float* f = a.feat + pix * D;
      if constexpr (D % 4 == 0) {
        if ((reinterpret_cast<uintptr_t>(a.feat) & 15) == 0) {
#pragma unroll
          for (int v = 0; v < VEC; ++v)
#pragma unroll
            for (int d = 0; d < D; d += 4)
              *reinterpret_cast<float4*>(f + v * D + d) = make_float4(x[d][v], x[d + 1][v], x[d + 2][v], x[d + 3][v]);
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v)
#pragma unroll
            for (int d = 0; d < D; ++d) f[v * D + d] = x[d][v];
        }
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v)
#pragma unroll
          for (int d = 0; d < D; ++d) f[v * D + d] = x[d][v];
      }
    }
  }

  // ---- running min / max of the raw score maps (reduced per block after the tile loop) ----------
  if ((om & OUT_MINMAX) && active) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int e = __float_as_int(eds[v]), m = __float_as_int(mspv[v]);
      emin = min(emin, e); emax = max(emax, e);
      mmin = min(mmin, m); mmax = max(mmax, m);
    }
  }

  // ---- fused confusion counts (block-shared histogram, flushed once after the tile loop) -----------
  if (om & OUT_CONF) {
    int bin[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) bin[v] = -1;
    if (active) {
      if (om & OUT_GT_U8) {
        unsigned char g[VEC];
        if constexpr (VEC == 4) {
          const uchar4 t = *reinterpret_cast<const uchar4*>(a.gt_u8 + pix);
          g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
        } else if constexpr (VEC == 2) {
          const uchar2 t = *reinterpret_cast<const uchar2*>(a.gt_u8 + pix);
          g[0] = t.x; g[1] = t.y;
        } else {
          g[0] = a.gt_u8[pix];
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v)
          if ((int)g[v] < a.crow && label[v] < a.ccol) bin[v] = (int)g[v] * a.ccol + label[v];
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const long long g = a.gt_i64[pix + v];
          if (g >= 0 && g < a.crow && label[v] < a.ccol) bin[v] = (int)g * a.ccol + label[v];
        }
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) warp_histogram_add(s_conf, bin[v]);
  }
  }  // tile loop

  if (om & OUT_MINMAX) {
    emin = __reduce_min_sync(0xffffffffu, emin); emax = __reduce_max_sync(0xffffffffu, emax);
    mmin = __reduce_min_sync(0xffffffffu, mmin); mmax = __reduce_max_sync(0xffffffffu, mmax);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { s_red[0][w] = emin; s_red[1][w] = emax; s_red[2][w] = mmin; s_red[3][w] = mmax; }
    __syncthreads();
    if (threadIdx.x < 4) {
      int r = s_red[threadIdx.x][0];
      const bool is_min = (threadIdx.x & 1) == 0;
#pragma unroll
      for (int i = 1; i < HEAD_THREADS / 32; ++i) r = is_min ? min(r, s_red[threadIdx.x][i]) : max(r, s_red[threadIdx.x][i]);
      const bool wanted = threadIdx.x < 2 ? a.want_eds_mm : a.want_msp_mm;
      if (wanted) {
        int* dst = a.minmax + b * 4 + threadIdx.x;
        if (is_min) atomicMin(dst, r); else atomicMax(dst, r);
      }
    }
  }
  if (om & OUT_CONF) {
    __syncthreads();
    for (int i = threadIdx.x; i < nbins; i += HEAD_THREADS) {
      const unsigned c = s_conf[i];
      if (c) atomicAdd(a.conf + i, (unsigned long long)c);
    }
  }
}

template <int D, int MODE, int VEC, bool EXTRA>
int launch_head(const HeadArgs& a, cudaStream_t stream) {
  constexpr int DP = (D + 3) & ~3;
  const long long per_block = (long long)HEAD_THREADS * VEC * HEAD_TILES_PER_BLOCK;
  dim3 grid((unsigned)((a.HW + per_block - 1) / per_block), (unsigned)a.B);
  size_t smem = (EXTRA ? (size_t)a.n_novel * D * sizeof(double) : 0) + (MODE == HEAD_DENSE ? (size_t)a.K * DP * sizeof(float) : 0) +
                (a.conf ? (size_t)a.crow * a.ccol * sizeof(unsigned) : 0);
  if constexpr (MODE == HEAD_MSS) {
    const int tiles_x = (a.ms_W + MS_TILE_THREADS_X * VEC - 1) / (MS_TILE_THREADS_X * VEC);
    const int tiles_y = (a.ms_H + MS_TILE_ROWS - 1) / MS_TILE_ROWS;
    grid.x = (unsigned)(tiles_x * tiles_y);
    smem = ((smem + 15) & ~(size_t)15) + (size_t)a.ms_smem_floats * sizeof(float);
  }
  head_kernel<D, MODE, VEC, EXTRA><<<grid, HEAD_THREADS, smem, stream>>>(a);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

// per-translation-unit dispatch over a range of D (keeps each TU's compile time bounded)
#define DML_HEAD_CASE_I(Dv, MD)                                                                \
  if (extra) return vec >= 2 ? launch_head<Dv, MD, 2, true>(a, s) : launch_head<Dv, MD, 1, true>(a, s); \
  return vec == 4 ? launch_head<Dv, MD, 4, false>(a, s) : vec == 2 ? launch_head<Dv, MD, 2, false>(a, s) : launch_head<Dv, MD, 1, false>(a, s);
#define DML_HEAD_CASE(Dv)                                                                      \
  case Dv:                                                                                     \
    if (mode == HEAD_IDENT) { DML_HEAD_CASE_I(Dv, HEAD_IDENT) }                                \
    if (mode == HEAD_MS)                                                                       \
      return vec >= 2 ? launch_head<Dv, HEAD_MS, 2, false>(a, s) : launch_head<Dv, HEAD_MS, 1, false>(a, s); \
    if (mode == HEAD_MSS)                                                                      \
      return vec >= 2 ? launch_head<Dv, HEAD_MSS, 2, false>(a, s) : launch_head<Dv, HEAD_MSS, 1, false>(a, s); \
    if (mode == HEAD_LOGITS)                                                                   \
      return vec == 4 ? launch_head<Dv, HEAD_LOGITS, 4, false>(a, s) : vec == 2 ? launch_head<Dv, HEAD_LOGITS, 2, false>(a, s) : launch_head<Dv, HEAD_LOGITS, 1, false>(a, s); \
    DML_HEAD_CASE_I(Dv, HEAD_DENSE)

int head_dispatch_1_8(int D, int mode, int vec, bool extra, const HeadArgs& a, cudaStream_t s);
int head_dispatch_9_16(int D, int mode, int vec, bool extra, const HeadArgs& a, cudaStream_t s);
int head_dispatch_17_24(int D, int mode, int vec, bool extra, const HeadArgs& a, cudaStream_t s);
int head_dispatch_25_32(int D, int mode, int vec, bool extra, const HeadArgs& a, cudaStream_t s);

}  // namespace dml
