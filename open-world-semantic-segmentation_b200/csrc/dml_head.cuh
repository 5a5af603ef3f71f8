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

namespace dml {

constexpr int HEAD_THREADS = 256;
// prototype modes: dense [K,D] table; m*I fast path; input already holds the logits z (scores only)
constexpr int HEAD_DENSE = 0, HEAD_IDENT = 1, HEAD_LOGITS = 2;
constexpr int HEAD_MAX_NOVEL = 8;
constexpr int HEAD_MAX_CONF_BINS = 32 * 32;

struct HeadArgs {
  const float* x;
  const float* mu;
  float diag_m;
  float msp_scale;  // 2 * diag_m * log2(e): softmax over z_k == softmax over 2 m x_k when mu = m I
  int first;
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
};

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

// EXTRA = the rarely used outputs that keep x[][] live to the end and need fp64 (NPM novel prototypes,
// novel_dist, NHWC feature copy); compiled separately so the lean scoring path keeps a small register file.
template <int D, int MODE, int VEC, bool EXTRA>
__global__ void __launch_bounds__(HEAD_THREADS) head_kernel(const HeadArgs a) {
  constexpr bool IDENT = (MODE == HEAD_IDENT);
  constexpr bool LOGITS = (MODE == HEAD_LOGITS);
  constexpr bool DENSE = (MODE == HEAD_DENSE);
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

  for (int i = threadIdx.x; i < n_novel * D; i += HEAD_THREADS) s_novel[i] = a.mu_novel[i];
  if constexpr (DENSE) {
    for (int i = threadIdx.x; i < K * DP; i += HEAD_THREADS) {
      int k = i / DP, d = i - k * DP;
      s_mu[i] = d < D ? a.mu[k * D + d] : 0.f;
    }
  }
  for (int i = threadIdx.x; i < nbins; i += HEAD_THREADS) s_conf[i] = 0u;
  const bool need_sync = (n_novel > 0) || DENSE || nbins > 0;
  if (need_sync) __syncthreads();

  const int b = blockIdx.y;
  const long long p0 = ((long long)blockIdx.x * HEAD_THREADS + threadIdx.x) * VEC;
  const bool active = p0 < a.HW;
  const uint64_t pol = policy_evict_first();

  float x[D][VEC];
  {
    const float* xb = a.x + ((long long)b * D) * a.HW + p0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      FVec<VEC> t;
      if (active) {
        t = ld_stream<VEC>(xb + (long long)d * a.HW, pol);
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) t.v[v] = 0.f;
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) x[d][v] = t.v[v];
    }
  }

  float dmin[VEC], smin[VEC], eds[VEC], ssum[VEC];
  int arg[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    dmin[v] = __int_as_float(0x7f800000);
    smin[v] = __int_as_float(0x7f800000);
    eds[v] = 0.f;
    ssum[v] = 0.f;
    arg[v] = 0;
  }
  const bool want_msp = (a.msp != nullptr) || a.want_msp_mm;
  float* lg = (a.logits && !LOGITS) ? a.logits + ((long long)b * K) * a.HW + p0 : nullptr;

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
    // softmax_k(z) == softmax_k(2 m x_k) when mu = m I (z_k - z_j = 2 m (x_k - x_j) exactly), so the
    // max-softmax is 1 / sum_k exp(2m (x_k - x_ext)) with x_ext the max (m > 0) or min (m < 0) of x.
    float xext[VEC];
    if (want_msp) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float t = a.msp_scale >= 0.f ? __int_as_float(0xff800000) : __int_as_float(0x7f800000);
#pragma unroll
        for (int k = 0; k < D; ++k)
          if (k >= a.first) t = a.msp_scale >= 0.f ? fmaxf(t, x[k][v]) : fminf(t, x[k][v]);
        xext[v] = t;
      }
    }
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
        if (dk < dmin[v]) { dmin[v] = dk; arg[v] = k; }
        if (k >= a.first) {
          smin[v] = fminf(smin[v], dk);
          eds[v] += dk;
          if (want_msp) ssum[v] += ex2_approx(a.msp_scale * (x[k][v] - xext[v]));
        }
        z.v[v] = -dk;
      }
      if (lg && active) st_stream<VEC>(lg + (long long)k * a.HW, z);
    }
  } else if constexpr (LOGITS) {
    // input channels ARE the logits z_k (anomaly path: distances taken at stride 8, then upsampled
    // and averaged over scales by the caller, anomaly/eval_ood_traditional.py:198-210): d_k = -z_k
    constexpr float LOG2E = 1.4426950408889634f;
    float zmax[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float t = __int_as_float(0xff800000);
#pragma unroll
      for (int k = 0; k < D; ++k)
        if (k >= a.first) t = fmaxf(t, x[k][v]);
      zmax[v] = t;
    }
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float dk = -x[k][v];
        if (dk < dmin[v]) { dmin[v] = dk; arg[v] = k; }
        if (k >= a.first) {
          eds[v] += dk;
          if (want_msp) ssum[v] += ex2_approx((x[k][v] - zmax[v]) * LOG2E);
        }
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) smin[v] = -zmax[v];
  } else {
    // dense prototypes: direct form, prototypes broadcast from shared memory; online softmax
    constexpr float LOG2E = 1.4426950408889634f;
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
        if (dk < dmin[v]) { dmin[v] = dk; arg[v] = k; }
        if (k >= a.first) {
          if (want_msp) {
            const float nm = fminf(smin[v], dk);
            // sum_k exp(-(d_k - dmin)) maintained online
            ssum[v] = ssum[v] * ex2_approx((nm - smin[v]) * LOG2E) + ex2_approx((nm - dk) * LOG2E);
            smin[v] = nm;
          } else {
            smin[v] = fminf(smin[v], dk);
          }
          eds[v] += dk;
        }
        z.v[v] = -dk;
      }
      if (lg && active) st_stream<VEC>(lg + (long long)k * a.HW, z);
    }
  }

  // ---- per-pixel results ---------------------------------------------------------------
  int label[VEC];
  float mspv[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    label[v] = arg[v];
    if (a.clamp > 0.f) eds[v] = (eds[v] >= a.clamp) ? a.clamp : eds[v];
    mspv[v] = want_msp ? (1.0f / ssum[v]) : 0.f;
  }

  // NPM override (float64 like the reference): label <- novel_base + j where
  // z_nov > thr && z_nov > max_k z_k  (max over ALL classes: test_embedding.py:445)
  if constexpr (EXTRA) {
    for (int j = 0; j < n_novel; ++j) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const double zn = novel_neg_dist<D, VEC>(x, v, s_novel + j * D);
        const double zmax = (double)(-dmin[v]);
        if (zn > a.novel_thr && zn > zmax) label[v] = a.novel_base + j;
        if (a.novel_dist && active)
          a.novel_dist[((long long)j * a.B + b) * a.HW + p0 + v] = zn;
      }
    }
  }

  const long long pix = (long long)b * a.HW + p0;
  if (active) {
    if (a.label_u8) {
      if constexpr (VEC == 4) {
        *reinterpret_cast<uchar4*>(a.label_u8 + pix) =
            make_uchar4((unsigned char)label[0], (unsigned char)label[1], (unsigned char)label[2], (unsigned char)label[3]);
      } else if constexpr (VEC == 2) {
        *reinterpret_cast<uchar2*>(a.label_u8 + pix) = make_uchar2((unsigned char)label[0], (unsigned char)label[1]);
      } else {
        a.label_u8[pix] = (unsigned char)label[0];
      }
    }
    if (a.label_i64) {
#pragma unroll
      for (int v = 0; v < VEC; v += 2) {
        if constexpr (VEC >= 2) {
          *reinterpret_cast<longlong2*>(a.label_i64 + pix + v) = make_longlong2(label[v], label[v + 1]);
        } else {
          a.label_i64[pix] = label[0];
        }
      }
    }
    if (a.maxlogit) {
      FVec<VEC> t;
#pragma unroll
      for (int v = 0; v < VEC; ++v) t.v[v] = -smin[v];
      st_keep<VEC>(a.maxlogit + pix, t);
    }
    if (a.eds) {
      FVec<VEC> t;
#pragma unroll
      for (int v = 0; v < VEC; ++v) t.v[v] = eds[v];
      st_keep<VEC>(a.eds + pix, t);
    }
    if (a.msp) {
      FVec<VEC> t;
#pragma unroll
      for (int v = 0; v < VEC; ++v) t.v[v] = mspv[v];
      st_keep<VEC>(a.msp + pix, t);
    }
    if constexpr (EXTRA) if (a.feat) {
      float* f = a.feat + pix * D;
#pragma unroll
      for (int v = 0; v < VEC; ++v)
#pragma unroll
        for (int d = 0; d < D; ++d) f[v * D + d] = x[d][v];
    }
  }

  // ---- per-image min / max of the raw score maps (for the min-max normalisation) ---------
  if (a.minmax && (a.want_eds_mm || a.want_msp_mm)) {
    int emin = 0x7fffffff, emax = (int)0x80000000, mmin = 0x7fffffff, mmax = (int)0x80000000;
    if (active) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int e = __float_as_int(eds[v]), m = __float_as_int(mspv[v]);
        emin = min(emin, e); emax = max(emax, e);
        mmin = min(mmin, m); mmax = max(mmax, m);
      }
    }
    emin = warp_reduce_min_i(emin); emax = warp_reduce_max_i(emax);
    mmin = warp_reduce_min_i(mmin); mmax = warp_reduce_max_i(mmax);
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

  // ---- fused confusion counts -------------------------------------------------------------
  if (nbins > 0) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      int bin = -1;
      if (active) {
        const long long g = a.gt_u8 ? (long long)a.gt_u8[pix + v] : a.gt_i64[pix + v];
        if (g >= 0 && g < a.crow && label[v] < a.ccol) bin = (int)g * a.ccol + label[v];
      }
      const unsigned peers = __match_any_sync(0xffffffffu, bin);
      if (bin >= 0 && (threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&s_conf[bin], (unsigned)__popc(peers));
    }
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
  const long long per_block = (long long)HEAD_THREADS * VEC;
  dim3 grid((unsigned)((a.HW + per_block - 1) / per_block), (unsigned)a.B);
  size_t smem = (EXTRA ? (size_t)a.n_novel * D * sizeof(double) : 0) + (MODE == HEAD_DENSE ? (size_t)a.K * DP * sizeof(float) : 0) +
                (a.conf ? (size_t)a.crow * a.ccol * sizeof(unsigned) : 0);
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
    if (mode == HEAD_LOGITS)                                                                   \
      return vec == 4 ? launch_head<Dv, HEAD_LOGITS, 4, false>(a, s) : vec == 2 ? launch_head<Dv, HEAD_LOGITS, 2, false>(a, s) : launch_head<Dv, HEAD_LOGITS, 1, false>(a, s); \
    DML_HEAD_CASE_I(Dv, HEAD_DENSE)

int head_dispatch_1_8(int D, int mode, int vec, bool extra, const HeadArgs& a, cudaStream_t s);
int head_dispatch_9_16(int D, int mode, int vec, bool extra, const HeadArgs& a, cudaStream_t s);
int head_dispatch_17_24(int D, int mode, int vec, bool extra, const HeadArgs& a, cudaStream_t s);
int head_dispatch_25_32(int D, int mode, int vec, bool extra, const HeadArgs& a, cudaStream_t s);

}  // namespace dml
