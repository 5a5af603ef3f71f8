// C-ABI entry points of the head path: dml_head_forward, dml_scores_finalize, dml_confusion,
// dml_plm_merge, plus library-level helpers.  Kernel body lives in dml_head.cuh.
#include <cmath>
#include <cstdlib>
#include "dml_head.cuh"

namespace dml {

thread_local int g_last_cuda_error = 0;
unsigned long long g_kernel_launches = 0;

__global__ void minmax_init_kernel(int* mm, int n4) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) mm[i] = (i & 1) ? 0 /* max := +0.0f */ : 0x7f800000 /* min := +inf */;
}

// Per-image min-max normalisation + EDS/MMSP mix.  fp32 arithmetic mirrors NumPy's:
// (v - lo) / (hi - lo) with separately rounded sub/sub/div; c = 1/(1+exp(lambda*(e-thr))).
// anomaly/eval_ood_traditional.py:101-106,305,435,447-448.
__global__ void __launch_bounds__(256) finalize_kernel(const float* __restrict__ eds, const float* __restrict__ msp,
                                                       const float* __restrict__ minmax, long long hw, float lambda,
                                                       float thr, int complement, float* eds_n, float* msp_n, float* mix) {
  const int b = blockIdx.y;
  const float elo = minmax[b * 4 + 0], ehi = minmax[b * 4 + 1];
  const float mlo = minmax[b * 4 + 2], mhi = minmax[b * 4 + 3];
  const float eden = __fsub_rn(ehi, elo), mden = __fsub_rn(mhi, mlo);
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += (long long)gridDim.x * blockDim.x) {
    const long long i = (long long)b * hw + p;
    float e = 0.f, m = 0.f;
    if (eds) e = __fdiv_rn(__fsub_rn(eds[i], elo), eden);
    if (msp) m = __fdiv_rn(__fsub_rn(msp[i], mlo), mden);
    if (mix) {
      // NumPy: 1 / (1 + np.exp(lamda * (x - thre))) in float32
      const float c = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(__fmul_rn(lambda, __fsub_rn(e, thr)))));
      mix[i] = __fadd_rn(__fmul_rn(c, e), __fmul_rn(__fsub_rn(1.0f, c), m));
    }
    if (eds_n) eds_n[i] = complement ? __fsub_rn(1.0f, e) : e;
    if (msp_n) msp_n[i] = m;
  }
}

template <typename GT, typename PR>
__global__ void __launch_bounds__(256) confusion_kernel(const GT* __restrict__ gt, const PR* __restrict__ pred, long long n,
                                                        int rows, int cols, unsigned long long* out) {
  extern __shared__ unsigned int s_bins[];
  const int nbins = rows * cols;
  for (int i = threadIdx.x; i < nbins; i += blockDim.x) s_bins[i] = 0u;
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long start = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // uniform trip count per warp so that __match_any_sync sees all 32 lanes
  const long long iters = (n + stride - 1) / stride;
  for (long long it = 0; it < iters; ++it) {
    const long long i = start + it * stride;
    int bin = -1;
    if (i < n) {
      const long long g = (long long)gt[i];
      const long long p = (long long)pred[i];
      if (g >= 0 && g < rows && p >= 0 && p < cols) bin = (int)(g * cols + p);
    }
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (bin >= 0 && (threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&s_bins[bin], (unsigned)__popc(peers));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nbins; i += blockDim.x) {
    const unsigned c = s_bins[i];
    if (c) atomicAdd(out + i, (unsigned long long)c);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) plm_merge_kernel(T* base, const T* __restrict__ head, long long n, int novel) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if ((long long)head[i] == novel) base[i] = (T)novel;
}

static int pick_vec(const dml_head_params* p, long long hw) {
  auto al = [](const void* q, size_t a) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % a) == 0; };
  // measured on B200 (profiles/): 2 pixels/thread (48-56 registers, 4 resident CTAs/SM) beats 4 for
  // D > 8 because load and compute phases of more co-resident CTAs overlap; small D keeps 4.
  int vmax = p->D > 8 ? 2 : 4;
  if (const char* e = getenv("DML_HEAD_VEC")) {  // tuning knob: cap the pixels per thread
    const int v = atoi(e);
    if (v == 1 || v == 2 || v == 4) vmax = v;
  }
  for (int vec = vmax; vec >= 2; vec >>= 1) {
    if (hw % vec) continue;
    const size_t fa = 4 * (size_t)vec;
    if (!al(p->x, fa) || !al(p->logits, fa) || !al(p->maxlogit, fa) || !al(p->eds, fa) || !al(p->msp, fa)) continue;
    if (!al(p->label_u8, vec) || !al(p->label_i64, 16) || !al(p->gt_u8, vec)) continue;
    // very wide embeddings: keep the register footprint (D * VEC floats) bounded
    if (p->D * vec > 96) continue;
    return vec;
  }
  return 1;
}

}  // namespace dml

using namespace dml;

extern "C" {
#pragma GCC visibility push(default)

int dml_abi_version(void) { return DML_B200_ABI_VERSION; }
int dml_max_dim(void) { return DML_MAX_DIM; }
int dml_last_cuda_error(void) { return g_last_cuda_error; }
unsigned long long dml_kernel_launches(void) { return g_kernel_launches; }

const char* dml_error_string(int code) {
  switch (code) {
    case DML_OK: return "ok";
    case DML_ERR_INVALID_ARG: return "invalid argument";
    case DML_ERR_UNSUPPORTED_DIM: return "embedding dim / class count outside the compiled range";
    case DML_ERR_CUDA: return "CUDA runtime error";
    case DML_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown error";
  }
}

int dml_head_forward(const dml_head_params* p, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!p || p->struct_bytes != sizeof(dml_head_params)) return DML_ERR_INVALID_ARG;
  if (p->B < 0 || p->D < 1 || p->K < 1 || p->H < 0 || p->W < 0) return DML_ERR_INVALID_ARG;
  if (p->B == 0 || p->H == 0 || p->W == 0) return DML_OK;  // empty batch / image: nothing to do
  if (!p->x) return DML_ERR_INVALID_ARG;
  if (p->D > DML_MAX_DIM || p->K > 255) return DML_ERR_UNSUPPORTED_DIM;
  const int mode = p->input_is_logits ? HEAD_LOGITS : (p->mu == nullptr ? HEAD_IDENT : HEAD_DENSE);
  if (mode != HEAD_DENSE && p->K != p->D) return DML_ERR_INVALID_ARG;
  if (mode == HEAD_LOGITS && (p->mu || p->n_novel > 0 || p->features_nhwc || p->logits || p->novel_dist)) return DML_ERR_INVALID_ARG;
  if (p->score_first_class < 0 || p->score_first_class > 1 || p->score_first_class >= p->K) return DML_ERR_INVALID_ARG;
  if (p->n_novel < 0 || p->n_novel > HEAD_MAX_NOVEL || (p->n_novel > 0 && !p->mu_novel)) return DML_ERR_INVALID_ARG;
  if (p->n_novel > 0 && p->novel_label_base + p->n_novel > 256) return DML_ERR_INVALID_ARG;
  if ((p->want_eds_minmax || p->want_msp_minmax) && !p->minmax) return DML_ERR_INVALID_ARG;
  if (p->confusion) {
    if (!p->gt_u8 == !p->gt_i64) return DML_ERR_INVALID_ARG;
    if (p->conf_rows < 1 || p->conf_cols < 1 || p->conf_rows * p->conf_cols > HEAD_MAX_CONF_BINS) return DML_ERR_INVALID_ARG;
  }
  if (p->B > 65535) return DML_ERR_INVALID_ARG;
  if (p->reference_order && (mode != HEAD_IDENT || p->D >= 16 || !p->logits)) return DML_ERR_INVALID_ARG;
  const long long hw = (long long)p->H * p->W;
  if (p->B == 0 || hw == 0) return DML_OK;

  HeadArgs a = {};
  a.x = p->x; a.mu = p->mu; a.diag_m = p->diag_m; a.ref_order = p->reference_order ? 1 : 0;
  a.msp_scale = 2.0f * p->diag_m * 1.4426950408889634f;
  a.first = p->score_first_class; a.clamp = p->eds_clamp;
  a.mu_novel = p->mu_novel; a.n_novel = p->n_novel; a.novel_base = p->novel_label_base; a.novel_thr = p->novel_thr;
  a.logits = p->logits; a.label_u8 = p->label_u8; a.label_i64 = (long long*)p->label_i64;
  a.maxlogit = p->maxlogit; a.eds = p->eds; a.msp = p->msp; a.feat = p->features_nhwc; a.novel_dist = p->novel_dist;
  a.minmax = reinterpret_cast<int*>(p->minmax);
  a.want_eds_mm = p->want_eds_minmax; a.want_msp_mm = p->want_msp_minmax;
  a.gt_u8 = p->gt_u8; a.gt_i64 = (const long long*)p->gt_i64; a.conf = p->confusion;
  a.crow = p->conf_rows; a.ccol = p->conf_cols;
  a.B = p->B; a.K = p->K; a.HW = hw;
  a.out_mask = (a.label_u8 ? OUT_LABEL_U8 : 0u) | (a.label_i64 ? OUT_LABEL_I64 : 0u) | (a.maxlogit ? OUT_MAXLOGIT : 0u) |
               (a.eds ? OUT_EDS : 0u) | (a.msp ? OUT_MSP : 0u) |
               ((a.minmax && (a.want_eds_mm || a.want_msp_mm)) ? OUT_MINMAX : 0u) | (a.conf ? OUT_CONF : 0u) |
               (a.gt_u8 ? OUT_GT_U8 : 0u) | (a.logits ? OUT_LOGITS : 0u) | (a.feat ? OUT_FEAT : 0u) |
               (a.novel_dist ? OUT_NOVEL_DIST : 0u);

  if (a.minmax && (a.want_eds_mm || a.want_msp_mm)) {
    const int n4 = p->B * 4;
    minmax_init_kernel<<<ceil_div_i(n4, 256), 256, 0, stream>>>(a.minmax, n4);
    DML_LAUNCH_CHECK();
  }
  bool extra = a.n_novel > 0 || a.feat != nullptr || a.novel_dist != nullptr || a.logits != nullptr;
  // DML_HEAD_LEAN=0: score mu = m*I inputs through the per-class-distance instantiation instead of the closed-form
  // lean path (cross-check / A-B timing knob; read per call, no state kept)
  if (mode == HEAD_IDENT && !extra) {
    const char* e = getenv("DML_HEAD_LEAN");
    if (e && e[0] == '0') extra = true;
  }
  int vec = pick_vec(p, hw);
  if (extra && vec > 2) vec = 2;
  const int D = p->D;
  if (D <= 8) return head_dispatch_1_8(D, mode, vec, extra, a, stream);
  if (D <= 16) return head_dispatch_9_16(D, mode, vec, extra, a, stream);
  if (D <= 24) return head_dispatch_17_24(D, mode, vec, extra, a, stream);
  return head_dispatch_25_32(D, mode, vec, extra, a, stream);
}

int dml_multiscale_head_forward(const dml_multiscale_params* p, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!p || p->struct_bytes != sizeof(dml_multiscale_params)) return DML_ERR_INVALID_ARG;
  if (p->B < 0 || p->K < 1 || p->H < 0 || p->W < 0) return DML_ERR_INVALID_ARG;
  if (p->n_scales < 1 || p->n_scales > HEAD_MAX_SCALES) return DML_ERR_INVALID_ARG;
  if (p->K > DML_MAX_DIM) return DML_ERR_UNSUPPORTED_DIM;
  if (p->score_first_class < 0 || p->score_first_class > 1 || p->score_first_class >= p->K) return DML_ERR_INVALID_ARG;
  if ((p->want_eds_minmax || p->want_msp_minmax) && !p->minmax) return DML_ERR_INVALID_ARG;
  if (p->confusion) {
    if (!p->gt_u8 == !p->gt_i64) return DML_ERR_INVALID_ARG;
    if (p->conf_rows < 1 || p->conf_cols < 1 || p->conf_rows * p->conf_cols > HEAD_MAX_CONF_BINS) return DML_ERR_INVALID_ARG;
  }
  if (p->B > 65535) return DML_ERR_INVALID_ARG;
  if (p->B == 0 || p->H == 0 || p->W == 0) return DML_OK;
  for (int s = 0; s < p->n_scales; ++s) {
    if (!p->z[s] || p->h[s] < 1 || p->w[s] < 1) return DML_ERR_INVALID_ARG;
    if ((long long)p->K * p->h[s] * p->w[s] > 0x7fffffffLL) return DML_ERR_INVALID_ARG;
  }
  const long long hw = (long long)p->H * p->W;

  HeadArgs a = {};
  a.first = p->score_first_class; a.clamp = p->eds_clamp;
  a.logits = p->scores; a.label_u8 = p->label_u8; a.label_i64 = (long long*)p->label_i64;
  a.maxlogit = p->maxlogit; a.eds = p->eds; a.msp = p->msp;
  a.minmax = reinterpret_cast<int*>(p->minmax);
  a.want_eds_mm = p->want_eds_minmax; a.want_msp_mm = p->want_msp_minmax;
  a.gt_u8 = p->gt_u8; a.gt_i64 = (const long long*)p->gt_i64; a.conf = p->confusion;
  a.crow = p->conf_rows; a.ccol = p->conf_cols;
  a.B = p->B; a.K = p->K; a.HW = hw;
  a.out_mask = (a.label_u8 ? OUT_LABEL_U8 : 0u) | (a.label_i64 ? OUT_LABEL_I64 : 0u) | (a.maxlogit ? OUT_MAXLOGIT : 0u) |
               (a.eds ? OUT_EDS : 0u) | (a.msp ? OUT_MSP : 0u) |
               ((a.minmax && (a.want_eds_mm || a.want_msp_mm)) ? OUT_MINMAX : 0u) | (a.conf ? OUT_CONF : 0u) |
               (a.gt_u8 ? OUT_GT_U8 : 0u) | (a.logits ? OUT_LOGITS : 0u);
  a.ms_n = p->n_scales; a.ms_W = p->W; a.ms_recip = p->reciprocal_average ? 1 : 0;
  a.ms_div = (float)p->n_scales; a.ms_inv = 1.0f / (float)p->n_scales;
  for (int s = 0; s < p->n_scales; ++s) {
    a.ms_z[s] = p->z[s]; a.ms_h[s] = p->h[s]; a.ms_w[s] = p->w[s];
    // torch area_pixel_compute_scale<float>(in, out, align_corners=false, nullopt) = float(in) / out
    a.ms_rh[s] = (float)p->h[s] / (float)p->H;
    a.ms_rw[s] = (float)p->w[s] / (float)p->W;
  }
  if (a.minmax && (a.want_eds_mm || a.want_msp_mm)) {
    const int n4 = p->B * 4;
    minmax_init_kernel<<<ceil_div_i(n4, 256), 256, 0, stream>>>(a.minmax, n4);
    DML_LAUNCH_CHECK();
  }
  // VEC horizontally adjacent pixels per thread must stay inside one output row
  auto al = [](const void* q, size_t n) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % n) == 0; };
  int vec = 1;
  if (p->W % 2 == 0 && al(p->scores, 8) && al(p->maxlogit, 8) && al(p->eds, 8) && al(p->msp, 8) && al(p->label_u8, 2) &&
      al(p->label_i64, 16) && al(p->gt_u8, 2))
    vec = 2;
  if (const char* e = getenv("DML_MS_VEC")) {
    if (atoi(e) == 1) vec = 1;
  }
  const int D = p->K;
  // staged variant: the low-resolution footprint of a 16 x (64*vec) output tile, all scales, must fit in the
  // default 48 KB of dynamic shared memory (always true for up-sampling by >= ~3x); otherwise gather from L1/L2
  int mode = HEAD_MSS;
  {
    const int tile_w = MS_TILE_THREADS_X * vec;
    long long floats = 0;
    for (int s = 0; s < p->n_scales; ++s) {
      // rows spanned by the taps of MS_TILE_ROWS consecutive outputs: ceil(scale*(rows-1)) + 2, +1 for fp32 rounding
      const int fh = (int)fmin((double)p->h[s], ceil((double)a.ms_rh[s] * (MS_TILE_ROWS - 1)) + 3.0);
      const int fw = (int)fmin((double)p->w[s], ceil((double)a.ms_rw[s] * (tile_w - 1)) + 3.0);
      a.ms_fh[s] = fh; a.ms_fw[s] = fw; a.ms_soff[s] = (int)floats;
      floats += (long long)fh * fw * ms_stride(D);
    }
    const size_t conf_bytes = p->confusion ? (((size_t)p->conf_rows * p->conf_cols * sizeof(unsigned) + 15) & ~(size_t)15) : 0;
    if (floats * 4 + (long long)conf_bytes > 48 * 1024 - 64) mode = HEAD_MS;
    a.ms_smem_floats = (int)floats;
    a.ms_H = p->H;
    if (const char* e = getenv("DML_MS_DIRECT")) {
      if (atoi(e) == 1) mode = HEAD_MS;
    }
  }
  if (D <= 8) return head_dispatch_1_8(D, mode, vec, false, a, stream);
  if (D <= 16) return head_dispatch_9_16(D, mode, vec, false, a, stream);
  if (D <= 24) return head_dispatch_17_24(D, mode, vec, false, a, stream);
  return head_dispatch_25_32(D, mode, vec, false, a, stream);
}

int dml_scores_finalize(const float* eds, const float* msp, const float* minmax, int32_t B, int64_t hw, float lambda,
                        float thr, int32_t complement, float* eds_norm, float* msp_norm, float* mix,
                        dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!minmax || B < 0 || hw < 0 || B > 65535) return DML_ERR_INVALID_ARG;
  if ((eds_norm && !eds) || (msp_norm && !msp) || (mix && (!eds || !msp))) return DML_ERR_INVALID_ARG;
  if (B == 0 || hw == 0) return DML_OK;
  const int gx = (int)((hw + 256 * 4 - 1) / (256 * 4));
  dim3 grid(gx < 1 ? 1 : (gx > 4096 ? 4096 : gx), B);
  finalize_kernel<<<grid, 256, 0, stream>>>(eds, msp, minmax, hw, lambda, thr, complement, eds_norm, msp_norm, mix);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

int dml_confusion(const uint8_t* gt_u8, const int64_t* gt_i64, const uint8_t* pred_u8, const int64_t* pred_i64,
                  int64_t n, int32_t rows, int32_t cols, unsigned long long* confusion, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!confusion || (!gt_u8 == !gt_i64) || (!pred_u8 == !pred_i64) || n < 0) return DML_ERR_INVALID_ARG;
  if (rows < 1 || cols < 1 || rows * cols > 64 * 64) return DML_ERR_INVALID_ARG;
  if (n == 0) return DML_OK;
  long long blocks = (n + 256 * 16 - 1) / (256 * 16);
  if (blocks > 148 * 16) blocks = 148 * 16;
  const size_t smem = (size_t)rows * cols * sizeof(unsigned);
  const long long* g64 = (const long long*)gt_i64;
  const long long* p64 = (const long long*)pred_i64;
  if (gt_u8 && pred_u8) confusion_kernel<<<(int)blocks, 256, smem, stream>>>(gt_u8, pred_u8, (long long)n, rows, cols, confusion);
  else if (gt_u8) confusion_kernel<<<(int)blocks, 256, smem, stream>>>(gt_u8, p64, (long long)n, rows, cols, confusion);
  else if (pred_u8) confusion_kernel<<<(int)blocks, 256, smem, stream>>>(g64, pred_u8, (long long)n, rows, cols, confusion);
  else confusion_kernel<<<(int)blocks, 256, smem, stream>>>(g64, p64, (long long)n, rows, cols, confusion);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

int dml_plm_merge(uint8_t* base_u8, int64_t* base_i64, const uint8_t* head_u8, const int64_t* head_i64, int64_t n,
                  int32_t novel_label, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0) return DML_ERR_INVALID_ARG;
  const bool u8 = base_u8 && head_u8 && !base_i64 && !head_i64;
  const bool i64 = base_i64 && head_i64 && !base_u8 && !head_u8;
  if (!u8 && !i64) return DML_ERR_INVALID_ARG;
  if (n == 0) return DML_OK;
  long long blocks = (n + 256 * 8 - 1) / (256 * 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (u8) plm_merge_kernel<<<(int)blocks, 256, 0, stream>>>(base_u8, head_u8, (long long)n, novel_label);
  else plm_merge_kernel<<<(int)blocks, 256, 0, stream>>>((long long*)base_i64, (const long long*)head_i64, (long long)n, novel_label);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

#pragma GCC visibility pop
}  // extern "C"
