// C-ABI entry points of the loss (b) and class-sum (c) paths.  Kernels: dml_loss.cuh, dml_reduce.cuh.
#include "dml_loss.cuh"
#include "dml_reduce.cuh"

namespace dml {
namespace {

int loss_grid_x(long long hw, int vec) {
  long long gx = (hw + (long long)LOSS_THREADS * vec - 1) / ((long long)LOSS_THREADS * vec);
  if (gx > LOSS_MAX_BLOCKS_X) gx = LOSS_MAX_BLOCKS_X;
  return gx < 1 ? 1 : (int)gx;
}

// fixed-order reduction of the per-block partials -> (loss, CE, VL, Inter, n_valid).  One CTA of 1024 threads; thread t
// takes the blocks t, t + 1024, ... with four 32-byte entries in flight (the first version walked a contiguous range per
// thread one dependent load at a time: 27 us for 11520 blocks, 15 % of the forward pass), then a fixed shared-memory tree.
constexpr int LOSS_FIN_THREADS = 1024;
__global__ void __launch_bounds__(LOSS_FIN_THREADS) loss_finalize_kernel(const double* __restrict__ partials, int n_blocks, int B,
                                                                         long long hw, double alpha, double beta, double* out5) {
  __shared__ double s_r[4][LOSS_FIN_THREADS];
  const int tid = threadIdx.x;
  double r[4] = {0.0, 0.0, 0.0, 0.0};
  const double2* p2 = reinterpret_cast<const double2*>(partials);     // (the workspace is 16-byte aligned: 4 doubles per block)
  for (int i0 = tid; i0 < n_blocks; i0 += 4 * LOSS_FIN_THREADS) {
    double2 v[4][2];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * LOSS_FIN_THREADS;
      v[u][0] = i < n_blocks ? p2[(size_t)i * 2] : make_double2(0.0, 0.0);
      v[u][1] = i < n_blocks ? p2[(size_t)i * 2 + 1] : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { r[0] += v[u][0].x; r[1] += v[u][0].y; r[2] += v[u][1].x; r[3] += v[u][1].y; }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) s_r[j][tid] = r[j];
  __syncthreads();
  for (int o = LOSS_FIN_THREADS / 2; o > 0; o >>= 1) {
    if (tid < o)
#pragma unroll
      for (int j = 0; j < 4; ++j) s_r[j][tid] += s_r[j][tid + o];
    __syncthreads();
  }
  if (tid == 0) {
    const double nv = s_r[3][0];
    const double ce = s_r[0][0] / nv;  // mean over valid pixels (NaN when there is none, like torch)
    const double vl = s_r[1][0] / (double)hw;
    const double in = s_r[2][0] / (double)hw;
    out5[0] = (ce + alpha * vl + beta * in) / (double)B;
    out5[1] = ce;
    out5[2] = vl;
    out5[3] = in;
    out5[4] = nv;
  }
}

int reduce_grid_x(long long hw) {
  long long gx = (hw + RED_THREADS * 8 - 1) / (RED_THREADS * 8);
  if (gx > 64) gx = 64;
  return gx < 1 ? 1 : (int)gx;
}

// sums[b,c,d] = sum over blocks (fixed order); counts likewise
__global__ void __launch_bounds__(256) class_sums_finalize_kernel(const double* __restrict__ partials, int gx, int n_cls,
                                                                  int D, double* sums, long long* counts) {
  const int b = blockIdx.y;
  const int row = D + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over n_cls * row
  if (i >= n_cls * row) return;
  // eight independent loads in flight, combined in a fixed order (one dependent load at a time: 33 us for 64 blocks,
  // 17 % of the whole reduction)
  double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int g0 = 0; g0 < gx; g0 += 8) {
    double v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = g0 + q < gx ? partials[((size_t)b * gx + g0 + q) * n_cls * row + i] : 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] += v[q];
  }
  const double t = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  const int c = i / row, d = i - c * row;
  if (d < D) sums[((size_t)b * n_cls + c) * D + d] = t;
  else counts[(size_t)b * n_cls + c] = (long long)(t + 0.5);
}

int loss_common(bool bwd, bool is_logits, const float* x, const float* mu, float diag_m, const uint8_t* t_u8, const int64_t* t_i64,
                int64_t ignore_index, int B, int D, int K, int H, int W, double alpha, double beta, void* partials,
                double* out5, const float* grad_out, float* dx, cudaStream_t stream) {
  if (!x || (!t_u8 == !t_i64) || B < 1 || D < 1 || K < 1 || H < 1 || W < 1 || !out5) return DML_ERR_INVALID_ARG;
  if (D > DML_MAX_DIM || K > DML_MAX_DIM) return DML_ERR_UNSUPPORTED_DIM;
  if (B > 65535) return DML_ERR_INVALID_ARG;
  if (is_logits && mu) return DML_ERR_INVALID_ARG;
  const int mode = is_logits ? LOSS_LOGITS : (mu == nullptr ? LOSS_IDENT : LOSS_DENSE);
  const bool ident = mode != LOSS_DENSE;  // vectorised, K == D paths
  if (ident && K != D) return DML_ERR_INVALID_ARG;
  if (bwd ? (!grad_out || !dx) : !partials) return DML_ERR_INVALID_ARG;
  if (!bwd && (reinterpret_cast<uintptr_t>(partials) & 15) != 0) return DML_ERR_INVALID_ARG;   // read back as double2
  const long long hw = (long long)H * W;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  // (vector target loads: uint8 targets 4-byte aligned, int64 targets 16-byte aligned)
  const bool t_al = t_u8 ? (reinterpret_cast<uintptr_t>(t_u8) & 3) == 0 : al16(t_i64);
  int vec = (ident && hw % 4 == 0 && al16(x) && t_al && (!bwd || al16(dx))) ? 4 : 1;
  if (D > 24) vec = 1;
  LossArgs a;
  a.x = x; a.mu = mu; a.diag_m = diag_m; a.t_u8 = t_u8; a.t_i64 = (const long long*)t_i64; a.ignore = ignore_index;
  a.B = B; a.K = K; a.HW = hw; a.alpha = alpha; a.beta = beta;
  a.partials = reinterpret_cast<double*>(partials); a.out5 = out5; a.grad_out = grad_out; a.dx = dx;
  const int gx = loss_grid_x(hw, vec);
  int rc;
  if (D <= 8) rc = loss_dispatch_1_8(D, mode, vec, bwd, a, gx, stream);
  else if (D <= 16) rc = loss_dispatch_9_16(D, mode, vec, bwd, a, gx, stream);
  else if (D <= 24) rc = loss_dispatch_17_24(D, mode, vec, bwd, a, gx, stream);
  else rc = loss_dispatch_25_32(D, mode, vec, bwd, a, gx, stream);
  if (rc != DML_OK || bwd) return rc;
  loss_finalize_kernel<<<1, LOSS_FIN_THREADS, 0, stream>>>(a.partials, gx * B, B, hw, alpha, beta, out5);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

}  // namespace
}  // namespace dml

using namespace dml;

extern "C" {
#pragma GCC visibility push(default)

size_t dml_loss_workspace_bytes(int32_t B, int32_t H, int32_t W) {
  if (B < 1 || H < 1 || W < 1) return 256;
  // vec = 1 gives the largest grid
  return (size_t)B * loss_grid_x((long long)H * W, 1) * 4 * sizeof(double);
}

int dml_loss_forward(const float* x, int32_t x_is_logits, const float* mu, float diag_m, const uint8_t* target_u8,
                     const int64_t* target_i64, int64_t ignore_index, int32_t B, int32_t D, int32_t K, int32_t H, int32_t W,
                     double alpha, double beta, void* partials, double* out5, dml_stream_t stream) {
  return loss_common(false, x_is_logits != 0, x, mu, diag_m, target_u8, target_i64, ignore_index, B, D, K, H, W, alpha, beta, partials, out5,
                     nullptr, nullptr, (cudaStream_t)stream);
}

int dml_loss_backward(const float* x, int32_t x_is_logits, const float* mu, float diag_m, const uint8_t* target_u8,
                      const int64_t* target_i64, int64_t ignore_index, int32_t B, int32_t D, int32_t K, int32_t H, int32_t W,
                      double alpha, double beta, const double* out5, const float* grad_out, float* dx, dml_stream_t stream) {
  return loss_common(true, x_is_logits != 0, x, mu, diag_m, target_u8, target_i64, ignore_index, B, D, K, H, W, alpha, beta, nullptr,
                     const_cast<double*>(out5), grad_out, dx, (cudaStream_t)stream);
}

size_t dml_class_sums_workspace_bytes(int32_t B, int32_t D, int32_t n_cls, int64_t pixels_per_image) {
  if (B < 1 || D < 1 || n_cls < 1 || pixels_per_image < 1) return 256;
  return (size_t)B * reduce_grid_x(pixels_per_image) * n_cls * (D + 1) * sizeof(double);
}

int dml_class_sums(const float* x, int32_t nhwc, const uint8_t* label_u8, const int64_t* label_i64, int32_t B, int32_t D,
                   int64_t hw, int32_t n_cls, void* workspace, double* sums, long long* counts, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!x || (!label_u8 == !label_i64) || !workspace || !sums || !counts) return DML_ERR_INVALID_ARG;
  if (B < 1 || D < 1 || n_cls < 1 || hw < 1 || B > 65535) return DML_ERR_INVALID_ARG;
  if (D > DML_MAX_DIM || n_cls > 64) return DML_ERR_UNSUPPORTED_DIM;
  ReduceArgs a;
  a.x = x; a.nhwc = nhwc; a.l_u8 = label_u8; a.l_i64 = (const long long*)label_i64;
  a.B = B; a.n_cls = n_cls; a.HW = hw; a.partials = reinterpret_cast<double*>(workspace);
  const int gx = reduce_grid_x(hw);
  int rc;
  if (D <= 8) rc = reduce_dispatch_1_8(D, a, gx, stream);
  else if (D <= 16) rc = reduce_dispatch_9_16(D, a, gx, stream);
  else if (D <= 24) rc = reduce_dispatch_17_24(D, a, gx, stream);
  else rc = reduce_dispatch_25_32(D, a, gx, stream);
  if (rc != DML_OK) return rc;
  const int n = n_cls * (D + 1);
  class_sums_finalize_kernel<<<dim3((unsigned)ceil_div_i(n, 256), (unsigned)B), 256, 0, stream>>>(a.partials, gx, n_cls, D, sums, counts);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

#pragma GCC visibility pop
}  // extern "C"
