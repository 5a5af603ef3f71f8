// Synchronised batch normalisation for one-process-per-GPU training: the statistics and
// element-wise kernels of
//   anomaly/lib/nn/modules/batchnorm.py:57-88    forward: sum / square-sum per channel -> (mean, inv_std) of ALL devices ->
//                                                (x - mean) * (inv_std * weight) + bias
//   anomaly/lib/nn/modules/batchnorm.py:121-139  _compute_mean_std: mean = sum / n; sumvar = ssum - sum * mean;
//                                                inv_std = clamp(sumvar / n, eps) ** -0.5; moving averages of mean and the
//                                                unbiased variance through _tmp_running_* / _running_iter
// The reference reduces over the devices of one process (DataParallel master / slave pipes); here every rank computes its
// partial sums, NCCL all-reduces 2C doubles (host side: syncbn.py) and every rank finishes identically.  The backward
// pass is the derivative of exactly that expression, with the cross-rank sums of dy and dy (x - mean) all-reduced the same
// way and the clamp's zero gradient (variance below eps) honoured.
//
// All four kernels stream NCHW fp32 once (HBM-bound: 4 B / element for the statistics, 8 for apply, 12 for backward-apply).
// Partial sums are accumulated in fp32 over 16 elements per thread and in fp64 above that, then combined in a fixed order
// (partials -> second kernel), so results do not depend on the launch order of the CTAs.
#include "dml_common.cuh"

namespace dml {
namespace {

constexpr int BN_THREADS = 256;
constexpr int BN_CHUNK = 4096;         // floats per (CTA, step): 16 per thread

struct BnGeom {
  int B, C;
  long long HW;
  int chunks_per_image;                // ceil(HW / BN_CHUNK)
  int S;                               // CTAs per channel
};

BnGeom bn_geom(int B, int C, long long HW) {
  BnGeom g;
  g.B = B; g.C = C; g.HW = HW;
  g.chunks_per_image = (int)((HW + BN_CHUNK - 1) / BN_CHUNK);
  const long long items = (long long)B * g.chunks_per_image;
  long long s = (148 * 8 + C - 1) / C;                // ~8 CTAs per SM over the whole grid
  if (s > items) s = items;
  if (s < 1) s = 1;
  if (s > 1024) s = 1024;
  g.S = (int)s;
  return g;
}

// BWD = false: (sum x, sum x^2);  BWD = true: (sum dy, sum dy (x - mean_c))
template <bool BWD>
__global__ void __launch_bounds__(BN_THREADS) bn_stats_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                              const float* __restrict__ mean, BnGeom g, double* __restrict__ part) {
  const int c = blockIdx.x, s = blockIdx.y;
  const int tid = threadIdx.x;
  const float m = BWD ? mean[c] : 0.f;
  const long long items = (long long)g.B * g.chunks_per_image;
  const bool vec = (g.HW & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (!BWD || (reinterpret_cast<uintptr_t>(dy) & 15) == 0);
  double a0 = 0.0, a1 = 0.0;
  for (long long it = s; it < items; it += g.S) {
    const int b = (int)(it / g.chunks_per_image);
    const long long i0 = (it - (long long)b * g.chunks_per_image) * BN_CHUNK;
    const long long i1 = min(i0 + BN_CHUNK, g.HW);
    const size_t base = ((size_t)b * g.C + c) * (size_t)g.HW;
    float f0 = 0.f, f1 = 0.f;
    if (vec) {
      for (long long i = i0 + 4 * tid; i < i1; i += 4 * BN_THREADS) {
        const float4 v = *reinterpret_cast<const float4*>(x + base + i);
        if constexpr (BWD) {
          const float4 d = *reinterpret_cast<const float4*>(dy + base + i);
          f0 += (d.x + d.y) + (d.z + d.w);
          f1 += (d.x * (v.x - m) + d.y * (v.y - m)) + (d.z * (v.z - m) + d.w * (v.w - m));
        } else {
          f0 += (v.x + v.y) + (v.z + v.w);
          f1 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
      }
    } else {
      for (long long i = i0 + tid; i < i1; i += BN_THREADS) {
        const float v = x[base + i];
        if constexpr (BWD) {
          const float d = dy[base + i];
          f0 += d;
          f1 += d * (v - m);
        } else {
          f0 += v;
          f1 += v * v;
        }
      }
    }
    a0 += (double)f0;
    a1 += (double)f1;
  }
  __shared__ double s_r[2][BN_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_down_sync(0xffffffffu, a0, o);
    a1 += __shfl_down_sync(0xffffffffu, a1, o);
  }
  if ((tid & 31) == 0) { s_r[0][tid >> 5] = a0; s_r[1][tid >> 5] = a1; }
  __syncthreads();
  if (tid == 0) {
    double t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int w = 0; w < BN_THREADS / 32; ++w) { t0 += s_r[0][w]; t1 += s_r[1][w]; }
    part[((size_t)c * g.S + s) * 2] = t0;
    part[((size_t)c * g.S + s) * 2 + 1] = t1;
  }
}

// sums[c] = sum_s part[c][s][0], sums[C + c] = sum_s part[c][s][1], in slice order
__global__ void bn_stats_reduce_kernel(const double* __restrict__ part, int C, int S, double* __restrict__ sums) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double t0 = 0.0, t1 = 0.0;
  for (int s = 0; s < S; ++s) { t0 += part[((size_t)c * S + s) * 2]; t1 += part[((size_t)c * S + s) * 2 + 1]; }
  sums[c] = t0;
  sums[C + c] = t1;
}

// batchnorm.py:121-139 on the all-reduced sums (fp32 arithmetic of the reference on fp32-rounded sums)
__global__ void bn_finalize_kernel(const double* __restrict__ sums, const double* __restrict__ count, float eps, float keep, int C, float* tmp_mean,
                                   float* tmp_var, float* running_iter, float* running_mean, float* running_var,
                                   float* __restrict__ mean_out, float* __restrict__ inv_std_out, unsigned char* __restrict__ clamped) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float size = (float)count[0];
  const float sum_ = (float)sums[c], ssum = (float)sums[C + c];
  const float mean = __fdiv_rn(sum_, size);
  const float sumvar = __fsub_rn(ssum, __fmul_rn(sum_, mean));
  const float unbias_var = __fdiv_rn(sumvar, size - 1.f);
  const float bias_var = __fdiv_rn(sumvar, size);
  const bool cl = !(bias_var > eps);
  const float v = cl ? eps : bias_var;
  mean_out[c] = mean;
  inv_std_out[c] = __fdiv_rn(1.f, __fsqrt_rn(v));
  clamped[c] = cl ? 1 : 0;
  if (tmp_mean) {
    // every thread reads the OLD iteration count; bn_iter_kernel publishes the new one after this kernel
    const float it_new = __fadd_rn(__fmul_rn(running_iter[0], keep), 1.f);
    const float tm = __fadd_rn(__fmul_rn(tmp_mean[c], keep), mean);
    const float tv = __fadd_rn(__fmul_rn(tmp_var[c], keep), unbias_var);
    tmp_mean[c] = tm;
    tmp_var[c] = tv;
    running_mean[c] = __fdiv_rn(tm, it_new);
    running_var[c] = __fdiv_rn(tv, it_new);
  }
}
__global__ void bn_iter_kernel(float* running_iter, float keep) { running_iter[0] = __fadd_rn(__fmul_rn(running_iter[0], keep), 1.f); }

// y = (x - mean) * (inv_std * weight) + bias        (batchnorm.py:80-84)
__global__ void __launch_bounds__(BN_THREADS) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                              const float* __restrict__ inv_std, const float* __restrict__ weight,
                                                              const float* __restrict__ bias, int C, long long HW, float* __restrict__ y) {
  const int bc = blockIdx.x;
  const int c = bc % C;
  const float m = mean[c];
  const float sc = weight ? __fmul_rn(inv_std[c], weight[c]) : inv_std[c];
  const float sh = bias ? bias[c] : 0.f;
  const size_t base = (size_t)bc * (size_t)HW;
  const bool vec = (HW & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0;
  const long long i0 = (long long)blockIdx.y * BN_CHUNK;
  const long long i1 = min(i0 + BN_CHUNK, HW);
  if (vec) {
    for (long long i = i0 + 4 * threadIdx.x; i < i1; i += 4 * BN_THREADS) {
      const float4 v = *reinterpret_cast<const float4*>(x + base + i);
      float4 o;
      o.x = __fadd_rn(__fmul_rn(__fsub_rn(v.x, m), sc), sh);
      o.y = __fadd_rn(__fmul_rn(__fsub_rn(v.y, m), sc), sh);
      o.z = __fadd_rn(__fmul_rn(__fsub_rn(v.z, m), sc), sh);
      o.w = __fadd_rn(__fmul_rn(__fsub_rn(v.w, m), sc), sh);
      *reinterpret_cast<float4*>(y + base + i) = o;
    }
  } else {
    for (long long i = i0 + threadIdx.x; i < i1; i += BN_THREADS) y[base + i] = __fadd_rn(__fmul_rn(__fsub_rn(x[base + i], m), sc), sh);
  }
}

// dx = w inv_std [dy - sum(dy) / n - (x - mean) inv_std^2 sum(dy (x - mean)) / n]   (last term 0 where the variance was clamped)
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                  const float* __restrict__ mean, const float* __restrict__ inv_std,
                                                                  const float* __restrict__ weight, const unsigned char* __restrict__ clamped,
                                                                  const double* __restrict__ sums, const double* __restrict__ count_p, int C, long long HW,
                                                                  float* __restrict__ dx) {
  const double count = count_p[0];
  const int bc = blockIdx.x;
  const int c = bc % C;
  const float m = mean[c], is = inv_std[c];
  const float g = weight ? is * weight[c] : is;
  const float k0 = (float)(sums[c] / count);
  const float k1 = clamped[c] ? 0.f : (float)(sums[C + c] / count) * is * is;
  const size_t base = (size_t)bc * (size_t)HW;
  const bool vec = (HW & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(dx) & 15) == 0;
  const long long i0 = (long long)blockIdx.y * BN_CHUNK;
  const long long i1 = min(i0 + BN_CHUNK, HW);
  if (vec) {
    for (long long i = i0 + 4 * threadIdx.x; i < i1; i += 4 * BN_THREADS) {
      const float4 v = *reinterpret_cast<const float4*>(x + base + i);
      const float4 d = *reinterpret_cast<const float4*>(dy + base + i);
      float4 o;
      o.x = g * (d.x - k0 - (v.x - m) * k1);
      o.y = g * (d.y - k0 - (v.y - m) * k1);
      o.z = g * (d.z - k0 - (v.z - m) * k1);
      o.w = g * (d.w - k0 - (v.w - m) * k1);
      *reinterpret_cast<float4*>(dx + base + i) = o;
    }
  } else {
    for (long long i = i0 + threadIdx.x; i < i1; i += BN_THREADS) dx[base + i] = g * (dy[base + i] - k0 - (x[base + i] - m) * k1);
  }
}

}  // namespace
}  // namespace dml

using namespace dml;

extern "C" {
#pragma GCC visibility push(default)

size_t dml_bn_workspace_bytes(int32_t B, int32_t C, int64_t HW) {
  if (B < 1 || C < 1 || HW < 1) return 256;
  const BnGeom g = bn_geom(B, C, HW);
  return (size_t)C * g.S * 2 * sizeof(double) + 256;
}

int dml_bn_stats(const float* x, const float* dy, const float* mean, int32_t B, int32_t C, int64_t HW, double* sums, void* workspace,
                 size_t workspace_bytes, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!sums || B < 0 || C < 1 || HW < 0 || C > 65535 || (dy != nullptr) != (mean != nullptr)) return DML_ERR_INVALID_ARG;
  if (B == 0 || HW == 0) {
    DML_CUDA_TRY(cudaMemsetAsync(sums, 0, (size_t)2 * C * sizeof(double), stream));
    return DML_OK;
  }
  if (!x || !workspace) return DML_ERR_INVALID_ARG;
  if (workspace_bytes < dml_bn_workspace_bytes(B, C, HW)) return DML_ERR_WORKSPACE;
  const BnGeom g = bn_geom(B, C, HW);
  double* part = reinterpret_cast<double*>(workspace);
  dim3 grid((unsigned)C, (unsigned)g.S);
  if (dy) bn_stats_kernel<true><<<grid, BN_THREADS, 0, stream>>>(x, dy, mean, g, part);
  else bn_stats_kernel<false><<<grid, BN_THREADS, 0, stream>>>(x, nullptr, nullptr, g, part);
  DML_LAUNCH_CHECK();
  bn_stats_reduce_kernel<<<(C + 127) / 128, 128, 0, stream>>>(part, C, g.S, sums);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

int dml_bn_finalize(const double* sums, const double* count, float eps, float momentum, int32_t C, float* tmp_running_mean, float* tmp_running_var,
                    float* running_iter, float* running_mean, float* running_var, float* mean, float* inv_std, uint8_t* clamped,
                    dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!sums || !count || !mean || !inv_std || !clamped || C < 1) return DML_ERR_INVALID_ARG;
  const bool track = tmp_running_mean != nullptr;
  if (track && (!tmp_running_var || !running_iter || !running_mean || !running_var)) return DML_ERR_INVALID_ARG;
  const float keep = 1.f - momentum;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(sums, count, eps, keep, C, tmp_running_mean, tmp_running_var, running_iter,
                                                        running_mean, running_var, mean, inv_std, clamped);
  DML_LAUNCH_CHECK();
  if (track) {
    bn_iter_kernel<<<1, 1, 0, stream>>>(running_iter, keep);
    DML_LAUNCH_CHECK();
  }
  return DML_OK;
}

int dml_bn_apply(const float* x, const float* mean, const float* inv_std, const float* weight, const float* bias, int32_t B, int32_t C,
                 int64_t HW, float* y, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!mean || !inv_std || B < 0 || C < 1 || HW < 0 || (HW + BN_CHUNK - 1) / BN_CHUNK > 65535) return DML_ERR_INVALID_ARG;
  if (B == 0 || HW == 0) return DML_OK;
  if (!x || !y) return DML_ERR_INVALID_ARG;
  dim3 grid((unsigned)((long long)B * C), (unsigned)((HW + BN_CHUNK - 1) / BN_CHUNK));
  bn_apply_kernel<<<grid, BN_THREADS, 0, stream>>>(x, mean, inv_std, weight, bias, C, HW, y);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

int dml_bn_backward_apply(const float* x, const float* dy, const float* mean, const float* inv_std, const float* weight,
                          const uint8_t* clamped, const double* sums, const double* count, int32_t B, int32_t C, int64_t HW, float* dx,
                          dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!mean || !inv_std || !clamped || !sums || !count || B < 0 || C < 1 || HW < 0 || (HW + BN_CHUNK - 1) / BN_CHUNK > 65535)
    return DML_ERR_INVALID_ARG;
  if (B == 0 || HW == 0) return DML_OK;
  if (!x || !dy || !dx) return DML_ERR_INVALID_ARG;
  dim3 grid((unsigned)((long long)B * C), (unsigned)((HW + BN_CHUNK - 1) / BN_CHUNK));
  bn_bwd_apply_kernel<<<grid, BN_THREADS, 0, stream>>>(x, dy, mean, inv_std, weight, clamped, sums, count, C, HW, dx);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

#pragma GCC visibility pop
}  // extern "C"
