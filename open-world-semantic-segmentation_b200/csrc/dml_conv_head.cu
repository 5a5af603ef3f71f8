// Final 1x1 classifier conv fused into the distance head.
//
// Replaces, for inference, the last layer of the reference's decoders together with the distance block that follows it:
//   anomaly/models/models.py:609        nn.Conv2d(512, num_class, kernel_size=1)      (conv_last[4])
//   anomaly/models/models.py:636-657    z_k = -sum_d (emb_d - 3 [d = k])^2
//   DeepLabV3Plus-Pytorch/network/utils.py:23   nn.Conv2d(256, num_classes, 1)         (DeepLabHeadV3Plus.classifier[3])
// One thread owns one stride-8 (stride-4) pixel: it streams the C input channels of that pixel (NCHW: for a fixed
// channel consecutive pixels are consecutive floats, so every warp load is one coalesced 128-byte line), multiplies by
// the [K, C] weight table held in shared memory (broadcast reads) into K fp32 accumulators, adds the bias and emits the
// embedding and / or the distance logits -- the [B, K, h, w] embedding never makes a round trip through HBM and one
// launch replaces two.
//
// Why FFMA and not the tensor cores: the contraction is [pixels x C] . [C x K] with K <= 32: 2 K flops per 4 input
// bytes, i.e. 6.5 flop / B at K = 13 -- 42 TFLOP/s fp32 would saturate the measured 6.5 TB/s, the FFMA pipes deliver
// ~75, so the kernel is bound by reading the C-channel feature map either way (29.5 MB per 720x1280 image and scale:
// 4.7 us at peak), while TF32 / BF16 operands (10 / 7 mantissa bits) would break the 1e-5 bar on the logits.
#include "dml_common.cuh"

namespace dml {
namespace {

constexpr int CH_THREADS = 128;
constexpr int CH_CHUNK = 128;   // input channels per shared-memory weight stage

template <int K>
__global__ void __launch_bounds__(CH_THREADS) conv_head_kernel(const float* __restrict__ f, const float* __restrict__ wgt,
                                                               const float* __restrict__ bias, float diag_m, int C, long long HW,
                                                               float* __restrict__ emb, float* __restrict__ logits) {
  constexpr int KP = (K + 3) & ~3;
  __shared__ __align__(16) float s_w[CH_CHUNK][KP];   // channel-major, K contiguous: float4 broadcast reads
  const int b = blockIdx.y;
  const long long p = (long long)blockIdx.x * CH_THREADS + threadIdx.x;
  const bool active = p < HW;
  const float* src = f + ((long long)b * C) * HW + (active ? p : 0);
  float acc[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) acc[k] = 0.f;
  for (int c0 = 0; c0 < C; c0 += CH_CHUNK) {
    const int nc = min(CH_CHUNK, C - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < nc * KP; i += CH_THREADS) {
      const int c = i / KP, k = i - c * KP;
      s_w[c][k] = k < K ? wgt[(long long)k * C + c0 + c] : 0.f;
    }
    __syncthreads();
    // 8 independent channel loads in flight per thread
    int c = 0;
    for (; c + 8 <= nc; c += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = active ? __ldg(src + (long long)(c0 + c + j) * HW) : 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4* wr = reinterpret_cast<const float4*>(s_w[c + j]);
#pragma unroll
        for (int q = 0; q < KP / 4; ++q) {
          const float4 w4 = wr[q];
          acc[4 * q + 0] = fmaf(v[j], w4.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(v[j], w4.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(v[j], w4.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(v[j], w4.w, acc[4 * q + 3]);
        }
      }
    }
    for (; c < nc; ++c) {
      const float v = active ? __ldg(src + (long long)(c0 + c) * HW) : 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] = fmaf(v, s_w[c][k], acc[k]);
    }
  }
  if (!active) return;
  if (bias) {
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] += bias[k];
  }
  if (emb) {
    float* e = emb + ((long long)b * K) * HW + p;
#pragma unroll
    for (int k = 0; k < K; ++k) e[(long long)k * HW] = acc[k];
  }
  if (logits) {
    // d_k = sum_{d != k} x_d^2 + (x_k - m)^2 from positive terms only (no cancellation against ||x||^2): prefix / suffix
    // sums of the squares give the leave-one-out sum
    float sq[K], pre[K], suf[K];
#pragma unroll
    for (int k = 0; k < K; ++k) sq[k] = acc[k] * acc[k];
    pre[0] = 0.f;
#pragma unroll
    for (int k = 1; k < K; ++k) pre[k] = pre[k - 1] + sq[k - 1];
    suf[K - 1] = 0.f;
#pragma unroll
    for (int k = K - 2; k >= 0; --k) suf[k] = suf[k + 1] + sq[k + 1];
    float* z = logits + ((long long)b * K) * HW + p;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float t = acc[k] - diag_m;
      z[(long long)k * HW] = -fmaf(t, t, pre[k] + suf[k]);
    }
  }
}

template <int K>
int launch_conv_head(const float* f, const float* w, const float* bias, float m, int B, int C, long long HW, float* emb, float* logits,
                     cudaStream_t stream) {
  dim3 grid((unsigned)((HW + CH_THREADS - 1) / CH_THREADS), (unsigned)B);
  conv_head_kernel<K><<<grid, CH_THREADS, 0, stream>>>(f, w, bias, m, C, HW, emb, logits);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

}  // namespace
}  // namespace dml

using namespace dml;

extern "C" {
#pragma GCC visibility push(default)

int dml_conv1x1_head_forward(const float* features, const float* weight, const float* bias, float diag_m, int32_t B, int32_t C,
                             int32_t K, int32_t H, int32_t W, float* embedding, float* logits, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!features || !weight || (!embedding && !logits) || B < 0 || C < 1 || K < 1 || H < 0 || W < 0 || B > 65535) return DML_ERR_INVALID_ARG;
  if (K > DML_MAX_DIM) return DML_ERR_UNSUPPORTED_DIM;
  if (B == 0 || H == 0 || W == 0) return DML_OK;
  const long long hw = (long long)H * W;
#define DML_CH_CASE(Kv) case Kv: return launch_conv_head<Kv>(features, weight, bias, diag_m, B, C, hw, embedding, logits, stream);
  switch (K) {
    DML_CH_CASE(1) DML_CH_CASE(2) DML_CH_CASE(3) DML_CH_CASE(4) DML_CH_CASE(5) DML_CH_CASE(6) DML_CH_CASE(7) DML_CH_CASE(8)
    DML_CH_CASE(9) DML_CH_CASE(10) DML_CH_CASE(11) DML_CH_CASE(12) DML_CH_CASE(13) DML_CH_CASE(14) DML_CH_CASE(15) DML_CH_CASE(16)
    DML_CH_CASE(17) DML_CH_CASE(18) DML_CH_CASE(19) DML_CH_CASE(20) DML_CH_CASE(21) DML_CH_CASE(22) DML_CH_CASE(23) DML_CH_CASE(24)
    DML_CH_CASE(25) DML_CH_CASE(26) DML_CH_CASE(27) DML_CH_CASE(28) DML_CH_CASE(29) DML_CH_CASE(30) DML_CH_CASE(31) DML_CH_CASE(32)
    default: return DML_ERR_UNSUPPORTED_DIM;
  }
#undef DML_CH_CASE
}

#pragma GCC visibility pop
}  // extern "C"
