// Masked per-class feature sums (north-star kernel (c)): the few-shot novel-prototype mean
// (DeepLabV3Plus-Pytorch/test_embedding.py:413-419) and the per-class means of
// DeepLabV3Plus-Pytorch/utils/loss.py:65-67, as a warp-shuffle segmented reduction.
//
// A thread owns one pixel (its D channels in registers).  Segmentation labels are spatially
// coherent, so a warp usually sees 1-2 distinct classes: for every class present in the warp
// (found with __match_any_sync) the members' channel values are summed with xor-shuffles and the
// group leader adds the result to a WARP-PRIVATE float64 accumulator in shared memory (plain
// LDS/STS, no atomics).  Warps are then reduced in fixed order, blocks write double partials and a
// finalize kernel adds them in fixed order => bit-reproducible results.
#pragma once
#include "dml_common.cuh"

namespace dml {

constexpr int RED_THREADS = 256;
constexpr int RED_WARPS = RED_THREADS / 32;

struct ReduceArgs {
  const float* x;
  int nhwc;
  const uint8_t* l_u8;
  const long long* l_i64;
  int B, n_cls;
  long long HW;
  double* partials;  // [B][gridDim.x][n_cls][D+1]  (last column = count)
};

template <int D>
__global__ void __launch_bounds__(RED_THREADS) class_sums_kernel(const ReduceArgs a) {
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  double* s_acc = reinterpret_cast<double*>(smem_dyn);  // [RED_WARPS][n_cls][D+1]
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int row = D + 1;
  const int per_warp = a.n_cls * row;
  for (int i = tid; i < RED_WARPS * per_warp; i += RED_THREADS) s_acc[i] = 0.0;
  __syncthreads();
  double* my = s_acc + w * per_warp;

  const long long stride = (long long)gridDim.x * RED_THREADS;
  const long long iters = (a.HW + stride - 1) / stride;  // warp-uniform trip count
  for (long long it = 0; it < iters; ++it) {
    const long long p = (long long)blockIdx.x * RED_THREADS + tid + it * stride;
    const bool in = p < a.HW;
    float x[D];
    int lab = -1;
    if (in) {
      const long long pix = (long long)b * a.HW + p;
      const long long l = a.l_u8 ? (long long)a.l_u8[pix] : a.l_i64[pix];
      if (l >= 0 && l < a.n_cls) lab = (int)l;
      if (a.nhwc) {
        const float* src = a.x + pix * D;
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = src[d];
      } else {
        const float* src = a.x + ((long long)b * D) * a.HW + p;
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = __ldg(src + (long long)d * a.HW);
      }
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d) x[d] = 0.f;
    }
    // one round per distinct class present in the warp
    unsigned todo = __ballot_sync(0xffffffffu, lab >= 0);
    while (todo) {
      const int leader = __ffs(todo) - 1;
      const int cls = __shfl_sync(0xffffffffu, lab, leader);
      const bool mine = (lab == cls);
      const unsigned members = __ballot_sync(0xffffffffu, mine);
      float s[D];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        float v = mine ? x[d] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        s[d] = v;
      }
      if (lane == leader) {
        double* dst = my + cls * row;
#pragma unroll
        for (int d = 0; d < D; ++d) dst[d] += (double)s[d];
        dst[D] += (double)__popc(members);
      }
      todo &= ~members;
    }
  }
  __syncthreads();
  double* out = a.partials + ((size_t)b * gridDim.x + blockIdx.x) * per_warp;
  for (int i = tid; i < per_warp; i += RED_THREADS) {
    double t = 0.0;
#pragma unroll
    for (int ww = 0; ww < RED_WARPS; ++ww) t += s_acc[ww * per_warp + i];
    out[i] = t;
  }
}

template <int D>
int launch_class_sums(const ReduceArgs& a, int grid_x, cudaStream_t stream) {
  const size_t smem = (size_t)RED_WARPS * a.n_cls * (D + 1) * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(class_sums_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e);
  }
  class_sums_kernel<D><<<dim3((unsigned)grid_x, (unsigned)a.B), RED_THREADS, smem, stream>>>(a);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

#define DML_REDUCE_CASE(Dv) \
  case Dv:                  \
    return launch_class_sums<Dv>(a, gx, s);

int reduce_dispatch_1_8(int D, const ReduceArgs& a, int gx, cudaStream_t s);
int reduce_dispatch_9_16(int D, const ReduceArgs& a, int gx, cudaStream_t s);
int reduce_dispatch_17_24(int D, const ReduceArgs& a, int gx, cudaStream_t s);
int reduce_dispatch_25_32(int D, const ReduceArgs& a, int gx, cudaStream_t s);

}  // namespace dml
