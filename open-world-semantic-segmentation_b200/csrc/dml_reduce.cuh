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

// ---- lane-owns-class variant (n_cls <= 32) ---------------------------------------------------------------------
// Lane c of every warp owns the running float64 sums of class c (a private shared-memory row that only this lane ever
// touches: no atomics, no conflicts between lanes; kept out of the register file for occupancy).  A thread owns VEC consecutive pixels (16-byte loads per channel); when its VEC labels agree (the
// usual case: label maps are piecewise constant) they are pre-added and handled as one weighted pixel.  Per slot the
// warp picks the cheaper of two exchanges:
//   * few classes, many members (coherent labels): for every distinct class a xor-shuffle reduction (5 D shuffles) whose
//     result lane `class` keeps;
//   * many classes, few members each (incoherent labels -- the old kernel's 20x worst case): every lane fetches the
//     values of the members of ITS class one by one with indexed shuffles (max-members x D shuffles; the member masks
//     of all 32 classes come from 5 ballots on the label bits).
// Fixed evaluation order => bit-reproducible.  Replaces DeepLabV3Plus-Pytorch/test_embedding.py:413-419,
// utils/loss.py:65-67 like class_sums_kernel.
template <int D, int VEC>
__global__ void __launch_bounds__(RED_THREADS, 2) class_sums_lane_kernel(const ReduceArgs a) {
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  double* s_out = reinterpret_cast<double*>(smem_dyn);   // [RED_WARPS][32][D + 1]
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  constexpr int ROW = D + 1;
  double* acc = s_out + ((size_t)w * 32 + lane) * ROW;   // this lane's class row
#pragma unroll
  for (int d = 0; d < D; ++d) acc[d] = 0.0;
  unsigned long long cnt = 0ull;
  unsigned lt;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));

  // `weight`: pixels this lane's value stands for (per lane: ignored pixels of a thread are left out of its pre-sum)
  auto slot = [&](int lab, const float (&val)[D], unsigned weight) {
    const bool valid = lab >= 0;
    const unsigned vm = __ballot_sync(0xffffffffu, valid);
    if (vm == 0u) return;
    unsigned m_c = vm, own = vm;     // members of class `lane` / of my own label
#pragma unroll
    for (int bit = 0; bit < 5; ++bit) {
      const unsigned bb = __ballot_sync(0xffffffffu, valid && ((lab >> bit) & 1));
      m_c &= ((lane >> bit) & 1) ? bb : ~bb;
      own &= ((lab >> bit) & 1) ? bb : ~bb;
    }
    if (lane >= a.n_cls) m_c = 0u;
    const bool leader = valid && (own & lt) == 0u;
    unsigned todo = __ballot_sync(0xffffffffu, leader);
    const int nd = __popc(todo);
    const int mm = (int)__reduce_max_sync(0xffffffffu, (unsigned)__popc(m_c));
    if (5 * nd <= mm) {
      while (todo) {
        const int ld = __ffs(todo) - 1;
        todo &= todo - 1u;
        const int cls = __shfl_sync(0xffffffffu, lab, ld);
        const bool mine = valid && lab == cls;
        const unsigned members = __ballot_sync(0xffffffffu, mine);
#pragma unroll
        for (int d = 0; d < D; ++d) {
          float v = mine ? val[d] : 0.f;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == cls) acc[d] += (double)v;
        }
        const unsigned wsum = __reduce_add_sync(0xffffffffu, mine ? weight : 0u);
        if (lane == cls) cnt += wsum;
      }
    } else {
      float part[D];
#pragma unroll
      for (int d = 0; d < D; ++d) part[d] = 0.f;
      unsigned mask = m_c, wpart = 0u;
      for (int i = 0; i < mm; ++i) {
        const bool have = mask != 0u;
        const int src = have ? __ffs(mask) - 1 : lane;
        mask &= mask - 1u;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float v = __shfl_sync(0xffffffffu, val[d], src);
          part[d] += have ? v : 0.f;
        }
        const unsigned wv = __shfl_sync(0xffffffffu, weight, src);
        wpart += have ? wv : 0u;
      }
#pragma unroll
      for (int d = 0; d < D; ++d) acc[d] += (double)part[d];
      cnt += wpart;
    }
  };

  const long long nvec = (a.HW + VEC - 1) / VEC;
  const long long stride = (long long)gridDim.x * RED_THREADS;
  const long long iters = (nvec + stride - 1) / stride;   // warp-uniform trip count
  for (long long it = 0; it < iters; ++it) {
    const long long q = (long long)blockIdx.x * RED_THREADS + tid + it * stride;
    const long long p = q * VEC;
    float x[D][VEC];
    int lab[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) lab[v] = -1;
    if (p < a.HW) {   // (VEC > 1 is only launched when HW % VEC == 0: the VEC pixels of a thread are all inside)
      const long long pix = (long long)b * a.HW + p;
      if constexpr (VEC == 4) {
        if (a.l_u8) {
          const uchar4 t = *reinterpret_cast<const uchar4*>(a.l_u8 + pix);
          lab[0] = t.x; lab[1] = t.y; lab[2] = t.z; lab[3] = t.w;
        } else {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const long long l = a.l_i64[pix + v];
            lab[v] = (l >= 0 && l < a.n_cls) ? (int)l : -1;
          }
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) if (lab[v] >= a.n_cls) lab[v] = -1;
        const float* src = a.x + ((long long)b * D) * a.HW + p;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(src + (long long)d * a.HW));
          x[d][0] = t.x; x[d][1] = t.y; x[d][2] = t.z; x[d][3] = t.w;
        }
      } else {
        const long long l = a.l_u8 ? (long long)a.l_u8[pix] : a.l_i64[pix];
        lab[0] = (l >= 0 && l < a.n_cls) ? (int)l : -1;
        if (a.nhwc) {
          const float* src = a.x + pix * D;
#pragma unroll
          for (int d = 0; d < D; ++d) x[d][0] = src[d];
        } else {
          const float* src = a.x + ((long long)b * D) * a.HW + p;
#pragma unroll
          for (int d = 0; d < D; ++d) x[d][0] = __ldg(src + (long long)d * a.HW);
        }
      }
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d)
#pragma unroll
        for (int v = 0; v < VEC; ++v) x[d][v] = 0.f;
    }
    // a thread whose VALID pixels all carry one class is pre-added (ignored pixels left out) and travels as one
    // weighted pixel; only threads straddling a class boundary force the warp onto the per-pixel rounds
    int labu = -1;
    unsigned nvalid = 0;
    bool uniform = true;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      if (lab[v] >= 0) {
        if (labu < 0) labu = lab[v];
        uniform = uniform && lab[v] == labu;
        ++nvalid;
      }
    }
    if (VEC == 1 || __all_sync(0xffffffffu, uniform)) {
      float s[D];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        float t = 0.f;
#pragma unroll
        for (int v = 0; v < VEC; ++v) t += (lab[v] >= 0) ? x[d][v] : 0.f;
        s[d] = t;
      }
      slot(labu, s, nvalid);
    } else {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float s[D];
#pragma unroll
        for (int d = 0; d < D; ++d) s[d] = x[d][v];
        slot(lab[v], s, 1u);
      }
    }
  }
  // fixed-order reduction over the warps
  acc[D] = (double)cnt;
  __syncthreads();
  const int per_block = a.n_cls * ROW;
  double* out = a.partials + ((size_t)b * gridDim.x + blockIdx.x) * per_block;
  for (int i = tid; i < per_block; i += RED_THREADS) {
    const int c = i / ROW, d = i - c * ROW;
    double t = 0.0;
#pragma unroll
    for (int ww = 0; ww < RED_WARPS; ++ww) t += s_out[((size_t)ww * 32 + c) * ROW + d];
    out[i] = t;
  }
}

template <int D>
int launch_class_sums(const ReduceArgs& a, int grid_x, cudaStream_t stream) {
  if (a.n_cls <= 32) {
    const size_t smem = (size_t)RED_WARPS * 32 * (D + 1) * sizeof(double);
    auto al = [](const void* q, size_t n) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % n) == 0; };
    const bool vec4 = !a.nhwc && a.HW % 4 == 0 && al(a.x, 16) && al(a.l_u8, 4) && D <= 24;
    cudaError_t e = vec4 ? cudaFuncSetAttribute(class_sums_lane_kernel<D, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                         : cudaFuncSetAttribute(class_sums_lane_kernel<D, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e);
    const dim3 grid((unsigned)grid_x, (unsigned)a.B);
    if (vec4) class_sums_lane_kernel<D, 4><<<grid, RED_THREADS, smem, stream>>>(a);
    else class_sums_lane_kernel<D, 1><<<grid, RED_THREADS, smem, stream>>>(a);
    DML_LAUNCH_CHECK();
    return DML_OK;
  }
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
