// Shared helpers for the dml_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/dml_b200.h"

#define DML_MAX_DIM 32

namespace dml {

extern thread_local int g_last_cuda_error;
extern unsigned long long g_kernel_launches;  // kernels enqueued by this library (bench bookkeeping; racy by design)

inline int cuda_fail(cudaError_t e) {
  g_last_cuda_error = (int)e;
  return DML_ERR_CUDA;
}

#define DML_CUDA_TRY(expr)                                   \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return ::dml::cuda_fail(_e);      \
  } while (0)

#define DML_LAUNCH_CHECK()                 \
  do {                                     \
    ++::dml::g_kernel_launches;            \
    DML_CUDA_TRY(cudaGetLastError());      \
  } while (0)

static inline int ceil_div_i(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- cache-hinted global access -------------------------------------------------------
// Streaming inputs are read exactly once: keep them out of L1 and mark them evict-first in
// L2 so that the small maps the next kernel re-reads (eds, keys) stay L2-resident.
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

template <int VEC> struct FVec;
template <> struct FVec<1> { float v[1]; };
template <> struct __align__(8) FVec<2> { float v[2]; };
template <> struct __align__(16) FVec<4> { float v[4]; };

template <int VEC>
__device__ __forceinline__ FVec<VEC> ld_stream(const float* p, uint64_t pol) {
  FVec<VEC> r;
  if constexpr (VEC == 4) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p), "l"(pol));
  } else if constexpr (VEC == 2) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;"
                 : "=f"(r.v[0]), "=f"(r.v[1]) : "l"(p), "l"(pol));
  } else {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;"
                 : "=f"(r.v[0]) : "l"(p), "l"(pol));
  }
  return r;
}

// streaming store (written once, not re-read by this pipeline)
template <int VEC>
__device__ __forceinline__ void st_stream(float* p, const FVec<VEC>& r) {
  if constexpr (VEC == 4) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]) : "memory");
  } else if constexpr (VEC == 2) {
    asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]) : "memory");
  } else {
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(r.v[0]) : "memory");
  }
}

// default-policy store (likely re-read soon by the next kernel: keep in L2)
template <int VEC>
__device__ __forceinline__ void st_keep(float* p, const FVec<VEC>& r) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(r.v[0], r.v[1]);
  } else {
    *p = r.v[0];
  }
}

// ---- packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2): two IEEE round-to-nearest operations per issue slot ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 r, float& lo, float& hi) {
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r));
}
__device__ __forceinline__ f32x2 fma2_rn(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2_rn(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2_rn(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ int warp_reduce_min_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_reduce_max_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_reduce_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_reduce_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace dml
