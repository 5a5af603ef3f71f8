// Micro-benchmarks behind the design of the minority-rank metric path:
// throughput of the primitives a "rank every negative against the sorted positives" kernel is built from, on the
// box it runs on.  Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/microbench_rank tools/microbench_rank.cu
//   (1) shared-memory atomics WITHOUT return value (ATOMS.POPC.INC / RED-like), spread addresses
//   (2) shared-memory atomics WITH return value (rank inside a tile: what a non-stable scatter needs)
//   (3) global RED, spread over a table of 2^22 / 2^25 counters, and on one hot address
//   (4) binary search in a shared-memory table (14 dependent LDS at random addresses) vs LUT + 3 probes
//   (5) match.any.sync vs 8 ballots
// Prints one JSON object per line: {"bench": ..., "keys": N, "us": t, "Gkeys_per_s": r}
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

__global__ void fill_kernel(uint32_t* k, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) k[i] = hash32((uint32_t)i);
}

template <int NC>
__global__ void __launch_bounds__(1024) smem_atomic_noret(const uint32_t* __restrict__ keys, size_t n, uint32_t* out) {
  extern __shared__ uint32_t s_c[];
  for (int i = threadIdx.x; i < NC; i += blockDim.x) s_c[i] = 0;
  __syncthreads();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    atomicAdd(&s_c[keys[i] & (NC - 1)], 1u);
  __syncthreads();
  uint32_t a = 0;
  for (int i = threadIdx.x; i < NC; i += blockDim.x) a += s_c[i];
  if (a == 0xffffffffu) out[0] = a;
}

template <int NC>
__global__ void __launch_bounds__(1024) smem_atomic_ret(const uint32_t* __restrict__ keys, size_t n, uint32_t* out) {
  extern __shared__ uint32_t s_c[];
  for (int i = threadIdx.x; i < NC; i += blockDim.x) s_c[i] = 0;
  __syncthreads();
  uint32_t acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc += atomicAdd(&s_c[keys[i] & (NC - 1)], 1u);
  if (acc == 0xfffffff1u) out[0] = acc;
}

__global__ void __launch_bounds__(256) global_red(const uint32_t* __restrict__ keys, size_t n, uint32_t* table, uint32_t mask) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    atomicAdd(&table[keys[i] & mask], 1u);
}

// binary search over NS sorted entries in shared memory (lower_bound), NS a power of two
template <int NS>
__global__ void __launch_bounds__(1024) smem_bsearch(const uint32_t* __restrict__ keys, size_t n, uint32_t* out) {
  extern __shared__ uint32_t s_t[];
  for (int i = threadIdx.x; i < NS; i += blockDim.x) s_t[i] = (uint32_t)(((unsigned long long)i << 32) / NS);
  __syncthreads();
  uint32_t acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t k = keys[i];
    int lo = 0;
#pragma unroll
    for (int step = NS / 2; step > 0; step >>= 1) lo = (s_t[lo + step - 1] < k) ? lo + step : lo;
    acc += lo;
  }
  if (acc == 0xfffffff1u) out[0] = acc;
}

// LUT (NL entries) -> start of a short range, then 3 probes
template <int NS, int NL>
__global__ void __launch_bounds__(1024) smem_lut_search(const uint32_t* __restrict__ keys, size_t n, uint32_t* out) {
  extern __shared__ uint32_t s_t[];
  uint32_t* s_lut = s_t + NS;
  for (int i = threadIdx.x; i < NS; i += blockDim.x) s_t[i] = (uint32_t)(((unsigned long long)i << 32) / NS);
  for (int i = threadIdx.x; i < NL; i += blockDim.x) s_lut[i] = (uint32_t)((unsigned long long)i * NS / NL);
  __syncthreads();
  uint32_t acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t k = keys[i];
    int lo = (int)s_lut[k >> (32 - __builtin_ctz(NL))];
#pragma unroll
    for (int p = 0; p < 3; ++p) lo += (lo < NS && s_t[lo] < k) ? 1 : 0;
    acc += lo;
  }
  if (acc == 0xfffffff1u) out[0] = acc;
}

__global__ void __launch_bounds__(256) match_any_kernel(const uint32_t* __restrict__ keys, size_t n, uint32_t* out) {
  uint32_t acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t nround = (n / stride) * stride;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nround; i += stride)
    acc += __popc(__match_any_sync(0xffffffffu, keys[i] & 255u));
  if (acc == 0xfffffff1u) out[0] = acc;
}

__global__ void __launch_bounds__(256) ballot8_kernel(const uint32_t* __restrict__ keys, size_t n, uint32_t* out) {
  uint32_t acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t nround = (n / stride) * stride;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nround; i += stride) {
    const uint32_t d = keys[i] & 255u;
    unsigned peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const bool bit = (d >> b) & 1u;
      const unsigned m = __ballot_sync(0xffffffffu, bit);
      peers &= bit ? m : ~m;
    }
    acc += __popc(peers);
  }
  if (acc == 0xfffffff1u) out[0] = acc;
}

template <typename F>
static float time_it(F f, int reps = 5) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best * 1e3f;
}

static void report(const char* name, size_t n, float us) {
  printf("{\"bench\": \"%s\", \"keys\": %zu, \"us\": %.1f, \"Gkeys_per_s\": %.1f}\n", name, n, us, n / us * 1e-3);
}

int main() {
  const size_t n = 46080000;   // one 50-image chunk of 720x1280 pixels
  uint32_t *keys, *out, *table;
  CK(cudaMalloc(&keys, n * 4));
  CK(cudaMalloc(&out, 256));
  CK(cudaMalloc(&table, (size_t)(1u << 25) * 4));
  CK(cudaMemset(table, 0, (size_t)(1u << 25) * 4));
  fill_kernel<<<148 * 8, 256>>>(keys, n);
  CK(cudaDeviceSynchronize());
  const int G1 = 148, G2 = 296;
#define SMEM_OPT(k, bytes) CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))
  SMEM_OPT(smem_atomic_noret<2048>, 2048 * 4);
  SMEM_OPT(smem_atomic_noret<32768>, 32768 * 4);
  SMEM_OPT(smem_atomic_ret<2048>, 2048 * 4);
  SMEM_OPT(smem_atomic_ret<32768>, 32768 * 4);
  SMEM_OPT(smem_bsearch<16384>, 16384 * 4);
  SMEM_OPT((smem_lut_search<16384, 4096>), (16384 + 4096) * 4);
  report("smem_atomic_noret_2048ctr_1024thr", n, time_it([&] { smem_atomic_noret<2048><<<G2, 1024, 2048 * 4>>>(keys, n, out); }));
  report("smem_atomic_noret_32768ctr_1024thr", n, time_it([&] { smem_atomic_noret<32768><<<G1, 1024, 32768 * 4>>>(keys, n, out); }));
  report("smem_atomic_ret_2048ctr_1024thr", n, time_it([&] { smem_atomic_ret<2048><<<G2, 1024, 2048 * 4>>>(keys, n, out); }));
  report("smem_atomic_ret_32768ctr_1024thr", n, time_it([&] { smem_atomic_ret<32768><<<G1, 1024, 32768 * 4>>>(keys, n, out); }));
  report("global_red_4M_counters", n, time_it([&] { global_red<<<148 * 8, 256>>>(keys, n, table, (1u << 22) - 1); }));
  report("global_red_32M_counters", n, time_it([&] { global_red<<<148 * 8, 256>>>(keys, n, table, (1u << 25) - 1); }));
  report("global_red_64K_counters", n, time_it([&] { global_red<<<148 * 8, 256>>>(keys, n, table, (1u << 16) - 1); }));
  report("global_red_one_address", n / 16, time_it([&] { global_red<<<148 * 8, 256>>>(keys, n / 16, table, 0u); }));
  report("smem_bsearch_16384_14steps", n, time_it([&] { smem_bsearch<16384><<<G1, 1024, 16384 * 4>>>(keys, n, out); }));
  report("smem_lut4096_plus_3probes", n, time_it([&] { smem_lut_search<16384, 4096><<<G1, 1024, (16384 + 4096) * 4>>>(keys, n, out); }));
  report("match_any_8bit", n, time_it([&] { match_any_kernel<<<148 * 8, 256>>>(keys, n, out); }));
  report("ballot8_8bit", n, time_it([&] { ballot8_kernel<<<148 * 8, 256>>>(keys, n, out); }));
  return 0;
}
