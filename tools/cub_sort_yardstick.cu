// External yardstick for the sorting stage: cub::DeviceRadixSort / DeviceSegmentedRadixSort on
// the bench's key counts, on the same box.  NOT part of the product (the product's sort is csrc/ood_sort.cu; the
// default metric path no longer sorts the negatives at all) -- a number to put next to ours.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/cub_sort_yardstick tools/cub_sort_yardstick.cu
#include <cstdint>
#include <cstdio>
#include <cub/cub.cuh>

__global__ void fill(uint32_t* k, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t x = (uint32_t)i * 2654435761u;
    x ^= x >> 15; x *= 0x2c1b3c6du; x ^= x >> 12;
    // conf-like keys: float bits of a value in [0.25, 1), shifted left by one, label bit
    k[i] = (((0x3e800000u + (x % 0x01000000u)) - 0u) << 1) | ((x >> 7) % 97 == 0);
  }
}

template <typename F>
static float best_ms(F f, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  const size_t seg = 921600, n_small = 50 * seg, n_big = 1500 * seg;
  uint32_t *in, *out;
  cudaMalloc(&in, n_big * 4); cudaMalloc(&out, n_big * 4);
  int* offs; cudaMallocManaged(&offs, 51 * sizeof(int));
  for (int i = 0; i <= 50; ++i) offs[i] = (int)(i * seg);
  void* tmp = nullptr; size_t tb = 0, tb2 = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, tb, in, out, n_big, 0, 32);
  cub::DeviceSegmentedRadixSort::SortKeys(nullptr, tb2, in, out, (int)n_small, 50, offs, offs + 1, 0, 32);
  if (tb2 > tb) tb = tb2;
  cudaMalloc(&tmp, tb);
  auto refill = [&](size_t n) { fill<<<148 * 8, 256>>>(in, n); cudaDeviceSynchronize(); };
  refill(n_small);
  float t1 = best_ms([&] { cub::DeviceRadixSort::SortKeys(tmp, tb, in, out, n_small, 0, 32); }, 5);
  float t1b = best_ms([&] { cub::DeviceRadixSort::SortKeys(tmp, tb, in, out, n_small, 0, 31); }, 5);
  float t2 = best_ms([&] { cub::DeviceSegmentedRadixSort::SortKeys(tmp, tb, in, out, (int)n_small, 50, offs, offs + 1, 0, 32); }, 3);
  refill(n_big);
  float t3 = best_ms([&] { cub::DeviceRadixSort::SortKeys(tmp, tb, in, out, n_big, 0, 32); }, 3);
  printf("{\"cub_version\": %d, \"keys_small\": %zu, \"DeviceRadixSort_32bit_ms\": %.3f, \"DeviceRadixSort_31bit_ms\": %.3f, "
         "\"DeviceSegmentedRadixSort_50x921600_ms\": %.3f, \"keys_big\": %zu, \"DeviceRadixSort_big_32bit_ms\": %.3f, "
         "\"Gkeys_per_s_small\": %.1f, \"Gkeys_per_s_big\": %.1f, \"cuda_error\": \"%s\"}\n",
         CUB_VERSION, n_small, t1, t1b, t2, n_big, t3, n_small / t1 * 1e-6, n_big / t3 * 1e-6, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
