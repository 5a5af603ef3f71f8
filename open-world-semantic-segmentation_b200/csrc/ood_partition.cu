// Range partition of packed ranking keys into G contiguous buckets (G = number of ranks): the
// "partition" exchange mode of the multi-GPU pooled metric.  Instead of sorting locally, exchanging the
// sorted shards and sorting (merging) again, every rank scatters its UNSORTED keys into the G key ranges
// chosen from an all-gathered sample, ships range r to rank r (NCCL all-to-all) and sorts only what it
// receives: one partition pass (12 B / key) replaces one full radix sort (36 B / key).
//
//   bucket(key) = #{ j : key >= bounds[j] },  bounds[0..G-2] ascending (positive bit cleared, so the two
//   keys of one score value never straddle two buckets);  out = [bucket 0 | bucket 1 | ...].
// The order inside a bucket is unspecified (the receiver sorts); bucket sizes are exact.
#include "dml_common.cuh"

namespace dml {
namespace {

constexpr int PART_THREADS = 256;
constexpr int PART_KPT = 16;                       // keys per thread (4 x 16-byte loads)
constexpr int PART_TILE = PART_THREADS * PART_KPT;  // 4096 keys per CTA tile
constexpr int PART_WARPS = PART_THREADS / 32;
constexpr int PART_MAX_BUCKETS = DML_MAX_PARTITIONS;

// The kernels are compiled for NB = 2, 4, 8, 16 buckets (unused bounds are padded with 0xffffffff and the bucket id
// is clamped to nb-1, so the one key value that reaches a padded bound stays in the last bucket).  Everything is
// statically indexed:
//   ge[j] = #{keys >= bounds[j]}  (one compare + predicated add per key and bound)
//   bucket(key) = sum_j (key >= bounds[j]);  count[g] = ge[g-1] - ge[g]  with ge[-1] = n, ge[NB-1] = 0.

// pass 1: exact ">= bound" counts of the whole input (grid-stride, 16-byte loads, one warp reduction and NB-1
// global atomics per CTA at the end)
template <int NB>
__global__ void __launch_bounds__(PART_THREADS) part_count_kernel(const uint32_t* __restrict__ keys, long long n,
                                                                  const uint32_t* __restrict__ bounds, int nb,
                                                                  unsigned long long* __restrict__ ge_tot) {
  uint32_t bnd[NB - 1];
#pragma unroll
  for (int j = 0; j < NB - 1; ++j) bnd[j] = j < nb - 1 ? bounds[j] : 0xffffffffu;
  unsigned ge[NB - 1];
#pragma unroll
  for (int j = 0; j < NB - 1; ++j) ge[j] = 0u;
  const long long n4 = n >> 2;
  const uint4* k4 = reinterpret_cast<const uint4*>(keys);
  unsigned long long acc[NB - 1];
#pragma unroll
  for (int j = 0; j < NB - 1; ++j) acc[j] = 0ull;
  for (long long base = (long long)blockIdx.x * PART_THREADS * 4; base < n4; base += (long long)gridDim.x * PART_THREADS * 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = base + u * PART_THREADS + threadIdx.x;
      if (i < n4) {
        const uint4 q = k4[i];
#pragma unroll
        for (int j = 0; j < NB - 1; ++j)
          ge[j] += (q.x >= bnd[j] ? 1u : 0u) + (q.y >= bnd[j] ? 1u : 0u) + (q.z >= bnd[j] ? 1u : 0u) + (q.w >= bnd[j] ? 1u : 0u);
      }
    }
    // 16 keys per step: flush the 32-bit counters into 64-bit ones long before they can overflow
#pragma unroll
    for (int j = 0; j < NB - 1; ++j) { acc[j] += ge[j]; ge[j] = 0u; }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {   // tail keys
    const uint32_t q = keys[(n4 << 2) + threadIdx.x];
#pragma unroll
    for (int j = 0; j < NB - 1; ++j) acc[j] += q >= bnd[j] ? 1u : 0u;
  }
  __shared__ unsigned long long s_acc[NB];
  if (threadIdx.x < NB) s_acc[threadIdx.x] = 0ull;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < NB - 1; ++j) {
    const unsigned long long v = warp_reduce_sum_u64(acc[j]);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_acc[j], v);
  }
  __syncthreads();
  if (threadIdx.x < nb - 1 && s_acc[threadIdx.x]) atomicAdd(&ge_tot[threadIdx.x], s_acc[threadIdx.x]);
}

// 16-bit lanes packed four to a 64-bit word (tile-local counts / offsets never exceed 4096)
template <int NB> struct Packed { unsigned long long w[(NB + 3) / 4]; };

template <int NB>
__device__ __forceinline__ Packed<NB> packed_shfl_up(const Packed<NB>& a, int delta) {
  Packed<NB> r;
#pragma unroll
  for (int i = 0; i < (NB + 3) / 4; ++i) r.w[i] = __shfl_up_sync(0xffffffffu, a.w[i], delta);
  return r;
}

// pass 2: every CTA step handles one 4096-key tile.  Thread-local bucket counts (from the static ">= bound"
// counters) are scanned over the CTA as packed 16-bit lanes, every thread drops its keys into a bucket-contiguous
// stage in shared memory, and each bucket run goes to a slice of the bucket's output area reserved with ONE global
// atomic per bucket and tile (coalesced stores; the order inside a bucket is irrelevant -- the receiver sorts).
template <int NB>
__global__ void __launch_bounds__(PART_THREADS) part_scatter_kernel(const uint32_t* __restrict__ keys, long long n,
                                                                    const uint32_t* __restrict__ bounds, int nb,
                                                                    const unsigned long long* __restrict__ ge_tot,
                                                                    unsigned long long* __restrict__ cursors,
                                                                    uint32_t* __restrict__ out, long long* __restrict__ counts) {
  constexpr int NW = (NB + 3) / 4;
  __shared__ unsigned long long s_wtot[PART_WARPS][NW];      // per-warp packed bucket totals
  __shared__ unsigned s_loff[NB + 1];                         // tile-local start of every bucket run
  __shared__ unsigned long long s_gpos[NB];                   // global start of this tile's slice per bucket
  __shared__ unsigned long long s_base[NB];                   // global start of every bucket
  __shared__ __align__(16) uint32_t s_keys[PART_TILE];
  uint32_t bnd[NB - 1];
#pragma unroll
  for (int j = 0; j < NB - 1; ++j) bnd[j] = j < nb - 1 ? bounds[j] : 0xffffffffu;
  if (threadIdx.x < NB) {
    // keys in buckets >= g number ge[g-1] (ge[-1] = n): bucket g starts at n - that
    const int g = threadIdx.x;
    const unsigned long long ge_prev = g == 0 ? (unsigned long long)n : (g - 1 < nb - 1 ? ge_tot[g - 1] : 0ull);
    const unsigned long long ge_this = g < nb - 1 ? ge_tot[g] : 0ull;
    s_base[g] = (unsigned long long)n - ge_prev;
    if (blockIdx.x == 0 && g < nb) counts[g] = (long long)(ge_prev - ge_this);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long n_tiles = (n + PART_TILE - 1) / PART_TILE;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long base = tile * PART_TILE;
    uint32_t key[PART_KPT];
    bool full = base + PART_TILE <= n;
    if (full) {
      const uint4* k4 = reinterpret_cast<const uint4*>(keys + base);
#pragma unroll
      for (int u = 0; u < PART_KPT / 4; ++u) {
        const uint4 q = k4[u * PART_THREADS + threadIdx.x];
        key[4 * u] = q.x; key[4 * u + 1] = q.y; key[4 * u + 2] = q.z; key[4 * u + 3] = q.w;
      }
    } else {
#pragma unroll
      for (int u = 0; u < PART_KPT / 4; ++u)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const long long idx = base + (long long)(u * PART_THREADS + threadIdx.x) * 4 + c;
          key[4 * u + c] = idx < n ? keys[idx] : 0xffffffffu;   // padding lands past every bound: dropped below
        }
    }
    // thread-local: bucket of every key and packed bucket counts
    unsigned long long bkt = 0ull;    // 16 x 4-bit bucket ids
    unsigned vmask = 0u;              // valid keys
    Packed<NB> cnt;
#pragma unroll
    for (int i = 0; i < NW; ++i) cnt.w[i] = 0ull;
#pragma unroll
    for (int i = 0; i < PART_KPT; ++i) {
      unsigned b = 0;
#pragma unroll
      for (int j = 0; j < NB - 1; ++j) b += key[i] >= bnd[j] ? 1u : 0u;
      b = min(b, (unsigned)(nb - 1));
      const bool valid = full || (base + (long long)((i >> 2) * PART_THREADS + threadIdx.x) * 4 + (i & 3) < n);
      bkt |= (unsigned long long)b << (4 * i);
      vmask |= valid ? (1u << i) : 0u;
      if (valid) {
        if constexpr (NW == 1) cnt.w[0] += 1ull << (16 * b);
        else {
#pragma unroll
          for (int wd = 0; wd < NW; ++wd) cnt.w[wd] += ((b >> 2) == (unsigned)wd) ? (1ull << (16 * (b & 3))) : 0ull;
        }
      }
    }
    // inclusive warp scan of the packed counts (lanes never carry into each other: totals <= 4096 < 65536)
    Packed<NB> inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const Packed<NB> up = packed_shfl_up<NB>(inc, d);
      if (lane >= d) {
#pragma unroll
        for (int i = 0; i < NW; ++i) inc.w[i] += up.w[i];
      }
    }
    __syncthreads();   // previous tile fully copied out: shared tables and stage are free again
    if (lane == 31) {
#pragma unroll
      for (int i = 0; i < NW; ++i) s_wtot[warp][i] = inc.w[i];
    }
    __syncthreads();
    // offsets of this thread per bucket: bucket start in the tile + totals of the warps before + lanes before
    Packed<NB> off;
#pragma unroll
    for (int i = 0; i < NW; ++i) off.w[i] = inc.w[i] - cnt.w[i];
    Packed<NB> tot;
#pragma unroll
    for (int i = 0; i < NW; ++i) tot.w[i] = 0ull;
#pragma unroll
    for (int w = 0; w < PART_WARPS; ++w) {
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        const unsigned long long t = s_wtot[w][i];
        if (w < warp) off.w[i] += t;
        tot.w[i] += t;
      }
    }
    // exclusive prefix over the buckets (every thread computes it redundantly: NB <= 16 adds)
    unsigned loff[NB + 1];
    loff[0] = 0;
#pragma unroll
    for (int g = 0; g < NB; ++g) loff[g + 1] = loff[g] + (unsigned)((tot.w[g >> 2] >> (16 * (g & 3))) & 0xffffull);
#pragma unroll
    for (int g = 0; g < NB; ++g) off.w[g >> 2] += (unsigned long long)loff[g] << (16 * (g & 3));
    if (threadIdx.x < NB) {
      const int g = threadIdx.x;
      // (static indexing: select this thread's bucket)
      unsigned start = 0, size = 0;
#pragma unroll
      for (int q = 0; q < NB; ++q) if (q == g) { start = loff[q]; size = loff[q + 1] - loff[q]; }
      s_loff[g] = start;
      if (g == NB - 1) s_loff[NB] = start + size;
      s_gpos[g] = size ? s_base[g] + atomicAdd(&cursors[g], (unsigned long long)size) : 0ull;
    }
#pragma unroll
    for (int i = 0; i < PART_KPT; ++i) {
      if (vmask & (1u << i)) {
        const unsigned b = (unsigned)(bkt >> (4 * i)) & 15u;
        unsigned pos;
        if constexpr (NW == 1) {
          pos = (unsigned)((off.w[0] >> (16 * b)) & 0xffffull);
          off.w[0] += 1ull << (16 * b);
        } else {
          unsigned long long sel = 0ull;
#pragma unroll
          for (int wd = 0; wd < NW; ++wd) {
            const bool mine = (b >> 2) == (unsigned)wd;
            sel = mine ? off.w[wd] : sel;
            off.w[wd] += mine ? (1ull << (16 * (b & 3))) : 0ull;
          }
          pos = (unsigned)((sel >> (16 * (b & 3))) & 0xffffull);
        }
        s_keys[pos] = key[i];
      }
    }
    __syncthreads();
    const unsigned tile_n = s_loff[NB];
    int g = 0;
    for (unsigned i = threadIdx.x; i < tile_n; i += PART_THREADS) {
      while (i >= s_loff[g + 1]) ++g;
      out[s_gpos[g] + (i - s_loff[g])] = s_keys[i];
    }
  }
}

template <int NB>
int launch_partition(const uint32_t* keys, long long n, const uint32_t* bounds, int nb, uint32_t* out, long long* counts,
                     unsigned long long* ge_tot, unsigned long long* cursors, cudaStream_t stream) {
  const long long n_tiles = (n + PART_TILE - 1) / PART_TILE;
  const int grid = (int)(n_tiles < 148 * 8 ? n_tiles : 148 * 8);   // persistent: up to 8 CTAs per SM
  part_count_kernel<NB><<<grid, PART_THREADS, 0, stream>>>(keys, n, bounds, nb, ge_tot);
  DML_LAUNCH_CHECK();
  part_scatter_kernel<NB><<<grid, PART_THREADS, 0, stream>>>(keys, n, bounds, nb, ge_tot, cursors, out, counts);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

}  // namespace
}  // namespace dml

using namespace dml;

extern "C" {
#pragma GCC visibility push(default)

size_t dml_ood_partition_workspace_bytes(int32_t n_buckets) {
  (void)n_buckets;
  return 2 * DML_MAX_PARTITIONS * sizeof(unsigned long long);
}

int dml_ood_partition(const uint32_t* keys, int64_t n, const uint32_t* bounds, int32_t n_buckets, uint32_t* out,
                      long long* counts, void* workspace, size_t workspace_bytes, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || n_buckets < 1 || n_buckets > PART_MAX_BUCKETS || !counts || !workspace) return DML_ERR_INVALID_ARG;
  if (n > 0 && (!keys || !out)) return DML_ERR_INVALID_ARG;
  if (n_buckets > 1 && !bounds) return DML_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(keys) & 15) != 0) return DML_ERR_INVALID_ARG;   // 16-byte loads
  if (workspace_bytes < dml_ood_partition_workspace_bytes(n_buckets)) return DML_ERR_WORKSPACE;
  unsigned long long* ge_tot = reinterpret_cast<unsigned long long*>(workspace);
  unsigned long long* cursors = ge_tot + PART_MAX_BUCKETS;
  DML_CUDA_TRY(cudaMemsetAsync(workspace, 0, 2 * PART_MAX_BUCKETS * sizeof(unsigned long long), stream));
  DML_CUDA_TRY(cudaMemsetAsync(counts, 0, (size_t)n_buckets * sizeof(long long), stream));
  if (n == 0) return DML_OK;
  if (n_buckets <= 2) return launch_partition<2>(keys, n, bounds, n_buckets, out, counts, ge_tot, cursors, stream);
  if (n_buckets <= 4) return launch_partition<4>(keys, n, bounds, n_buckets, out, counts, ge_tot, cursors, stream);
  if (n_buckets <= 8) return launch_partition<8>(keys, n, bounds, n_buckets, out, counts, ge_tot, cursors, stream);
  return launch_partition<16>(keys, n, bounds, n_buckets, out, counts, ge_tot, cursors, stream);
}

#pragma GCC visibility pop
}  // extern "C"
