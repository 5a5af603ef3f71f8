// Pooled (full-set) exact AUROC / AUPR / FPR@recall without sorting the negatives: the minority-rank idea of
// ood_rank.cu for ONE long ranking of up to 2^32 - 1 (score, label) pairs whose positives no longer fit a CTA.
//
//   dml_ood_pos_compact     packed keys -> score keys of the positives (any order)                   4 B / pair read
//   dml_ood_sort            (ood_sort.cu) radix sort of the positives only (~1 % of the pairs)
//   dml_ood_unique_counts   sorted positives -> distinct scores S[g] + multiplicities pc[g]
//   dml_ood_bucket_rank     negatives -> counters bt[g] (strictly between S[g-1] and S[g]) / eq[g] (== S[g]):
//       the G groups are cut into B = ceil(G / 12288) buckets of consecutive groups; one counting pass + one
//       NON-STABLE scatter pass (the order inside a bucket is irrelevant: ranks come from shared-memory atomics with
//       return value, no digit matching, no look-back) group the negatives by bucket (8 B / pair), then fixed-size units
//       of every bucket are ranked against the bucket's positives in shared memory exactly like rank_kernel does per
//       image (4 B / pair).  16 B / pair and ~1/3 of the instructions of the 4-pass LSD sort + scan (40 B / pair).
//   dml_ood_pooled_scan     scan over the G groups -> dml_ood_result (same arithmetic as rank_scan_kernel)
// The counters are plain sums over the negatives, so a multi-GPU evaluation needs no exchange of the negatives at
// all: every rank ranks ITS negatives against the all-gathered positives and the counters are all-reduced
// (distributed.pooled_measures(mode="rank")).  Reference semantics: anomaly/anom_utils.py:25-78 over all pixels.
#include <cstdlib>
#include "ood_rank.cuh"

namespace dml {
namespace {

constexpr int PR_PASS_CAP = 12288;        // positive groups per bucket (shared memory: 48 KB keys + 96 KB counters + 16 KB LUT)
constexpr int PR_MAX_BUCKETS = 4096;
constexpr long long PR_FLAT_BELOW = 1200;  // mean chunk length below which unit_rank walks its chunks as one flat sequence (A/B on
                                          // the B200, profiles/r2u_unit_rank_ab.txt: 787 -> 2.12 vs 2.36 ms, 1570 -> 3.82 vs 3.75, 6300 -> 13.9 vs 12.9)
constexpr int PR_MAX_SLICES = 148 * 2;    // slices of the key array: one per CTA of the counting / scatter passes, 2 CTAs per SM
constexpr long long PR_UNIT = 1ll << 20;  // keys per ranking unit, at most (and at least 1 / (8 * 148) of the keys, >= 64 K:
                                          // a 1/8 shard cut into 1 M-key units left 165 long units for 148 SMs -- two waves)
constexpr int UNIQ_TILE = 4096;
constexpr int SCAN_GROUPS_PER_BLOCK = 8192;

struct Unit { uint32_t bucket, s0, s1, pad; };   // bucket, slices [s0, s1)

// ---- positives of a key array --------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pos_compact_kernel(const uint32_t* __restrict__ keys, long long n, uint32_t* __restrict__ out,
                                                          long long capacity, unsigned long long* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const long long nvec = (n + 3) / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long q_end = ((nvec + stride - 1) / stride) * stride;
  const bool aligned = (reinterpret_cast<uintptr_t>(keys) & 15) == 0;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < q_end; q += stride) {
    uint32_t k[4] = {0u, 0u, 0u, 0u};
    const long long i0 = q * 4;
    if (i0 + 4 <= n && aligned) {
      const uint4 t = *reinterpret_cast<const uint4*>(keys + i0);
      k[0] = t.x; k[1] = t.y; k[2] = t.z; k[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) k[j] = (i0 + j < n) ? keys[i0 + j] : 0u;
    }
    const int c = (int)((k[0] & 1u) + (k[1] & 1u) + (k[2] & 1u) + (k[3] & 1u));
    if (__ballot_sync(0xffffffffu, c > 0) == 0u) continue;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int m = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += m;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long wbase = 0;
    if (lane == 31) wbase = atomicAdd(count, (unsigned long long)total);
    wbase = __shfl_sync(0xffffffffu, wbase, 31);
    long long dst = (long long)wbase + (incl - c);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (k[j] & 1u) {
        if (dst < capacity) out[dst] = k[j] >> 1;
        ++dst;
      }
    }
  }
}

// ---- distinct values of a sorted array ---------------------------------------------------------------------
__global__ void __launch_bounds__(1024) uniq_count_kernel(const uint32_t* __restrict__ a, long long n, uint32_t* __restrict__ bc) {
  __shared__ uint32_t s_w[33];
  const long long base = (long long)blockIdx.x * UNIQ_TILE + (long long)threadIdx.x * 4;
  uint32_t heads = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long i = base + j;
    if (i < n) heads += (i == 0 || a[i] != a[i - 1]) ? 1u : 0u;
  }
  heads = __reduce_add_sync(0xffffffffu, heads);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = heads;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int i = 0; i < 32; ++i) t += s_w[i];
    bc[blockIdx.x] = t;
  }
}

// exclusive scan of m 32-bit counts in place (single CTA), total -> *total_out (u64)
__global__ void __launch_bounds__(1024) scan_u32_kernel(uint32_t* __restrict__ v, long long m, unsigned long long* __restrict__ total_out) {
  __shared__ unsigned long long s_w[32];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const long long per = (m + 1023) / 1024;
  const long long b0 = min((long long)tid * per, m), b1 = min(b0 + per, m);
  unsigned long long sum = 0;
  for (long long i = b0; i < b1; ++i) sum += v[i];
  unsigned long long incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_w[w] = incl;
  __syncthreads();
  unsigned long long run = incl - sum, tot = 0;
  for (int i = 0; i < 32; ++i) { if (i < w) run += s_w[i]; tot += s_w[i]; }
  for (long long i = b0; i < b1; ++i) { const uint32_t c = v[i]; v[i] = (uint32_t)run; run += c; }
  if (tid == 0 && total_out) *total_out = tot;
}

__global__ void __launch_bounds__(1024) uniq_write_kernel(const uint32_t* __restrict__ a, long long n, const uint32_t* __restrict__ bc,
                                                          uint32_t* __restrict__ S, uint32_t* __restrict__ start) {
  __shared__ uint32_t s_w[33];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const long long base = (long long)blockIdx.x * UNIQ_TILE + (long long)tid * 4;
  uint32_t k[4];
  bool head[4];
  uint32_t c = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long i = base + j;
    k[j] = i < n ? a[i] : 0u;
    head[j] = i < n && (i == 0 || k[j] != a[i - 1]);
    c += head[j] ? 1u : 0u;
  }
  uint32_t incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_w[w] = incl;
  __syncthreads();
  uint32_t g = bc[blockIdx.x] + incl - c;
  for (int i = 0; i < w; ++i) g += s_w[i];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (head[j]) {
      S[g] = k[j];
      start[g] = (uint32_t)(base + j);
      ++g;
    }
  }
}

__global__ void __launch_bounds__(256) uniq_pc_kernel(const uint32_t* __restrict__ start, const unsigned long long* __restrict__ Gp,
                                                      long long n, uint32_t* __restrict__ pc) {
  const long long G = (long long)*Gp;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < G) pc[g] = (g + 1 < G ? start[g + 1] : (uint32_t)n) - start[g];
}

// ---- buckets of consecutive positive groups ---------------------------------------------------------------
// bucket b = groups [b * cap_b, min((b + 1) * cap_b, G)); U[b] = its largest score; a negative belongs to the first
// bucket whose U is >= its score, or to the last bucket when it lies above every positive
__global__ void bucket_bounds_kernel(const uint32_t* __restrict__ S, long long G, int B, long long cap_b, uint32_t* __restrict__ U) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    const long long e = min((long long)(b + 1) * cap_b, G);
    U[b] = S[e - 1];
  }
}

__device__ __forceinline__ int bucket_of(const SmemTable& tab, uint32_t sk, int B) {
  int lo, hi;
  tab.range(sk, lo, hi);
  const int b = tab.finish(sk, lo, hi);
  return b < B ? b : B - 1;
}

// ---- grouping the negatives by bucket ----------------------------------------------------------------------------
// The key array is cut into n_slices contiguous slices, one per CTA, the same cut in every pass.
//   bucket_count_kernel    negatives of slice s per bucket -> cnt[s][b]
//   slice_prefix_kernel    rel[s][b] = exclusive prefix of cnt[s][*], tot[s]
//   bucket_plan_kernel     sbase[s] = exclusive prefix of tot; ranking units = (bucket, run of slices) of ~PR_UNIT keys
//   bucket_scatter_kernel  slice s writes ITS negatives, grouped by bucket, into ITS OWN region
//                          part[sbase[s] + rel[s][b] ...]: the 700+ write streams of a CTA stay inside a few MB (the
//                          bucket-major layout tried first spread them over the whole 5.5 GB array -- one TLB miss and
//                          one 4-byte partial-sector write per key, 25 ms for 1.38 G keys); keys are staged through
//                          shared memory per 8192-key tile so that every (tile, bucket) run leaves as one contiguous
//                          store.  Positions inside a tile come from shared-memory atomics with return value (the
//                          order inside a bucket is irrelevant: non-stable on purpose, no digit matching, no look-back).
//   unit_rank_kernel       a unit walks the chunks (slice s, bucket b) of its slices and ranks them against the bucket's
//                          positives in shared memory
constexpr int PT_THREADS = 512;           // scatter pass: 2 resident CTAs per SM hide each other's barriers
constexpr int PT_ITEMS = 16;
constexpr int PT_TILE = PT_THREADS * PT_ITEMS;

__global__ void __launch_bounds__(RANK_THREADS) bucket_count_kernel(const uint32_t* __restrict__ keys, long long n, long long slice_len,
                                                                    const uint32_t* __restrict__ U, int B, uint32_t key_base,
                                                                    uint32_t* __restrict__ cnt) {
  extern __shared__ uint32_t s_mem[];
  uint32_t* s_U = s_mem;                                             // [B]
  uint32_t* s_c = s_U + B;                                           // [B]
  uint32_t* s_lut = s_c + B;                                         // [RANK_LUT]
  const int tid = threadIdx.x;
  for (int i = tid; i < B; i += RANK_THREADS) { s_U[i] = U[i]; s_c[i] = 0u; }
  __syncthreads();
  SmemTable tab;
  tab.build(s_U, s_lut, B, key_base);
  const long long k0 = (long long)blockIdx.x * slice_len;
  const long long k1 = min(k0 + slice_len, n);
  auto visit = [&](uint32_t k) {
    const int b = bucket_of(tab, k >> 1, B);      // for every key: only the atomic is conditional (predicated, no divergence region)
    if (!(k & 1u)) atomicAdd(&s_c[b], 1u);
  };
  const bool aligned = (reinterpret_cast<uintptr_t>(keys + k0) & 15) == 0;   // slice_len is a multiple of 4
  const long long nvec = aligned && k1 > k0 ? (k1 - k0) / 4 : 0;
  for (long long q = tid; q < nvec; q += RANK_THREADS) {
    const uint4 t = *reinterpret_cast<const uint4*>(keys + k0 + q * 4);
    visit(t.x); visit(t.y); visit(t.z); visit(t.w);
  }
  for (long long i = k0 + nvec * 4 + tid; i < k1; i += RANK_THREADS) visit(keys[i]);
  __syncthreads();
  uint32_t* mine = cnt + (size_t)blockIdx.x * B;
  for (int i = tid; i < B; i += RANK_THREADS) mine[i] = s_c[i];
}

// block-wide exclusive scan of up to PR_MAX_BUCKETS values held PR_PER = PR_MAX_BUCKETS / 1024 per thread (thread t owns
// entries [t * PR_PER, (t + 1) * PR_PER)); returns the total
constexpr int PR_PER = PR_MAX_BUCKETS / RANK_THREADS;
template <int PER>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t (&v)[PER], uint32_t* s_w /*[33]*/) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nw = (int)(blockDim.x >> 5);
  uint32_t sum = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) sum += v[j];
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();              // s_w may still be read from a previous call
  if (lane == 31) s_w[w] = incl;
  __syncthreads();
  // second level: warp 0 scans the (<= 32) warp totals in place -> s_w[i] = exclusive prefix, s_w[32] = grand total
  if (w == 0) {
    const uint32_t t = lane < nw ? s_w[lane] : 0u;
    uint32_t it = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, it, o);
      if (lane >= o) it += u;
    }
    s_w[lane] = it - t;
    if (lane == 31) s_w[32] = it;
  }
  __syncthreads();
  uint32_t run = incl - sum + s_w[w];
  const uint32_t tot = s_w[32];
#pragma unroll
  for (int j = 0; j < PER; ++j) { const uint32_t c = v[j]; v[j] = run; run += c; }
  return tot;
}

__global__ void __launch_bounds__(RANK_THREADS) slice_prefix_kernel(const uint32_t* __restrict__ cnt, int B, uint32_t* __restrict__ rel,
                                                                    uint32_t* __restrict__ tot) {
  __shared__ uint32_t s_w[33];
  const uint32_t* c = cnt + (size_t)blockIdx.x * B;
  uint32_t v[PR_PER];
#pragma unroll
  for (int j = 0; j < PR_PER; ++j) { const int b = threadIdx.x * PR_PER + j; v[j] = b < B ? c[b] : 0u; }
  const uint32_t t = block_excl_scan(v, s_w);
#pragma unroll
  for (int j = 0; j < PR_PER; ++j) { const int b = threadIdx.x * PR_PER + j; if (b < B) rel[(size_t)blockIdx.x * B + b] = v[j]; }
  if (threadIdx.x == 0) tot[blockIdx.x] = t;
}

constexpr int PLAN_MLP = 16;
// slice bases + ranking units (single CTA).  A unit = bucket b, slices [s0, s1): consecutive slices until ~unit_len keys.
__global__ void __launch_bounds__(RANK_THREADS) bucket_plan_kernel(const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ tot,
                                                                   int n_slices, int B, uint32_t* __restrict__ sbase,
                                                                   Unit* __restrict__ units, uint32_t* __restrict__ n_units,
                                                                   long long unit_len) {
  __shared__ uint32_t s_w[33];
  const int tid = threadIdx.x;
  // (1) slice bases (n_slices <= PR_MAX_BUCKETS: PR_PER per thread)
  {
    uint32_t v[PR_PER];
#pragma unroll
    for (int j = 0; j < PR_PER; ++j) { const int sidx = tid * PR_PER + j; v[j] = sidx < n_slices ? tot[sidx] : 0u; }
    block_excl_scan(v, s_w);
#pragma unroll
    for (int j = 0; j < PR_PER; ++j) { const int sidx = tid * PR_PER + j; if (sidx < n_slices) sbase[sidx] = v[j]; }
  }
  // (2) units per bucket: count, scan, emit.  Bucket b = tid + 1024 j in the two walks over the slices (every thread busy for
  // B <= 1024), blocked ownership only for the scan in between.
  __shared__ uint32_t s_nu[PR_MAX_BUCKETS], s_ub[PR_MAX_BUCKETS];
  // PLAN_MLP independent loads in flight: one dependent L2 round trip per slice made this single CTA 0.25 ms
  for (int b = tid; b < B; b += RANK_THREADS) {
    unsigned long long acc = 0;
    uint32_t nub = 0u;
    for (int sb = 0; sb < n_slices; sb += PLAN_MLP) {
      uint32_t c[PLAN_MLP];
#pragma unroll
      for (int q = 0; q < PLAN_MLP; ++q) c[q] = sb + q < n_slices ? cnt[(size_t)(sb + q) * B + b] : 0u;
#pragma unroll
      for (int q = 0; q < PLAN_MLP; ++q) {
        acc += c[q];
        if (acc >= (unsigned long long)unit_len) { ++nub; acc = 0; }
      }
    }
    if (acc) ++nub;
    s_nu[b] = nub;
  }
  __syncthreads();
  uint32_t ub[PR_PER];
#pragma unroll
  for (int j = 0; j < PR_PER; ++j) { const int b = tid * PR_PER + j; ub[j] = b < B ? s_nu[b] : 0u; }
  const uint32_t total = block_excl_scan(ub, s_w);
  if (tid == 0) *n_units = total;
#pragma unroll
  for (int j = 0; j < PR_PER; ++j) { const int b = tid * PR_PER + j; if (b < B) s_ub[b] = ub[j]; }
  __syncthreads();
  for (int b = tid; b < B; b += RANK_THREADS) {
    if (!s_nu[b]) continue;
    uint32_t u = s_ub[b];
    unsigned long long acc = 0;
    int s0 = 0;
    for (int sb = 0; sb < n_slices; sb += PLAN_MLP) {
      uint32_t c[PLAN_MLP];
#pragma unroll
      for (int q = 0; q < PLAN_MLP; ++q) c[q] = sb + q < n_slices ? cnt[(size_t)(sb + q) * B + b] : 0u;
#pragma unroll
      for (int q = 0; q < PLAN_MLP; ++q) {
        const int sidx = sb + q;
        acc += c[q];
        if (acc >= (unsigned long long)unit_len) {       // (padding entries add 0 to acc = 0 or to acc < unit_len)
          units[u++] = Unit{(uint32_t)b, (uint32_t)s0, (uint32_t)(sidx + 1), 0u};
          s0 = sidx + 1; acc = 0;
        }
      }
    }
    if (acc) units[u++] = Unit{(uint32_t)b, (uint32_t)s0, (uint32_t)n_slices, 0u};
  }
}

__global__ void __launch_bounds__(PT_THREADS, 2) bucket_scatter_kernel(const uint32_t* __restrict__ keys, long long n, long long slice_len,
                                                                      const uint32_t* __restrict__ U, int B, uint32_t key_base,
                                                                      const uint32_t* __restrict__ rel, const uint32_t* __restrict__ sbase,
                                                                      uint32_t* __restrict__ part) {
  extern __shared__ uint32_t s_mem[];
  uint32_t* s_U = s_mem;                                             // [B]
  uint32_t* s_cur = s_U + B;                                         // [B] next write position of bucket b in `part`
  uint32_t* s_tc = s_cur + B;                                        // [B] tile count, then tile-local start
  uint32_t* s_delta = s_tc + B;                                      // [B] part index of staged position 0 of the bucket's run
  uint32_t* s_lut = s_delta + B;                                     // [RANK_LUT]
  uint32_t* s_stage = s_lut + RANK_LUT;                              // [PT_TILE] keys grouped by bucket
  unsigned short* s_sb = reinterpret_cast<unsigned short*>(s_stage + PT_TILE);   // [PT_TILE] their buckets
  __shared__ uint32_t s_w[33];
  const int tid = threadIdx.x;
  const uint32_t base = sbase[blockIdx.x];
  for (int i = tid; i < B; i += PT_THREADS) {
    s_U[i] = U[i];
    s_cur[i] = base + rel[(size_t)blockIdx.x * B + i];
    s_tc[i] = 0u;
  }
  __syncthreads();
  SmemTable tab;
  tab.build<PT_THREADS>(s_U, s_lut, B, key_base);
  const long long k0 = (long long)blockIdx.x * slice_len;
  const long long k1 = min(k0 + slice_len, n);
  const bool aligned = (reinterpret_cast<uintptr_t>(keys + k0) & 15) == 0;
  for (long long t0 = k0; t0 < k1; t0 += PT_TILE) {
    uint32_t k[PT_ITEMS], br[PT_ITEMS];
#pragma unroll
    for (int v = 0; v < PT_ITEMS / 4; ++v) {
      const long long i0 = t0 + ((long long)v * PT_THREADS + tid) * 4;
      if (aligned && i0 + 4 <= k1) {
        const uint4 t = *reinterpret_cast<const uint4*>(keys + i0);
        k[4 * v] = t.x; k[4 * v + 1] = t.y; k[4 * v + 2] = t.z; k[4 * v + 3] = t.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) k[4 * v + j] = (i0 + j < k1) ? keys[i0 + j] : 1u;   // past the end: dropped like a positive
      }
    }
#pragma unroll
    for (int j = 0; j < PT_ITEMS; ++j) {
      // the bucket is looked up for every key (1 % positives): only the atomic stays conditional, so the compiler predicates
      // it instead of opening a divergence region around the whole search
      const int b = bucket_of(tab, k[j] >> 1, B);
      const bool neg = !(k[j] & 1u);
      uint32_t r = 0u;
      if (neg) r = atomicAdd(&s_tc[b], 1u);                         // rank inside (tile, bucket) < PT_TILE <= 65536
      br[j] = neg ? (((uint32_t)b << 16) | r) : 0xffffffffu;
    }
    __syncthreads();
    // tile-local starts of the buckets; advance the running cursors
    constexpr int SP = PR_MAX_BUCKETS / PT_THREADS;
    uint32_t c[SP], st[SP];
#pragma unroll
    for (int j = 0; j < SP; ++j) { const int b = tid * SP + j; c[j] = b < B ? s_tc[b] : 0u; st[j] = c[j]; }
    const uint32_t total = block_excl_scan(st, s_w);
#pragma unroll
    for (int j = 0; j < SP; ++j) {
      const int b = tid * SP + j;
      if (b < B) {
        s_tc[b] = st[j];
        s_delta[b] = s_cur[b] - st[j];
        s_cur[b] += c[j];
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PT_ITEMS; ++j) {
      if (br[j] != 0xffffffffu) {
        const uint32_t b = br[j] >> 16, p = s_tc[b] + (br[j] & 0xffffu);
        s_stage[p] = k[j];
        s_sb[p] = (unsigned short)b;
      }
    }
    __syncthreads();
    for (uint32_t p = tid; p < total; p += PT_THREADS) part[s_delta[s_sb[p]] + p] = s_stage[p];
    __syncthreads();
    for (int i = tid; i < B; i += PT_THREADS) s_tc[i] = 0u;
    __syncthreads();
  }
}

// one unit = the chunks (slice s, bucket b), s in [s0, s1), of one bucket: ~PR_UNIT grouped negatives, ranked against
// the bucket's positive groups in shared memory
template <bool FLAT>
__global__ void __launch_bounds__(RANK_THREADS, 1) unit_rank_kernel(const uint32_t* __restrict__ part, const Unit* __restrict__ units,
                                                                    const uint32_t* __restrict__ n_units, const uint32_t* __restrict__ cntt,
                                                                    const uint32_t* __restrict__ rel, const uint32_t* __restrict__ sbase,
                                                                    const uint32_t* __restrict__ S, long long G, long long cap_b, int B,
                                                                    uint32_t key_base, unsigned long long* __restrict__ cnt) {
  extern __shared__ uint32_t s_mem[];
  if (blockIdx.x >= *n_units) return;
  const Unit u = units[blockIdx.x];
  const long long g0 = (long long)u.bucket * cap_b;
  const int gn = (int)(min(g0 + cap_b, G) - g0);
  uint32_t* s_S = s_mem;                                                        // [cap_b + 1]: sentinel behind the last score
  uint32_t* s_cnt = s_S + cap_b + 1;                                            // [2 cap_b + 2]
  uint32_t* s_lut = s_cnt + 2 * cap_b + 2;
  const int tid = threadIdx.x;
  // The unit's chunks (slice, bucket) as ONE flat sequence of aligned 16-byte vectors: chunk j covers the vectors
  // [coff >> 2, (coff + len + 3) >> 2) of `part` (elements outside [coff, coff + len) masked), s_vpre = exclusive prefix of
  // the vector counts.  Every thread strides over the flat index, so the work is balanced whatever the chunk lengths are
  // (the r2l form -- long chunks by the whole CTA, short ones one warp each -- spent 60 % of a 1/8 shard's time fetching
  // descriptors one dependent L2 round trip at a time and at the final barrier behind the warp with the longest chunk).
  __shared__ uint32_t s_clen[PR_MAX_SLICES], s_coff[PR_MAX_SLICES], s_vpre[PR_MAX_SLICES + 1];
  const uint32_t n_ch = u.s1 - u.s0;
  for (uint32_t j = tid; j < n_ch; j += RANK_THREADS) {
    const size_t e = (size_t)(u.s0 + j) * B + u.bucket;
    const uint32_t len = cntt[e], off = sbase[u.s0 + j] + rel[e];
    s_clen[j] = len;
    s_coff[j] = off;
    s_vpre[j + 1] = len ? ((off + len + 3u) >> 2) - (off >> 2) : 0u;
  }
  for (int i = tid; i < gn; i += RANK_THREADS) s_S[i] = S[g0 + i];
  if (tid == 0) s_S[gn] = 0xffffffffu;              // no 31-bit score key equals it: no bound test on l below
  for (int i = tid; i < 2 * gn + 2; i += RANK_THREADS) s_cnt[i] = 0u;
  __syncthreads();
  if (tid < 32) {                                   // warp 0: inclusive scan of the vector counts, 32 at a time
    uint32_t carry = 0u;
    for (uint32_t b0 = 0; b0 < n_ch; b0 += 32) {
      const uint32_t j = b0 + tid;
      uint32_t v = j < n_ch ? s_vpre[j + 1] : 0u;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (tid >= o) v += t;
      }
      v += carry;
      if (j < n_ch) s_vpre[j + 1] = v;
      carry = __shfl_sync(0xffffffffu, v, 31);
    }
    if (tid == 0) s_vpre[0] = 0u;
  }
  SmemTable tab;
  tab.build(s_S, s_lut, gn, key_base);              // (barriers inside: s_vpre is complete behind them)
  auto rank_one = [&](uint32_t key) {
    const uint32_t sk = key >> 1;
    int lo, hi;
    tab.range(sk, lo, hi);
    const int l = tab.finish(sk, lo, hi);
    atomicAdd(&s_cnt[2 * l + (s_S[l] == sk ? 1 : 0)], 1u);
  };
  if constexpr (FLAT) {
    const uint32_t V = s_vpre[n_ch];
    const uint4* pv = reinterpret_cast<const uint4*>(part);           // (the workspace is 256-byte aligned)
    constexpr int U = 4;                                              // independent 16-byte loads in flight per thread
    uint32_t j = 0;                                                   // chunk of the previous vector: the flat index only grows
    for (uint32_t q0 = tid; q0 < V; q0 += U * RANK_THREADS) {
      uint4 v[U];
      uint32_t rel_e[U], len_e[U];
  #pragma unroll
      for (int k = 0; k < U; ++k) {
        const uint32_t q = q0 + k * RANK_THREADS;
        v[k] = make_uint4(0u, 0u, 0u, 0u);
        rel_e[k] = 0u; len_e[k] = 0u;
        if (q < V) {
          if (q >= s_vpre[j + 1]) {                                   // first chunk with s_vpre[j + 1] > q
            uint32_t lo = j + 1, hi = n_ch - 1;
            while (lo < hi) {
              const uint32_t mid = (lo + hi) >> 1;
              if (s_vpre[mid + 1] > q) hi = mid; else lo = mid + 1;
            }
            j = lo;
          }
          const uint32_t off = s_coff[j];
          const uint32_t vec = (off >> 2) + (q - s_vpre[j]);
          v[k] = pv[vec];
          rel_e[k] = vec * 4u - off;                                  // element c of the vector is chunk element rel_e + c (mod 2^32)
          len_e[k] = s_clen[j];
        }
      }
  #pragma unroll
      for (int k = 0; k < U; ++k) {
        if (rel_e[k] + 0u < len_e[k]) rank_one(v[k].x);
        if (rel_e[k] + 1u < len_e[k]) rank_one(v[k].y);
        if (rel_e[k] + 2u < len_e[k]) rank_one(v[k].z);
        if (rel_e[k] + 3u < len_e[k]) rank_one(v[k].w);
      }
    }
  } else {
    // long chunks by the whole CTA, short ones one warp each (32 chunks in flight per CTA)
    auto rank_chunk = [&](const uint32_t* p0, uint32_t len_all, int t, int nthr) {
      uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(p0) & 15)) & 15) / 4;
      if (head > len_all) head = len_all;
      if (t < (int)head) rank_one(p0[t]);
      const uint32_t* p = p0 + head;
      const uint32_t len = len_all - head;
      const uint32_t nvec = len / 4;
      for (uint32_t q = t; q < nvec; q += 4 * nthr) {
        uint4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t qq = q + k * nthr;
          v[k] = qq < nvec ? *reinterpret_cast<const uint4*>(p + (size_t)qq * 4) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (q + k * nthr < nvec) { rank_one(v[k].x); rank_one(v[k].y); rank_one(v[k].z); rank_one(v[k].w); }
        }
      }
      for (uint32_t i = nvec * 4 + t; i < len; i += nthr) rank_one(p[i]);
    };
    constexpr uint32_t LONG_CHUNK = 32768;
    for (uint32_t j = 0; j < n_ch; ++j) {
      const uint32_t len_all = s_clen[j];
      if (len_all >= LONG_CHUNK) rank_chunk(part + s_coff[j], len_all, tid, RANK_THREADS);
    }
    for (uint32_t j = (uint32_t)(tid >> 5); j < n_ch; j += RANK_THREADS / 32) {
      const uint32_t len_all = s_clen[j];
      if (len_all > 0u && len_all < LONG_CHUNK) rank_chunk(part + s_coff[j], len_all, tid & 31, 32);
    }
  }
  __syncthreads();
  unsigned long long* c = cnt + 2 * g0;
  for (int i = tid; i < 2 * gn + 1; i += RANK_THREADS)
    if (s_cnt[i]) atomicAdd(c + i, (unsigned long long)s_cnt[i]);
}

// ---- scan over the G positive groups (pc[g]; cnt[2g] = bt[g], cnt[2g+1] = eq[g], cnt[2G] = bt[G]) -----------------
struct BlockTot { unsigned long long T, F; };
struct BlockPart { unsigned long long au; double ap; long long gs; long long pad; };

__global__ void __launch_bounds__(1024) pscan_sums_kernel(const uint32_t* __restrict__ pc, const unsigned long long* __restrict__ cnt,
                                                          long long G, BlockTot* __restrict__ tot) {
  __shared__ unsigned long long s_T[32], s_F[32];
  const long long g0 = (long long)blockIdx.x * SCAN_GROUPS_PER_BLOCK + (long long)threadIdx.x * 8;
  unsigned long long T = 0, F = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const long long g = g0 + j;
    if (g < G) { T += pc[g]; F += cnt[2 * g] + cnt[2 * g + 1]; }
  }
  T = warp_reduce_sum_u64(T); F = warp_reduce_sum_u64(F);
  if ((threadIdx.x & 31) == 0) { s_T[threadIdx.x >> 5] = T; s_F[threadIdx.x >> 5] = F; }
  __syncthreads();
  if (threadIdx.x == 0) {
    BlockTot t = {0ull, 0ull};
    for (int i = 0; i < 32; ++i) { t.T += s_T[i]; t.F += s_F[i]; }
    tot[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(1024) pscan_prefix_kernel(BlockTot* __restrict__ tot, long long nblk) {
  __shared__ unsigned long long s_T[32], s_F[32];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const long long per = (nblk + 1023) / 1024;
  const long long b0 = min((long long)tid * per, nblk), b1 = min(b0 + per, nblk);
  unsigned long long T = 0, F = 0;
  for (long long i = b0; i < b1; ++i) { T += tot[i].T; F += tot[i].F; }
  unsigned long long iT = T, iF = F;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long a = __shfl_up_sync(0xffffffffu, iT, o), b = __shfl_up_sync(0xffffffffu, iF, o);
    if (lane >= o) { iT += a; iF += b; }
  }
  if (lane == 31) { s_T[w] = iT; s_F[w] = iF; }
  __syncthreads();
  unsigned long long rT = iT - T, rF = iF - F;
  for (int i = 0; i < w; ++i) { rT += s_T[i]; rF += s_F[i]; }
  for (long long i = b0; i < b1; ++i) {
    const BlockTot t = tot[i];
    tot[i].T = rT; tot[i].F = rF;
    rT += t.T; rF += t.F;
  }
}

__global__ void __launch_bounds__(1024) pscan_apply_kernel(const uint32_t* __restrict__ pc, const unsigned long long* __restrict__ cnt,
                                                           long long G, const BlockTot* __restrict__ tot, long long total_pos,
                                                           double recall_level, BlockPart* __restrict__ parts) {
  __shared__ unsigned long long s_T[32], s_F[32], s_au[32];
  __shared__ double s_ap[32];
  __shared__ long long s_gs[32];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const long long g0 = (long long)blockIdx.x * SCAN_GROUPS_PER_BLOCK + (long long)tid * 8;
  unsigned long long p[8], bt[8], eq[8], T = 0, F = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const long long g = g0 + j;
    const bool in = g < G;
    p[j] = in ? pc[g] : 0ull; bt[j] = in ? cnt[2 * g] : 0ull; eq[j] = in ? cnt[2 * g + 1] : 0ull;
    T += p[j]; F += bt[j] + eq[j];
  }
  unsigned long long iT = T, iF = F;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long a = __shfl_up_sync(0xffffffffu, iT, o), b = __shfl_up_sync(0xffffffffu, iF, o);
    if (lane >= o) { iT += a; iF += b; }
  }
  if (lane == 31) { s_T[w] = iT; s_F[w] = iF; }
  __syncthreads();
  unsigned long long rT = tot[blockIdx.x].T + iT - T, rF = tot[blockIdx.x].F + iF - F;
  for (int i = 0; i < w; ++i) { rT += s_T[i]; rF += s_F[i]; }
  const long long tstar = recall_threshold(total_pos, recall_level);
  unsigned long long au = 0ull;
  double ap = 0.0;
  long long gs = -1;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (g0 + j < G) {
      au += bt[j] * 2ull * rT;
      rT += p[j];
      rF += bt[j] + eq[j];
      au += eq[j] * (2ull * rT - p[j]);
      ap += (double)p[j] * ((double)rT / (double)(rT + rF));
      if ((long long)rT <= tstar) gs = g0 + j;
    }
  }
  au = warp_reduce_sum_u64(au);
  ap = warp_reduce_sum_d(ap);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gs = max(gs, __shfl_xor_sync(0xffffffffu, gs, o));
  if (lane == 0) { s_au[w] = au; s_ap[w] = ap; s_gs[w] = gs; }
  __syncthreads();
  if (tid == 0) {
    BlockPart b = {0ull, 0.0, -1, 0};
    for (int i = 0; i < 32; ++i) { b.au += s_au[i]; b.ap += s_ap[i]; b.gs = max(b.gs, s_gs[i]); }
    parts[blockIdx.x] = b;
  }
}

// cumulative (tps, fps) through group g (g >= 0), all threads of the CTA cooperate
__device__ void cta_prefix(const uint32_t* __restrict__ pc, const unsigned long long* __restrict__ cnt, const BlockTot* __restrict__ tot,
                           long long g, unsigned long long* s_red, unsigned long long& Tg, unsigned long long& Fg) {
  const long long blk = g / SCAN_GROUPS_PER_BLOCK;
  unsigned long long T = 0, F = 0;
  for (long long k = blk * SCAN_GROUPS_PER_BLOCK + threadIdx.x; k <= g; k += blockDim.x) { T += pc[k]; F += cnt[2 * k] + cnt[2 * k + 1]; }
  T = warp_reduce_sum_u64(T); F = warp_reduce_sum_u64(F);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5] = T; s_red[32 + (threadIdx.x >> 5)] = F; }
  __syncthreads();
  Tg = tot[blk].T; Fg = tot[blk].F;
  for (int i = 0; i < 32; ++i) { Tg += s_red[i]; Fg += s_red[32 + i]; }
}

__global__ void __launch_bounds__(1024) pscan_final_kernel(const uint32_t* __restrict__ pc, const unsigned long long* __restrict__ cnt,
                                                           long long G, const BlockTot* __restrict__ tot, const BlockPart* __restrict__ parts,
                                                           long long nblk, long long total_pos, long long total_n, long long n_nan,
                                                           double recall_level, dml_ood_result* __restrict__ result) {
  __shared__ unsigned long long s_red[64];
  __shared__ unsigned long long s_au[32];
  __shared__ double s_ap[32];
  __shared__ long long s_gs[32];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const long long P = total_pos, N = total_n - total_pos;
  dml_ood_result o;
  o.n_pos = P; o.n_neg = N; o.n_nan = n_nan; o.n_groups = -1;
  if (P <= 0 || N <= 0 || G <= 0) {
    if (tid == 0) {
      const double nan = __longlong_as_double(0x7ff8000000000000ll);
      o.auroc = o.aupr = o.fpr = nan;
      *result = o;
    }
    return;
  }
  // fixed-order reduction of the block partials: thread t owns a contiguous slice
  const long long per = (nblk + 1023) / 1024;
  const long long b0 = min((long long)tid * per, nblk), b1 = min(b0 + per, nblk);
  unsigned long long au = 0ull;
  double ap = 0.0;
  long long gs = -1;
  for (long long i = b0; i < b1; ++i) { au += parts[i].au; ap += parts[i].ap; gs = max(gs, parts[i].gs); }
  au = warp_reduce_sum_u64(au);
  ap = warp_reduce_sum_d(ap);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) gs = max(gs, __shfl_xor_sync(0xffffffffu, gs, off));
  if (lane == 0) { s_au[w] = au; s_ap[w] = ap; s_gs[w] = gs; }
  __syncthreads();
  au = 0ull; ap = 0.0; gs = -1;
  for (int i = 0; i < 32; ++i) { au += s_au[i]; ap += s_ap[i]; gs = max(gs, s_gs[i]); }
  au += 2ull * (unsigned long long)P * cnt[2 * G];
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  double da = inf, db = inf;
  unsigned long long a_fps = 0, b_fps = 0;
  {
    unsigned long long Tg = 0, Fg = 0;
    if (gs >= 0) cta_prefix(pc, cnt, tot, gs, s_red, Tg, Fg);
    const unsigned long long trail = (gs + 1 <= G - 1) ? cnt[2 * (gs + 1)] : 0ull;
    if (gs >= 0 || trail > 0) {
      da = fabs((double)Tg / (double)P - recall_level);
      a_fps = Fg + trail;
    }
  }
  if (gs + 1 <= G - 1) {
    unsigned long long Tg, Fg;
    cta_prefix(pc, cnt, tot, gs + 1, s_red, Tg, Fg);
    const unsigned long long trail = (gs + 2 <= G - 1) ? cnt[2 * (gs + 2)] : 0ull;
    db = fabs((double)Tg / (double)P - recall_level);
    b_fps = Fg + trail;
  }
  if (tid == 0) {
    o.auroc = (double)au / (2.0 * (double)P * (double)N);
    o.aupr = ap / (double)P;
    o.fpr = (double)(db <= da ? b_fps : a_fps) / (double)N;
    *result = o;
  }
}

// Fixed-capacity slot records (header word 0..1 = int64 key count, keys from word `hdr`) -> one contiguous key list, in
// slot order.  The records are what the ranks all-gather batch by batch while the per-image pass is still running
// (distributed.PositiveExchange); grid (n_slots, SG_SPLIT).
constexpr int SG_SPLIT = 8;
__global__ void __launch_bounds__(256) slots_gather_kernel(const uint32_t* __restrict__ slots, int n_slots, long long stride, int hdr,
                                                           long long cap, uint32_t* __restrict__ out, long long out_capacity) {
  __shared__ unsigned long long s_part[8];
  const int b = blockIdx.x;
  auto count_of = [&](int i) -> unsigned long long {
    const long long c = *reinterpret_cast<const long long*>(slots + (size_t)i * stride);
    return (unsigned long long)(c < 0 ? 0 : (c > cap ? cap : c));
  };
  unsigned long long acc = 0;
  for (int i = threadIdx.x; i < b; i += 256) acc += count_of(i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  unsigned long long base = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) base += s_part[w];
  const unsigned long long cnt = count_of(b);
  if (base + cnt > (unsigned long long)out_capacity) return;
  const uint32_t* src = slots + (size_t)b * stride + hdr;
  for (unsigned long long i = (unsigned long long)blockIdx.y * 256 + threadIdx.x; i < cnt; i += 256ull * SG_SPLIT) out[base + i] = src[i];
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct BucketPlan {
  int B, n_slices;
  long long cap_b, slice_len;
  long long max_units, unit_len;
  size_t off_part, off_U, off_cnt, off_rel, off_tot, off_sbase, off_units, off_nunits, off_end;
};
// returns false when the positives have more distinct scores than PR_MAX_BUCKETS buckets can hold
bool make_bucket_plan(long long n, long long G, BucketPlan& p) {
  p.B = (int)((G + PR_PASS_CAP - 1) / PR_PASS_CAP);
  if (p.B < 1) p.B = 1;
  if (p.B > PR_MAX_BUCKETS) return false;
  p.cap_b = (G + p.B - 1) / p.B;
  if (p.cap_b < 1) p.cap_b = 1;
  p.unit_len = n / (148 * 8);
  if (p.unit_len > PR_UNIT) p.unit_len = PR_UNIT;
  if (p.unit_len < 65536) p.unit_len = 65536;
  p.max_units = n / p.unit_len + p.B + 1;
  // one slice per CTA of the counting / scatter passes: 2 CTAs per SM, at least 16384 keys each
  long long ns = (n + 16383) / 16384;
  if (ns > PR_MAX_SLICES) ns = PR_MAX_SLICES;
  if (ns < 1) ns = 1;
  p.n_slices = (int)ns;
  p.slice_len = (((n + ns - 1) / ns) + 3) & ~3ll;
  p.off_part = 0;
  p.off_U = align256(p.off_part + (size_t)n * sizeof(uint32_t));
  p.off_cnt = align256(p.off_U + (size_t)p.B * sizeof(uint32_t));                               // [n_slices][B] negatives per (slice, bucket)
  p.off_rel = align256(p.off_cnt + (size_t)p.n_slices * p.B * sizeof(uint32_t));                // [n_slices][B] their start inside the slice region
  p.off_tot = align256(p.off_rel + (size_t)p.n_slices * p.B * sizeof(uint32_t));
  p.off_sbase = align256(p.off_tot + (size_t)p.n_slices * sizeof(uint32_t));
  p.off_units = align256(p.off_sbase + (size_t)p.n_slices * sizeof(uint32_t));
  p.off_nunits = align256(p.off_units + (size_t)p.max_units * sizeof(Unit));
  p.off_end = align256(p.off_nunits + 256);
  return true;
}

}  // namespace
}  // namespace dml

using namespace dml;

extern "C" {
#pragma GCC visibility push(default)

int dml_ood_pos_compact(const uint32_t* keys, int64_t n, uint32_t* pos_keys_out, int64_t capacity, long long* count,
                        dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || capacity < 0 || !count || (n > 0 && !keys) || (capacity > 0 && !pos_keys_out)) return DML_ERR_INVALID_ARG;
  DML_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(long long), stream));
  if (n == 0) return DML_OK;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  pos_compact_kernel<<<(unsigned)blocks, 256, 0, stream>>>(keys, n, pos_keys_out, capacity, (unsigned long long*)count);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

int dml_ood_slots_gather(const uint32_t* slots, int32_t n_slots, int64_t stride_words, int32_t hdr_words, int64_t slot_capacity,
                         uint32_t* out, int64_t out_capacity, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_slots < 0 || n_slots > 65535 || hdr_words < 2 || slot_capacity < 0 || stride_words < hdr_words + slot_capacity ||
      (stride_words & 1) || out_capacity < 0 || (n_slots > 0 && !slots) || (out_capacity > 0 && !out))
    return DML_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(slots) & 7) != 0) return DML_ERR_INVALID_ARG;      // the int64 counts are read in place
  if (n_slots == 0) return DML_OK;
  slots_gather_kernel<<<dim3((unsigned)n_slots, SG_SPLIT), 256, 0, stream>>>(slots, n_slots, stride_words, hdr_words, slot_capacity, out,
                                                                             out_capacity);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

size_t dml_ood_unique_workspace_bytes(int64_t n) {
  if (n <= 0) return 256;
  const size_t nblk = (size_t)((n + UNIQ_TILE - 1) / UNIQ_TILE);
  return align256(nblk * sizeof(uint32_t)) + align256((size_t)n * sizeof(uint32_t)) + 256;
}

int dml_ood_unique_counts(const uint32_t* sorted_keys, int64_t n, uint32_t* values_out, uint32_t* counts_out, long long* n_unique,
                          void* workspace, size_t workspace_bytes, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || n >= (1ll << 32) || !n_unique || (n > 0 && (!sorted_keys || !values_out || !counts_out || !workspace))) return DML_ERR_INVALID_ARG;
  if (n == 0) {
    DML_CUDA_TRY(cudaMemsetAsync(n_unique, 0, sizeof(long long), stream));
    return DML_OK;
  }
  if (workspace_bytes < dml_ood_unique_workspace_bytes(n)) return DML_ERR_WORKSPACE;
  const long long nblk = (n + UNIQ_TILE - 1) / UNIQ_TILE;
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  uint32_t* bc = reinterpret_cast<uint32_t*>(ws);
  uint32_t* start = reinterpret_cast<uint32_t*>(ws + align256((size_t)nblk * sizeof(uint32_t)));
  uniq_count_kernel<<<(unsigned)nblk, 1024, 0, stream>>>(sorted_keys, n, bc);
  DML_LAUNCH_CHECK();
  scan_u32_kernel<<<1, 1024, 0, stream>>>(bc, nblk, (unsigned long long*)n_unique);
  DML_LAUNCH_CHECK();
  uniq_write_kernel<<<(unsigned)nblk, 1024, 0, stream>>>(sorted_keys, n, bc, values_out, start);
  DML_LAUNCH_CHECK();
  uniq_pc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(start, (const unsigned long long*)n_unique, n, counts_out);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

size_t dml_ood_bucket_rank_workspace_bytes(int64_t n, int64_t n_groups) {
  BucketPlan p;
  if (n <= 0 || n_groups <= 0 || !make_bucket_plan(n, n_groups, p)) return 256;
  return p.off_end;
}

int dml_ood_bucket_rank(const uint32_t* keys, int64_t n, const uint32_t* group_scores, int64_t n_groups, uint32_t key_base,
                        unsigned long long* counters, void* workspace, size_t workspace_bytes, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || n >= (1ll << 32) || n_groups < 1 || !group_scores || !counters || !workspace || (n > 0 && !keys)) return DML_ERR_INVALID_ARG;
  BucketPlan p;
  if (!make_bucket_plan(n, n_groups, p)) return DML_ERR_UNSUPPORTED_DIM;
  if (workspace_bytes < p.off_end) return DML_ERR_WORKSPACE;
  DML_CUDA_TRY(cudaMemsetAsync(counters, 0, (2 * (size_t)n_groups + 2) * sizeof(unsigned long long), stream));
  if (n == 0) return DML_OK;
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  uint32_t* part = reinterpret_cast<uint32_t*>(ws + p.off_part);
  uint32_t* U = reinterpret_cast<uint32_t*>(ws + p.off_U);
  uint32_t* cntt = reinterpret_cast<uint32_t*>(ws + p.off_cnt);
  uint32_t* rel = reinterpret_cast<uint32_t*>(ws + p.off_rel);
  uint32_t* tot = reinterpret_cast<uint32_t*>(ws + p.off_tot);
  uint32_t* sbase = reinterpret_cast<uint32_t*>(ws + p.off_sbase);
  Unit* units = reinterpret_cast<Unit*>(ws + p.off_units);
  uint32_t* n_units = reinterpret_cast<uint32_t*>(ws + p.off_nunits);
  bucket_bounds_kernel<<<ceil_div_i(p.B, 256), 256, 0, stream>>>(group_scores, n_groups, p.B, p.cap_b, U);
  DML_LAUNCH_CHECK();
  const size_t smem_c = (size_t)2 * p.B * sizeof(uint32_t) + SmemTable::lut_bytes() + 16;
  const size_t smem_s = (size_t)4 * p.B * sizeof(uint32_t) + SmemTable::lut_bytes() + (size_t)PT_TILE * 6 + 16;
  DML_CUDA_TRY(cudaFuncSetAttribute(bucket_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
  DML_CUDA_TRY(cudaFuncSetAttribute(bucket_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
  bucket_count_kernel<<<(unsigned)p.n_slices, RANK_THREADS, smem_c, stream>>>(keys, n, p.slice_len, U, p.B, key_base, cntt);
  DML_LAUNCH_CHECK();
  slice_prefix_kernel<<<(unsigned)p.n_slices, RANK_THREADS, 0, stream>>>(cntt, p.B, rel, tot);
  DML_LAUNCH_CHECK();
  bucket_plan_kernel<<<1, RANK_THREADS, 0, stream>>>(cntt, tot, p.n_slices, p.B, sbase, units, n_units, p.unit_len);
  DML_LAUNCH_CHECK();
  bucket_scatter_kernel<<<(unsigned)p.n_slices, PT_THREADS, smem_s, stream>>>(keys, n, p.slice_len, U, p.B, key_base, rel, sbase, part);
  DML_LAUNCH_CHECK();
  const size_t smem_u = ((size_t)p.cap_b + 1) * 4 + ((size_t)2 * p.cap_b + 2) * 4 + SmemTable::lut_bytes() + 16;
  // chunk walk: flat (balanced for any chunk length) when the mean (slice, bucket) chunk is short, else long chunks by the
  // CTA + short ones per warp; DML_UNIT_RANK=flat|chunk overrides (A/B runs)
  bool flat = n / ((long long)p.n_slices * p.B) < PR_FLAT_BELOW;
  if (const char* e = getenv("DML_UNIT_RANK")) flat = e[0] == 'f';
  if (flat) {
    DML_CUDA_TRY(cudaFuncSetAttribute(unit_rank_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_u));
    unit_rank_kernel<true><<<(unsigned)p.max_units, RANK_THREADS, smem_u, stream>>>(part, units, n_units, cntt, rel, sbase, group_scores,
                                                                                   n_groups, p.cap_b, p.B, key_base, counters);
  } else {
    DML_CUDA_TRY(cudaFuncSetAttribute(unit_rank_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_u));
    unit_rank_kernel<false><<<(unsigned)p.max_units, RANK_THREADS, smem_u, stream>>>(part, units, n_units, cntt, rel, sbase, group_scores,
                                                                                    n_groups, p.cap_b, p.B, key_base, counters);
  }
  DML_LAUNCH_CHECK();
  return DML_OK;
}

size_t dml_ood_pooled_scan_workspace_bytes(int64_t n_groups) {
  if (n_groups <= 0) return 256;
  const size_t nblk = (size_t)((n_groups + SCAN_GROUPS_PER_BLOCK - 1) / SCAN_GROUPS_PER_BLOCK);
  return align256(nblk * sizeof(BlockTot)) + align256(nblk * sizeof(BlockPart)) + 256;
}

int dml_ood_pooled_scan(const uint32_t* group_counts, const unsigned long long* counters, int64_t n_groups, int64_t total_pos,
                        int64_t total_n, int64_t n_nan, double recall_level, void* workspace, size_t workspace_bytes,
                        dml_ood_result* result, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!result || n_groups < 0 || total_pos < 0 || total_n < total_pos) return DML_ERR_INVALID_ARG;
  if (n_groups > 0 && (!group_counts || !counters || !workspace)) return DML_ERR_INVALID_ARG;
  if (workspace_bytes < dml_ood_pooled_scan_workspace_bytes(n_groups)) return DML_ERR_WORKSPACE;
  const long long nblk = n_groups > 0 ? (n_groups + SCAN_GROUPS_PER_BLOCK - 1) / SCAN_GROUPS_PER_BLOCK : 0;
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  BlockTot* tot = reinterpret_cast<BlockTot*>(ws);
  BlockPart* parts = reinterpret_cast<BlockPart*>(ws + align256((size_t)nblk * sizeof(BlockTot)));
  if (nblk > 0) {
    pscan_sums_kernel<<<(unsigned)nblk, 1024, 0, stream>>>(group_counts, counters, n_groups, tot);
    DML_LAUNCH_CHECK();
    pscan_prefix_kernel<<<1, 1024, 0, stream>>>(tot, nblk);
    DML_LAUNCH_CHECK();
    pscan_apply_kernel<<<(unsigned)nblk, 1024, 0, stream>>>(group_counts, counters, n_groups, tot, total_pos, recall_level, parts);
    DML_LAUNCH_CHECK();
  }
  pscan_final_kernel<<<1, 1024, 0, stream>>>(group_counts, counters, n_groups, tot, parts, nblk, total_pos, total_n, n_nan, recall_level,
                                             result);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

#pragma GCC visibility pop
}  // extern "C"
