// Exact, tie-aware AUROC / AUPR / FPR@recall on the GPU (north-star kernel (d)).
//
//   key-gen  : score -> order-preserving packed u32 key, positive flag in bit 0
//   sort     : ood_sort.cu (segmented LSD radix sort)
//   scan     : segmented scan over distinct-score groups -> per-tile partial sums
//   finalize : fixed-order reduction of the partials -> (auroc, aupr, fpr) per segment
//
// Replaces anomaly/anom_utils.py:25-78 (fpr_and_fdr_at_recall, get_measures -> sklearn
// roc_auc_score / average_precision_score) and the callers' masking at
// anomaly/eval_ood_traditional.py:128-148.  Math, groups g of equal
// score in descending-score order, cumulative tps_g / fps_g:
//   AUROC = sum_g neg_g (2 tps_g - pos_g) / (2 P N)        [== trapezoid over the ROC points]
//   AUPR  = sum_g pos_g * tps_g / (tps_g + fps_g) / P
//   FPR   = fps_{g*} / N,  g* = argmin_g |tps_g / P - recall| over groups with tps_{g-1} < P,
//           ties -> the LATER group (anom_utils.py:57-65 scans from the lowest threshold down).
#include <cstdlib>

#include "ood_sort.cuh"
#include "ood_scan_thread.cuh"

namespace dml {

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;

// ---------------------------------------------------------------------------------------------
// key generation
// ---------------------------------------------------------------------------------------------
// pack_key(): ood_scan_thread.cuh (host / device, exercised on the CPU by tests/test_scan_emulation.py)

template <typename GT>
__device__ __forceinline__ bool is_positive(const GT* gt, const uint8_t* pos_u8, uint64_t mask, size_t i) {
  if (pos_u8) return pos_u8[i] != 0;
  const long long g = (long long)gt[i];
  return g >= 0 && g < 64 && ((mask >> g) & 1ull);
}

// Optional extra outputs of key generation: the normalised max-softmax map (MMSP) and the
// EDS/MMSP mix (anomaly/eval_ood_traditional.py:434-435,447-448), so that the raw EDS map is read
// once for normalisation, mix and ranking key.
struct KeygenFuse {
  const float* msp;   // raw max-softmax [n_seg*seg_len] (uses minmax slot 1)
  float* msp_norm;
  float* mix;
  float lambda, thr;
};

struct HistArgs {
  uint32_t* ghist;   // [n_seg][MAX_PASSES][RADIX] inside the sort workspace (zeroed by the caller)
  int n_passes;
  int shifts[MAX_PASSES];
  int top_atomic;    // count the (skewed) top digit place with shared atomics too instead of ballot groups
};
__device__ __forceinline__ unsigned lanemask_lt_() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
// HIST: the digit histograms of all radix passes (the sort's up-front counting read) are accumulated here, while
// the keys are still in registers: key-gen is HBM-bound and the histogram is shared-atomic-bound, so the two overlap
// and the sort's separate 4 B/key histogram pass (hist_kernel) disappears.  Same scheme as hist_kernel: warp-private
// shared histograms, plain shared atomics for the (near-uniform) lower digit places, ballot-grouped adds for the
// heavily skewed top place.
template <typename GT, int VEC, bool HIST>
__global__ void __launch_bounds__(256) keygen_kernel(const float* __restrict__ values, const float* __restrict__ minmax,
                                                     int slot, float* conf_out, const GT* __restrict__ gt,
                                                     uint64_t out_mask, const uint8_t* __restrict__ pos_u8, int kind,
                                                     long long seg_len, uint32_t key_base, uint32_t* __restrict__ keys,
                                                     unsigned long long* seg_stats, const KeygenFuse fz,
                                                     const HistArgs hz) {
  __shared__ uint32_t s_h[HIST ? SORT_WARPS : 1][HIST ? MAX_PASSES : 1][HIST ? RADIX : 1];
  if constexpr (HIST) {
    for (int i = threadIdx.x; i < SORT_WARPS * MAX_PASSES * RADIX; i += 256) (&s_h[0][0][0])[i] = 0u;
    __syncthreads();
  }
  const int seg = blockIdx.y;
  const size_t base = (size_t)seg * (size_t)seg_len;
  float lo = 0.f, den = 1.f, mlo = 0.f, mden = 1.f;
  const bool norm = minmax != nullptr;
  if (norm) {
    lo = minmax[seg * 4 + slot * 2];
    den = __fsub_rn(minmax[seg * 4 + slot * 2 + 1], lo);
    mlo = minmax[seg * 4 + 2];
    mden = __fsub_rn(minmax[seg * 4 + 3], mlo);
  }
  const bool fuse = fz.msp != nullptr;
  unsigned n_pos = 0, n_nan = 0, n_oow = 0;
  const long long nvec = seg_len / VEC;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // HIST needs warp-uniform trip counts (ballots): iterate to the block-uniform bound and mask the tail
  const long long q_end = HIST ? ((nvec + stride - 1) / stride) * stride : nvec;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < q_end; q += stride) {
    // a lane past the end re-reads the segment's last vector with every side effect masked, so that the whole warp
    // executes the same top-place ballots
    const bool live = !HIST || q < nvec;
    const size_t i = base + (size_t)(live ? q : nvec - 1) * VEC;
    float v[VEC];
    bool pos[VEC];
    if constexpr (VEC == 4) {
      const float4 t = *reinterpret_cast<const float4*>(values + i);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      if (pos_u8 || sizeof(GT) == 1) {
        const uchar4 g = *reinterpret_cast<const uchar4*>((pos_u8 ? pos_u8 : reinterpret_cast<const uint8_t*>(gt)) + i);
        const unsigned char gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) pos[j] = pos_u8 ? (gg[j] != 0) : (gg[j] < 64 && ((out_mask >> gg[j]) & 1ull));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) pos[j] = is_positive(gt, pos_u8, out_mask, i + j);
      }
    } else {
      v[0] = values[i];
      pos[0] = is_positive(gt, pos_u8, out_mask, i);
    }
    uint32_t key[VEC];
    float mn[VEC], mx[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      if (norm) v[j] = __fdiv_rn(__fsub_rn(v[j], lo), den);  // NumPy: (x - min) / (max - min), fp32
      unsigned c_nan = 0, c_oow = 0;
      key[j] = pack_key(v[j], kind, pos[j], key_base, c_nan, c_oow);
      if (live) { n_pos += pos[j]; n_nan += c_nan; n_oow += c_oow; }
    }
    if (fuse) {
      float m[VEC];
      if constexpr (VEC == 4) {
        const float4 t = *reinterpret_cast<const float4*>(fz.msp + i);
        m[0] = t.x; m[1] = t.y; m[2] = t.z; m[3] = t.w;
      } else {
        m[0] = fz.msp[i];
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        mn[j] = __fdiv_rn(__fsub_rn(m[j], mlo), mden);
        // NumPy: c = 1 / (1 + exp(lamda * (e - thre))); mix = c*e + (1-c)*mmsp   (float32)
        const float c = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(__fmul_rn(fz.lambda, __fsub_rn(v[j], fz.thr)))));
        mx[j] = __fadd_rn(__fmul_rn(c, v[j]), __fmul_rn(__fsub_rn(1.0f, c), mn[j]));
      }
    }
    if constexpr (HIST) {
      const int w = threadIdx.x >> 5;
      const unsigned lt = lanemask_lt_();
      const unsigned vm = ballot_all(live);   // the VEC keys of a thread share it
      // the fused form always counts the full 32-bit key: MAX_PASSES digit places at shifts 0, 8, 16, 24
      // (dml_ood_keygen checks the plan).  The lower places of float keys are close to uniform: plain shared atomics.
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        if (live) {
#pragma unroll
          for (int p = 0; p < MAX_PASSES - 1; ++p) atomicAdd(&s_h[w][p][(key[j] >> (p * RADIX_BITS)) & (RADIX - 1)], 1u);
        }
      }
      // the top place is heavily skewed (a warp usually holds 1-3 distinct values).  Default: shared atomics as well
      // (ATOMS.POPC.INC counts the same-address lanes in one operation).  Alternative: group equal digits with ballots,
      // the group's lowest lane adds the group size with a plain read-modify-write -- bare VOTEs (match8_full /
      // ballot_all): all lanes get here, the trip count is block-uniform.
      if (hz.top_atomic) {
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          if (live) atomicAdd(&s_h[w][MAX_PASSES - 1][key[j] >> ((MAX_PASSES - 1) * RADIX_BITS)], 1u);
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const uint32_t d = key[j] >> ((MAX_PASSES - 1) * RADIX_BITS);
          const unsigned peers = match8_full(d) & vm;
          if (live && (peers & lt) == 0u) s_h[w][MAX_PASSES - 1][d] += (uint32_t)__popc(peers);
          __syncwarp();
        }
      }
    }
    if (!live) continue;   // (after the ballots) nothing to store
    if constexpr (VEC == 4) {
      *reinterpret_cast<uint4*>(keys + i) = make_uint4(key[0], key[1], key[2], key[3]);
      if (conf_out) *reinterpret_cast<float4*>(conf_out + i) = make_float4(v[0], v[1], v[2], v[3]);
      if (fuse && fz.msp_norm) *reinterpret_cast<float4*>(fz.msp_norm + i) = make_float4(mn[0], mn[1], mn[2], mn[3]);
      if (fuse && fz.mix) *reinterpret_cast<float4*>(fz.mix + i) = make_float4(mx[0], mx[1], mx[2], mx[3]);
    } else {
      keys[i] = key[0];
      if (conf_out) conf_out[i] = v[0];
      if (fuse && fz.msp_norm) fz.msp_norm[i] = mn[0];
      if (fuse && fz.mix) fz.mix[i] = mx[0];
    }
  }
  if constexpr (HIST) {
    __syncthreads();
    for (int i = threadIdx.x; i < MAX_PASSES * RADIX; i += 256) {
      const int p = i >> RADIX_BITS, d = i & (RADIX - 1);
      uint32_t c = 0;
#pragma unroll
      for (int ww = 0; ww < SORT_WARPS; ++ww) c += s_h[ww][p][d];
      if (c) atomicAdd(hz.ghist + ((size_t)seg * MAX_PASSES + p) * RADIX + d, c);
    }
  }
  // block reduce the three counters
  __shared__ unsigned s_c[3][8];
  unsigned c[3] = {n_pos, n_nan, n_oow};
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    c[j] = __reduce_add_sync(0xffffffffu, c[j]);
    if ((threadIdx.x & 31) == 0) s_c[j][threadIdx.x >> 5] = c[j];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    unsigned t = 0;
    for (int wv = 0; wv < 8; ++wv) t += s_c[threadIdx.x][wv];
    if (t) atomicAdd(seg_stats + (size_t)seg * 4 + threadIdx.x, (unsigned long long)t);
  }
}

// min / max of the sortable key + NaN / positive counts (for choosing key_base on arbitrary scores)
__global__ void __launch_bounds__(256) keystats_kernel(const float* __restrict__ values, int kind, long long seg_len,
                                                       unsigned long long* seg_stats /*[seg,4]: min, max, n_nan, -*/) {
  const int seg = blockIdx.y;
  const size_t base = (size_t)seg * (size_t)seg_len;
  uint32_t mn = 0xffffffffu, mx = 0u;
  unsigned n_nan = 0;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < seg_len; p += (long long)gridDim.x * blockDim.x) {
    float f = values[base + (size_t)p];
    f = kind == 0 ? f : -f;
    if (f != f) { ++n_nan; continue; }
    if (f == 0.f) f = 0.f;
    const uint32_t u = __float_as_uint(f);
    const uint32_t srt = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    mn = min(mn, srt);
    mx = max(mx, srt);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    n_nan += __shfl_xor_sync(0xffffffffu, n_nan, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(seg_stats + (size_t)seg * 4 + 0, (unsigned long long)mn);
    atomicMax(seg_stats + (size_t)seg * 4 + 1, (unsigned long long)mx);
    if (n_nan) atomicAdd(seg_stats + (size_t)seg * 4 + 2, (unsigned long long)n_nan);
  }
}

__global__ void keystats_init_kernel(unsigned long long* s, int n_seg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_seg) { s[i * 4 + 0] = 0xffffffffull; s[i * 4 + 1] = 0ull; s[i * 4 + 2] = 0ull; s[i * 4 + 3] = 0ull; }
}

// ---------------------------------------------------------------------------------------------
// segmented scan over score groups
// ---------------------------------------------------------------------------------------------
// Agg / Carry / TilePartial and the per-thread mask arithmetic: ood_scan_thread.cuh
__device__ __forceinline__ Agg agg_shfl_up(const Agg& a, int o) {
  Agg r;
  r.pos = __shfl_up_sync(0xffffffffu, a.pos, o);
  r.spos = __shfl_up_sync(0xffffffffu, a.spos, o);
  r.slen = __shfl_up_sync(0xffffffffu, a.slen, o);
  r.head = __shfl_up_sync(0xffffffffu, a.head, o);
  return r;
}

struct RangeInfo {  // device-resident description of one scan range (segment)
  long long pos_before;  // positives ranked before this range
  long long idx_before;  // elements ranked before this range
  long long total_pos;   // P over the whole ranking
  long long total_n;     // P + N over the whole ranking
};

// Loads one tile into registers in BLOCKED order (thread t: elements t*16 .. t*16+15) through a
// padded shared buffer (coalesced global reads, conflict-free strided smem reads), plus the element
// just before / after each thread's run for group-boundary detection.
struct TileKeys {
  uint32_t k[SCAN_ITEMS];
  uint32_t prev, next;  // element before k[0] / after k[15]
  bool has_prev, has_next;
};

__device__ __forceinline__ int pad_idx(int i) { return i + (i >> 4); }

__device__ __forceinline__ void load_tile(const uint32_t* __restrict__ keys, long long n, long long tile_off,
                                          uint32_t* s_keys /*[SCAN_TILE + SCAN_TILE/16 + 2]*/, TileKeys& tk) {
  const int tid = threadIdx.x;
  if (tile_off + SCAN_TILE <= n) {
    // full tile (all but the last tile of a segment): no bounds tests, immediate offsets
    const uint32_t* src = keys + tile_off + tid;
    uint32_t* dst = s_keys + tid + (tid >> 4);
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) dst[j * (SCAN_THREADS + SCAN_THREADS / 16)] = src[j * SCAN_THREADS];
  } else {
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
      const int i = j * SCAN_THREADS + tid;
      const long long g = tile_off + i;
      s_keys[pad_idx(i)] = g < n ? keys[g] : 0u;
    }
  }
  __shared__ uint32_t s_edge[2];
  if (tid == 0) {
    s_edge[0] = tile_off > 0 ? keys[tile_off - 1] : 0u;
    s_edge[1] = tile_off + SCAN_TILE < n ? keys[tile_off + SCAN_TILE] : 0u;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) tk.k[j] = s_keys[pad_idx(tid * SCAN_ITEMS + j)];
  const long long first = tile_off + (long long)tid * SCAN_ITEMS;
  tk.has_prev = first > 0;
  tk.prev = tid > 0 ? s_keys[pad_idx(tid * SCAN_ITEMS - 1)] : s_edge[0];
  tk.has_next = first + SCAN_ITEMS < n;
  tk.next = tid < SCAN_THREADS - 1 ? s_keys[pad_idx((tid + 1) * SCAN_ITEMS)] : s_edge[1];
}

__device__ __forceinline__ Agg thread_aggregate(const TileKeys& tk, long long n, long long first) {
  Agg a = {0u, 0u, 0u, 0u};
  uint32_t prev = tk.prev;
  bool have_prev = tk.has_prev;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    if (first + j < n) {
      const uint32_t k = tk.k[j];
      const bool head = !have_prev || ((k >> 1) != (prev >> 1));
      const unsigned p = k & 1u;
      if (head) { a.spos = 0; a.slen = 0; a.head = 1; }
      a.pos += p; a.spos += p; a.slen += 1;
      prev = k; have_prev = true;
    }
  }
  return a;
}

// phase 1: per-tile aggregate.  BITS: the run's aggregate from its bit masks (ood_scan_thread.cuh); !BITS: the
// original one-step-per-key state machine, kept selectable (DML_SCAN_BITS=0) as the cross-check of the mask form.
template <bool BITS>
__global__ void __launch_bounds__(SCAN_THREADS) scan_agg_kernel(const uint32_t* __restrict__ keys, long long seg_len,
                                                                int tiles_per_seg, Agg* __restrict__ tile_agg) {
  __shared__ uint32_t s_keys[SCAN_TILE + SCAN_TILE / 16 + 2];
  __shared__ Agg s_w[SCAN_WARPS];
  const int seg = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t* k = keys + (size_t)seg * (size_t)seg_len;
  const long long tile_off = (long long)tile * SCAN_TILE;
  TileKeys tk;
  load_tile(k, seg_len, tile_off, s_keys, tk);
  const long long first = tile_off + (long long)tid * SCAN_ITEMS;
  Agg a;
  if constexpr (BITS) a = run_aggregate(run_masks(tk.k, tk.prev, tk.has_prev, tk.next, first, seg_len));
  else a = thread_aggregate(tk, seg_len, first);
  // ordered inclusive warp scan, keep the last lane's value
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    Agg nb = agg_shfl_up(a, o);
    if (lane >= o) a = agg_combine(nb, a);
  }
  if (lane == 31) s_w[w] = a;
  __syncthreads();
  if (tid == 0) {
    Agg t = s_w[0];
    for (int i = 1; i < SCAN_WARPS; ++i) t = agg_combine(t, s_w[i]);
    tile_agg[(size_t)seg * tiles_per_seg + tile] = t;
  }
}

// phase 2: exclusive scan of the tile aggregates of each segment -> carry-in per tile
constexpr int CARRY_THREADS = 1024;
struct CarryH { Carry c; unsigned head; };
__device__ __forceinline__ CarryH carryh_combine(const CarryH& a, const CarryH& b) {
  CarryH r;
  r.c.pos = a.c.pos + b.c.pos;
  r.c.spos = b.head ? b.c.spos : a.c.spos + b.c.spos;
  r.c.slen = b.head ? b.c.slen : a.c.slen + b.c.slen;
  r.head = a.head | b.head;
  return r;
}
// One CTA handles `per_chunk` consecutive tiles of a segment (chunks == 1: the whole segment, the per-image case).
// Long segments (the pooled metric: > 300 000 tiles) are split into chunks so that the scan is not serialised on one
// SM: PHASE 1 writes each chunk's total, scan_chunk_prefix_kernel turns the totals into exclusive chunk prefixes, and
// PHASE 2 redoes the in-chunk scan starting from its prefix.  PHASE 0 = single level (prefix = the range's carry-in).
template <int PHASE>
__global__ void __launch_bounds__(CARRY_THREADS) scan_carry_kernel(const Agg* __restrict__ tile_agg, int tiles_per_seg,
                                                                   int per_chunk, int chunks,
                                                                   const RangeInfo* __restrict__ info,
                                                                   CarryH* __restrict__ chunk_state,
                                                                   Carry* __restrict__ tile_carry) {
  __shared__ CarryH s_t[CARRY_THREADS];
  const int seg = blockIdx.x / chunks, chunk = blockIdx.x - seg * chunks, tid = threadIdx.x;
  const Agg* ag = tile_agg + (size_t)seg * tiles_per_seg;
  Carry* out = tile_carry + (size_t)seg * tiles_per_seg;
  const int t0 = chunk * per_chunk;
  int t1 = t0 + per_chunk;
  if (t1 > tiles_per_seg) t1 = tiles_per_seg;
  const int per = (t1 - t0 + CARRY_THREADS - 1) / CARRY_THREADS;
  int b = t0 + tid * per;
  if (b > t1) b = t1;
  int e = b + per;
  if (e > t1) e = t1;
  CarryH mine;
  mine.c.pos = mine.c.spos = mine.c.slen = 0ull;
  mine.head = 0;
  for (int i = b; i < e; ++i) carry_apply(mine.c, mine.head, ag[i]);
  s_t[tid] = mine;
  __syncthreads();
  // Hillis-Steele inclusive scan over the per-thread aggregates (ordered operator)
  for (int o = 1; o < CARRY_THREADS; o <<= 1) {
    CarryH v = s_t[tid];
    if (tid >= o) v = carryh_combine(s_t[tid - o], v);
    __syncthreads();
    s_t[tid] = v;
    __syncthreads();
  }
  if (PHASE == 1) {
    if (tid == CARRY_THREADS - 1) chunk_state[blockIdx.x] = s_t[tid];
    return;
  }
  CarryH run;
  if (PHASE == 2) {
    run = chunk_state[blockIdx.x];
  } else {
    run.c.pos = (unsigned long long)info[seg].pos_before;  // the range starts on a group boundary
    run.c.spos = run.c.slen = 0ull;
    run.head = 0;
  }
  if (tid > 0) run = carryh_combine(run, s_t[tid - 1]);
  for (int i = b; i < e; ++i) {
    out[i] = run.c;
    carry_apply(run.c, run.head, ag[i]);
  }
}

// chunk totals -> exclusive chunk prefixes (in place), one CTA per segment, chunks <= CARRY_THREADS
__global__ void __launch_bounds__(CARRY_THREADS) scan_chunk_prefix_kernel(CarryH* __restrict__ chunk_state, int chunks,
                                                                          const RangeInfo* __restrict__ info) {
  __shared__ CarryH s_t[CARRY_THREADS];
  const int seg = blockIdx.x, tid = threadIdx.x;
  CarryH mine;
  mine.c.pos = mine.c.spos = mine.c.slen = 0ull;
  mine.head = 0;
  if (tid < chunks) mine = chunk_state[(size_t)seg * chunks + tid];
  s_t[tid] = mine;
  __syncthreads();
  for (int o = 1; o < CARRY_THREADS; o <<= 1) {
    CarryH v = s_t[tid];
    if (tid >= o) v = carryh_combine(s_t[tid - o], v);
    __syncthreads();
    s_t[tid] = v;
    __syncthreads();
  }
  CarryH run;
  run.c.pos = (unsigned long long)info[seg].pos_before;
  run.c.spos = run.c.slen = 0ull;
  run.head = 0;
  if (tid > 0) run = carryh_combine(run, s_t[tid - 1]);
  if (tid < chunks) chunk_state[(size_t)seg * chunks + tid] = run;
}

// phase 3: per-tile group contributions (BITS: see scan_agg_kernel)
template <bool BITS>
__global__ void __launch_bounds__(SCAN_THREADS, BITS ? 4 : 3) scan_apply_kernel(const uint32_t* __restrict__ keys, long long seg_len,
                                                                  int tiles_per_seg, const Carry* __restrict__ tile_carry,
                                                                  const RangeInfo* __restrict__ info, double recall_level,
                                                                  TilePartial* __restrict__ partials) {
  __shared__ uint32_t s_keys[SCAN_TILE + SCAN_TILE / 16 + 2];
  __shared__ Agg s_w[SCAN_WARPS];
  __shared__ TilePartial s_p[SCAN_WARPS];
  __shared__ long long s_tstar;
  const int seg = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t* k = keys + (size_t)seg * (size_t)seg_len;
  const long long tile_off = (long long)tile * SCAN_TILE;
  const long long first = tile_off + (long long)tid * SCAN_ITEMS;
  const RangeInfo ri = info[seg];
  if (tid == 0) s_tstar = recall_threshold(ri.total_pos, recall_level);
  TileKeys tk;
  load_tile(k, seg_len, tile_off, s_keys, tk);   // contains the __syncthreads that publishes s_tstar
  const long long tstar = s_tstar;
  RunMasks rm = {};
  Agg mine;
  if constexpr (BITS) {
    rm = run_masks(tk.k, tk.prev, tk.has_prev, tk.next, first, seg_len);
    mine = run_aggregate(rm);
  } else {
    mine = thread_aggregate(tk, seg_len, first);
  }
  // exclusive block scan of the thread aggregates
  Agg incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    Agg nb = agg_shfl_up(incl, o);
    if (lane >= o) incl = agg_combine(nb, incl);
  }
  if (lane == 31) s_w[w] = incl;
  __syncthreads();
  Agg excl = agg_shfl_up(incl, 1);
  if (lane == 0) excl = Agg{0u, 0u, 0u, 0u};
  Agg wpre = {0u, 0u, 0u, 0u};
  for (int i = 0; i < w; ++i) wpre = agg_combine(wpre, s_w[i]);
  excl = agg_combine(wpre, excl);

  // Everything inside the tile is 32-bit and relative to this thread's first element; 64-bit values are
  // formed once per thread (and at the rare group ends that carry an open group in from earlier tiles).
  const Carry tc = tile_carry[(size_t)seg * tiles_per_seg + tile];
  const long long base_P = (long long)(tc.pos + excl.pos);                      // positives ranked before my run
  const long long open_pos = (long long)(excl.head ? excl.spos : tc.spos + excl.spos);   // open group before my run
  const long long open_len = (long long)(excl.head ? excl.slen : tc.slen + excl.slen);
  const long long Ptot = ri.total_pos;
  const long long first_idx = ri.idx_before + first;
  // thresholds in local terms, clamped to [-1, SCAN_ITEMS + 1]
  auto clamp_local = [](long long v) { return (int)(v < -1 ? -1 : (v > SCAN_ITEMS + 1 ? SCAN_ITEMS + 1 : v)); };
  const int t_local = clamp_local(tstar - base_P);            // tps <= T*        <=>  Pl <= t_local
  const int rem_local = clamp_local(Ptot - base_P);           // tps_{g-1} < Ptot <=>  Pl_start < rem_local
  const bool open_valid = (base_P - open_pos) < Ptot;         // same test for the group carried in

  unsigned long long auroc;
  double ap;
  int a_j = -1, a_Pl = 0, b_j = -1, b_Pl = 0, ngroups = 0;
  if constexpr (BITS) {
    const RunContribution rc = run_contribution(rm, base_P, open_pos, open_len, Ptot, first_idx, t_local, rem_local);
    auroc = rc.auroc; ap = rc.ap_sum; ngroups = rc.n_groups;
    a_j = rc.a_j; a_Pl = rc.a_Pl; b_j = rc.b_j; b_Pl = rc.b_Pl;
  } else {
  TilePartial acc;
  partial_init(acc);
  unsigned long long neg_P = 0ull, tie = 0ull;   // sum over negatives of Pl ; sum over mixed groups of neg_g * pos_g
  int n_neg = 0;
  int Pl = 0, ls = 0, ll = 0, Pl_start = 0;
  bool in_thread = false;                        // current group started inside my run

  bool head = !tk.has_prev || ((tk.k[0] >> 1) != (tk.prev >> 1));
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    const long long gi = first + j;
    if (gi < seg_len) {
      const uint32_t key = tk.k[j];
      const int p = (int)(key & 1u);
      if (head) { ls = 0; ll = 0; Pl_start = Pl; in_thread = true; }
      if (!p) { neg_P += (unsigned)Pl; ++n_neg; }
      Pl += p; ls += p; ll += 1;
      // group end: the next element (if any) opens a new group
      bool endg;
      if (gi + 1 >= seg_len) endg = true;
      else {
        const uint32_t nk = (j + 1 < SCAN_ITEMS) ? tk.k[(j + 1) % SCAN_ITEMS] : tk.next;
        endg = (nk >> 1) != (key >> 1);
      }
      if (endg) {
        const long long pos_g = in_thread ? (long long)ls : open_pos + ls;
        const long long len_g = in_thread ? (long long)ll : open_len + ll;
        if (pos_g) {
          const long long neg_g = len_g - pos_g;
          if (neg_g) tie += (unsigned long long)(neg_g * pos_g);
          acc.ap_sum += (double)pos_g * ((double)(base_P + Pl) / (double)(first_idx + j + 1));
        }
        const bool valid = in_thread ? (Pl_start < rem_local) : open_valid;   // groups up to the first with full recall
        if (valid) {
          if (Pl <= t_local) { a_j = j; a_Pl = Pl; }
          else if (b_j < 0 || Pl == b_Pl) { b_j = j; b_Pl = Pl; }
        }
        ++ngroups;
      }
      head = endg;
    }
  }
  // ---- block reduction.  Sums: fixed-order shuffles.  FPR candidates: the winning thread is found with one
  //      32-bit warp reduction on a tile-local key, only the winner materialises the 64-bit tuple. ----------
  auroc = 2ull * ((unsigned long long)n_neg * (unsigned long long)base_P + neg_P) + tie;
  ap = acc.ap_sum;
  }  // !BITS
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    auroc += __shfl_down_sync(0xffffffffu, auroc, o);
    ap += __shfl_down_sync(0xffffffffu, ap, o);
  }
  const int ng_w = __reduce_add_sync(0xffffffffu, ngroups);
  // a: the latest group end with tps <= T*  -> max tile-local element index
  const int a_key = a_j >= 0 ? tid * SCAN_ITEMS + a_j : -1;
  const int a_best = __reduce_max_sync(0xffffffffu, a_key);
  // b: smallest tps > T*, then the latest index.  tps grows with the index, so order by the positives counted
  //    inside the tile (excl.pos + Pl <= SCAN_TILE) and, among equals, by the reversed index.
  const unsigned b_key = b_j >= 0 ? (((unsigned)(excl.pos + (unsigned)b_Pl)) << 13) | (unsigned)(SCAN_TILE - 1 - (tid * SCAN_ITEMS + b_j))
                                  : 0xffffffffu;
  const unsigned b_best = __reduce_min_sync(0xffffffffu, b_key);
  if (lane == 0) {
    TilePartial t;
    partial_init(t);
    t.auroc_num = auroc; t.ap_sum = ap; t.n_groups = ng_w;
    s_p[w] = t;
  }
  __syncwarp();
  if (a_j >= 0 && a_key == a_best) {
    s_p[w].a_idx = first_idx + a_j; s_p[w].a_tps = base_P + a_Pl; s_p[w].a_fps = first_idx + a_j + 1 - (base_P + a_Pl);
  }
  if (b_j >= 0 && b_key == b_best) {
    s_p[w].b_idx = first_idx + b_j; s_p[w].b_tps = base_P + b_Pl; s_p[w].b_fps = first_idx + b_j + 1 - (base_P + b_Pl);
  }
  __syncthreads();
  if (tid == 0) {
    TilePartial t = s_p[0];
    for (int i = 1; i < SCAN_WARPS; ++i) partial_merge(t, s_p[i]);
    partials[(size_t)seg * tiles_per_seg + tile] = t;
  }
}

// phase 4: fixed-order reduction of the tile partials of each segment
constexpr int FIN_THREADS = 256;
// `chunks` > 1: block (seg, chunk) merges tiles [chunk*per_chunk, ...) of its segment into range_partials[seg*chunks + chunk]
// (first level of the two-level reduction of long segments); chunks == 1: whole segment.
__global__ void __launch_bounds__(FIN_THREADS) scan_finalize_kernel(const TilePartial* __restrict__ partials, int tiles_per_seg,
                                                                    int per_chunk, int chunks,
                                                                    const RangeInfo* __restrict__ info,
                                                                    const unsigned long long* __restrict__ seg_stats,
                                                                    double recall_level, dml_ood_result* __restrict__ results,
                                                                    TilePartial* __restrict__ range_partials) {
  __shared__ TilePartial s_t[FIN_THREADS];
  const int seg = blockIdx.x / chunks, chunk = blockIdx.x - seg * chunks, tid = threadIdx.x;
  const TilePartial* pp = partials + (size_t)seg * tiles_per_seg;
  TilePartial t;
  partial_init(t);
  // contiguous chunk per thread => the same summation order regardless of scheduling
  const int t0 = chunk * per_chunk;
  int t1 = t0 + per_chunk;
  if (t1 > tiles_per_seg) t1 = tiles_per_seg;
  const int per = (t1 - t0 + FIN_THREADS - 1) / FIN_THREADS;
  int b = t0 + tid * per;
  if (b > t1) b = t1;
  int e = b + per;
  if (e > t1) e = t1;
  for (int i = b; i < e; ++i) partial_merge(t, pp[i]);
  s_t[tid] = t;
  __syncthreads();
  for (int o = FIN_THREADS / 2; o > 0; o >>= 1) {
    if (tid < o) {
      TilePartial a = s_t[tid];
      partial_merge(a, s_t[tid + o]);
      s_t[tid] = a;
    }
    __syncthreads();
  }
  if (tid == 0) {
    const TilePartial r = s_t[0];
    if (range_partials) range_partials[blockIdx.x] = r;
    if (results) {
      const RangeInfo ri = info[seg];
      const double P = (double)ri.total_pos, N = (double)(ri.total_n - ri.total_pos);
      dml_ood_result o;
      o.n_pos = ri.total_pos;
      o.n_neg = ri.total_n - ri.total_pos;
      o.n_nan = seg_stats ? (long long)seg_stats[(size_t)seg * 4 + 1] : 0;
      o.n_groups = r.n_groups;
      if (ri.total_pos > 0 && o.n_neg > 0) {
        o.auroc = (double)r.auroc_num / (2.0 * P * N);
        o.aupr = r.ap_sum / P;
        // |recall - level| of the two candidates in float64, ties -> the later group (b)
        const double inf = __longlong_as_double(0x7ff0000000000000ll);
        const double da = r.a_idx >= 0 ? fabs((double)r.a_tps / P - recall_level) : inf;
        const double db = r.b_tps != NO_B ? fabs((double)r.b_tps / P - recall_level) : inf;
        o.fpr = (double)(db <= da ? r.b_fps : r.a_fps) / N;
      } else {
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        o.auroc = o.aupr = o.fpr = nan;
      }
      results[seg] = o;
    }
  }
}

// first index i with sorted[i] >= q (one thread per query; used to cut sorted shards at the splitters)
__global__ void lower_bound_kernel(const uint32_t* __restrict__ sorted, long long n, const uint32_t* __restrict__ queries,
                                   int nq, long long* __restrict__ pos) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq) return;
  const uint32_t q = queries[t];
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = lo + ((hi - lo) >> 1);
    if (sorted[mid] < q) lo = mid + 1; else hi = mid;
  }
  pos[t] = lo;
}

// positives (bit 0) in a key range
__global__ void __launch_bounds__(256) count_pos_kernel(const uint32_t* __restrict__ keys, long long n, unsigned long long* out) {
  unsigned c = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) c += keys[i] & 1u;
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

// RangeInfo of whole segments: everything starts at zero, totals from the key-gen statistics
__global__ void range_info_from_stats_kernel(const unsigned long long* __restrict__ seg_stats, long long seg_len, int n_seg,
                                             RangeInfo* __restrict__ info) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_seg) {
    RangeInfo r;
    r.pos_before = 0; r.idx_before = 0;
    r.total_pos = (long long)seg_stats[(size_t)i * 4 + 0];
    r.total_n = seg_len;
    info[i] = r;
  }
}

// ---------------------------------------------------------------------------------------------
// Second FPR@recall convention: fpr[tpr >= recall][0] on sklearn.metrics.roc_curve(drop_intermediate=True)
// (the softmax-baseline evaluator, DeepLabV3Plus-Pytorch/test.py:241-244).  Runs after the scan on the same
// workspace: the per-tile carries locate the tile that holds the T-th positive (T = the smallest count whose
// float64 recall T/P reaches the level), one block finds the element, and warp 0 then walks the score groups:
// the ROC point of a group is dropped by roc_curve iff the NEXT group has the same (positives, negatives)
// counts (both second differences vanish), so the answer is the false-positive count at the end of the run of
// equal-count groups that starts at the group holding the T-th positive (the first and last points always stay).
// ---------------------------------------------------------------------------------------------
struct GroupSpan { long long end; unsigned long long pos; };   // last index of the group, positives in [from, end]

// warp-cooperative: extends the group of `score` (key >> 1) from index `from` (whose score is `score`) to its end
__device__ __forceinline__ GroupSpan warp_group_end(const uint32_t* __restrict__ k, long long n, long long from, uint32_t score) {
  const int lane = threadIdx.x & 31;
  GroupSpan g;
  g.end = from - 1;
  g.pos = 0ull;
  for (long long base = from; base < n; base += 32) {
    const long long i = base + lane;
    const uint32_t key = i < n ? k[i] : 0u;
    const bool same = i < n && (key >> 1) == score;
    const unsigned m = __ballot_sync(0xffffffffu, same);
    const int run = m == 0xffffffffu ? 32 : __ffs(~m) - 1;             // leading lanes still inside the group
    const unsigned in_run = run == 32 ? 0xffffffffu : ((1u << run) - 1u);
    g.pos += __popc(__ballot_sync(0xffffffffu, same && (key & 1u)) & in_run);
    g.end += run;
    if (run < 32) break;
  }
  return g;
}
// ... and backwards: first index of the group that contains `from`, positives in [start, from)
__device__ __forceinline__ GroupSpan warp_group_start(const uint32_t* __restrict__ k, long long from, uint32_t score) {
  const int lane = threadIdx.x & 31;
  GroupSpan g;
  g.end = from;   // (re-used as "start")
  g.pos = 0ull;
  for (long long base = from - 1; base >= 0; base -= 32) {
    const long long i = base - lane;
    const uint32_t key = i >= 0 ? k[i] : 0u;
    const bool same = i >= 0 && (key >> 1) == score;
    const unsigned m = __ballot_sync(0xffffffffu, same);
    const int run = m == 0xffffffffu ? 32 : __ffs(~m) - 1;
    const unsigned in_run = run == 32 ? 0xffffffffu : ((1u << run) - 1u);
    g.pos += __popc(__ballot_sync(0xffffffffu, same && (key & 1u)) & in_run);
    g.end -= run;
    if (run < 32) break;
  }
  return g;
}

__global__ void __launch_bounds__(SCAN_THREADS) roc_fpr_kernel(const uint32_t* __restrict__ keys, long long seg_len,
                                                               int tiles_per_seg, const Carry* __restrict__ tile_carry,
                                                               const unsigned long long* __restrict__ seg_stats,
                                                               double recall_level, double* __restrict__ out) {
  __shared__ unsigned s_warp[SCAN_WARPS];
  __shared__ long long s_e0;
  const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t* k = keys + (size_t)seg * (size_t)seg_len;
  const long long P = (long long)seg_stats[(size_t)seg * 4 + 0];
  const long long N = seg_len - P;
  if (P <= 0 || N <= 0) {
    if (tid == 0) out[seg] = __longlong_as_double(0x7ff8000000000000ll);
    return;
  }
  // T = smallest t in [1, P] with (double)t / P >= recall_level
  long long T;
  {
    const double dP = (double)P;
    double g = ceil(recall_level * dP);
    T = g < 1.0 ? 1 : (g > dP ? P : (long long)g);
    while (T > 1 && (double)(T - 1) / dP >= recall_level) --T;
    while (T < P && (double)T / dP < recall_level) ++T;
  }
  // tile holding the T-th positive: the last tile whose carry (positives before it) is < T
  const Carry* carry = tile_carry + (size_t)seg * tiles_per_seg;
  int lo = 0, hi = tiles_per_seg - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if ((long long)carry[mid].pos < T) lo = mid; else hi = mid - 1;
  }
  const long long tile_off = (long long)lo * SCAN_TILE;
  const long long need = T - (long long)carry[lo].pos;            // 1-based rank of the positive inside the tile
  // blocked positives count per thread, block-wide inclusive scan
  const long long first = tile_off + (long long)tid * SCAN_ITEMS;
  unsigned cnt = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) cnt += (first + j < seg_len) ? (k[first + j] & 1u) : 0u;
  unsigned incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned nb = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += nb;
  }
  if (lane == 31) s_warp[w] = incl;
  __syncthreads();
  unsigned before = incl - cnt;
  for (int i = 0; i < w; ++i) before += s_warp[i];
  if ((long long)before < need && need <= (long long)(before + cnt)) {
    unsigned c = before;
    for (int j = 0; j < SCAN_ITEMS; ++j) {
      if (first + j < seg_len && (k[first + j] & 1u)) {
        if ((long long)(++c) == need) { s_e0 = first + j; break; }
      }
    }
  }
  __syncthreads();
  if (w != 0) return;
  // ---- warp 0: the group of the T-th positive, then the run of equal-count groups -------------------
  const long long e0 = s_e0;
  const uint32_t score0 = k[e0] >> 1;
  const GroupSpan fwd = warp_group_end(k, seg_len, e0, score0);              // positives in [e0, end]
  const GroupSpan bwd = warp_group_start(k, e0, score0);                      // positives in [start, e0)
  long long end = fwd.end;
  long long tps = (T - 1) + (long long)fwd.pos;                                // cumulative positives at the group's end
  long long g_pos = (long long)(fwd.pos + bwd.pos);
  long long g_neg = (end - bwd.end + 1) - g_pos;
  const bool is_first_group = bwd.end == 0;
  if (!is_first_group) {
    while (end + 1 < seg_len) {
      const uint32_t sc = k[end + 1] >> 1;
      const GroupSpan nx = warp_group_end(k, seg_len, end + 1, sc);
      const long long n_pos = (long long)nx.pos, n_neg = (nx.end - end) - n_pos;
      if (n_pos != g_pos || n_neg != g_neg) break;                            // this point has a non-zero second difference: kept
      end = nx.end;                                                            // collinear: dropped, move to the next point
      tps += n_pos;
    }
  }
  if (lane == 0) out[seg] = (double)((end + 1) - tps) / (double)N;
}

constexpr int SCAN_CHUNK_TILES = 2048;       // tiles per chunk of the two-level carry / finalize (long segments only)
constexpr int SCAN_CHUNK_MIN_TILES = 8192;   // segments shorter than this stay single-level
struct MetricsPlan {
  SortPlan sort;
  int tiles_per_seg;  // scan tiles
  int chunks;         // > 1: two-level carry / finalize
  size_t off_agg, off_carry, off_partial, off_info, off_chunk_state, off_chunk_partial, off_end;
};

MetricsPlan make_metrics_plan(int n_seg, long long seg_len) {
  MetricsPlan m;
  m.sort = make_sort_plan(n_seg, seg_len, 0, 32);
  m.tiles_per_seg = (int)((seg_len + SCAN_TILE - 1) / SCAN_TILE);
  if (m.tiles_per_seg < 1) m.tiles_per_seg = 1;
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t nt = (size_t)n_seg * m.tiles_per_seg;
  m.off_agg = m.sort.off_end;
  m.off_carry = align(m.off_agg + nt * sizeof(Agg));
  m.off_partial = align(m.off_carry + nt * sizeof(Carry));
  m.off_info = align(m.off_partial + nt * sizeof(TilePartial));
  m.chunks = 1;
  if (m.tiles_per_seg >= SCAN_CHUNK_MIN_TILES) {
    m.chunks = (m.tiles_per_seg + SCAN_CHUNK_TILES - 1) / SCAN_CHUNK_TILES;
    if (m.chunks > CARRY_THREADS) m.chunks = 1;   // > 8.6 G keys per segment: not reachable (seg_len < 2^32)
  }
  m.off_chunk_state = align(m.off_info + (size_t)n_seg * sizeof(RangeInfo));
  m.off_chunk_partial = align(m.off_chunk_state + (size_t)n_seg * m.chunks * sizeof(CarryH));
  m.off_end = align(m.off_chunk_partial + (size_t)n_seg * m.chunks * sizeof(TilePartial));
  return m;
}

int run_scan(const uint32_t* sorted, const MetricsPlan& m, unsigned char* ws, const RangeInfo* info, double recall_level,
             const unsigned long long* seg_stats, dml_ood_result* results, TilePartial* range_partials, cudaStream_t stream) {
  Agg* agg = reinterpret_cast<Agg*>(ws + m.off_agg);
  Carry* carry = reinterpret_cast<Carry*>(ws + m.off_carry);
  TilePartial* partial = reinterpret_cast<TilePartial*>(ws + m.off_partial);
  dim3 grid((unsigned)m.tiles_per_seg, (unsigned)m.sort.n_seg);
  // DML_SCAN_BITS=0 selects the original per-key state machine (cross-check / A-B timing); read per call, no state kept
  const char* sb = getenv("DML_SCAN_BITS");
  const bool bits = !(sb && sb[0] == '0');
  if (bits) scan_agg_kernel<true><<<grid, SCAN_THREADS, 0, stream>>>(sorted, m.sort.seg_len, m.tiles_per_seg, agg);
  else scan_agg_kernel<false><<<grid, SCAN_THREADS, 0, stream>>>(sorted, m.sort.seg_len, m.tiles_per_seg, agg);
  DML_LAUNCH_CHECK();
  CarryH* chunk_state = reinterpret_cast<CarryH*>(ws + m.off_chunk_state);
  TilePartial* chunk_partial = reinterpret_cast<TilePartial*>(ws + m.off_chunk_partial);
  if (m.chunks > 1) {
    scan_carry_kernel<1><<<m.sort.n_seg * m.chunks, CARRY_THREADS, 0, stream>>>(agg, m.tiles_per_seg, SCAN_CHUNK_TILES, m.chunks, info, chunk_state, carry);
    DML_LAUNCH_CHECK();
    scan_chunk_prefix_kernel<<<m.sort.n_seg, CARRY_THREADS, 0, stream>>>(chunk_state, m.chunks, info);
    DML_LAUNCH_CHECK();
    scan_carry_kernel<2><<<m.sort.n_seg * m.chunks, CARRY_THREADS, 0, stream>>>(agg, m.tiles_per_seg, SCAN_CHUNK_TILES, m.chunks, info, chunk_state, carry);
    DML_LAUNCH_CHECK();
  } else {
    scan_carry_kernel<0><<<m.sort.n_seg, CARRY_THREADS, 0, stream>>>(agg, m.tiles_per_seg, m.tiles_per_seg, 1, info, chunk_state, carry);
    DML_LAUNCH_CHECK();
  }
  if (bits) scan_apply_kernel<true><<<grid, SCAN_THREADS, 0, stream>>>(sorted, m.sort.seg_len, m.tiles_per_seg, carry, info, recall_level, partial);
  else scan_apply_kernel<false><<<grid, SCAN_THREADS, 0, stream>>>(sorted, m.sort.seg_len, m.tiles_per_seg, carry, info, recall_level, partial);
  DML_LAUNCH_CHECK();
  if (m.chunks > 1) {
    // level 1: chunk partials (fixed order inside a chunk); level 2: the chunk partials of each segment
    scan_finalize_kernel<<<m.sort.n_seg * m.chunks, FIN_THREADS, 0, stream>>>(partial, m.tiles_per_seg, SCAN_CHUNK_TILES, m.chunks, info, nullptr, recall_level, nullptr, chunk_partial);
    DML_LAUNCH_CHECK();
    scan_finalize_kernel<<<m.sort.n_seg, FIN_THREADS, 0, stream>>>(chunk_partial, m.chunks, m.chunks, 1, info, seg_stats, recall_level, results, range_partials);
    DML_LAUNCH_CHECK();
  } else {
    scan_finalize_kernel<<<m.sort.n_seg, FIN_THREADS, 0, stream>>>(partial, m.tiles_per_seg, m.tiles_per_seg, 1, info, seg_stats, recall_level, results, range_partials);
    DML_LAUNCH_CHECK();
  }
  return DML_OK;
}

// blocks 0..MAX_PASSES-1: pooled[p][d] (+)= sum over segments of seg_hist[seg][p][d] (one thread per (pass, digit),
// coalesced over digits); block MAX_PASSES: pooled_stats[c] (+)= sum over segments of seg_stats[seg][c]
__global__ void __launch_bounds__(RADIX) pool_hist_kernel(const uint32_t* __restrict__ seg_hist, int n_seg,
                                                          uint32_t* __restrict__ pooled, int reset,
                                                          const unsigned long long* __restrict__ seg_stats,
                                                          unsigned long long* __restrict__ pooled_stats) {
  if (blockIdx.x == MAX_PASSES) {
    if (!seg_stats || !pooled_stats) return;
    __shared__ unsigned long long s_s[4][RADIX];
    unsigned long long a[4] = {0ull, 0ull, 0ull, 0ull};
    for (int sg = threadIdx.x; sg < n_seg; sg += RADIX)
#pragma unroll
      for (int c = 0; c < 4; ++c) a[c] += seg_stats[(size_t)sg * 4 + c];
#pragma unroll
    for (int c = 0; c < 4; ++c) s_s[c][threadIdx.x] = a[c];
    __syncthreads();
    for (int o = RADIX / 2; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o)
#pragma unroll
        for (int c = 0; c < 4; ++c) s_s[c][threadIdx.x] += s_s[c][threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x < 4) pooled_stats[threadIdx.x] = (reset ? 0ull : pooled_stats[threadIdx.x]) + s_s[threadIdx.x][0];
    return;
  }
  if (!seg_hist || !pooled) return;
  const int i = blockIdx.x * RADIX + threadIdx.x;   // p * RADIX + d
  uint32_t c = reset ? 0u : pooled[i];
  for (int s = 0; s < n_seg; ++s) c += seg_hist[(size_t)s * MAX_PASSES * RADIX + i];
  pooled[i] = c;
}
}  // namespace
}  // namespace dml

using namespace dml;

extern "C" {
#pragma GCC visibility push(default)

int dml_ood_keystats(const float* values, int32_t score_kind, int32_t n_seg, int64_t seg_len, long long* seg_stats,
                     dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!values || !seg_stats || n_seg < 0 || seg_len < 0 || n_seg > 65535) return DML_ERR_INVALID_ARG;
  if (n_seg == 0) return DML_OK;
  keystats_init_kernel<<<ceil_div_i(n_seg, 256), 256, 0, stream>>>((unsigned long long*)seg_stats, n_seg);
  DML_LAUNCH_CHECK();
  if (seg_len == 0) return DML_OK;
  long long bx = (seg_len + 256 * 16 - 1) / (256 * 16);
  const long long cap = n_seg >= 148 * 8 ? 8 : (148 * 16) / n_seg + 1;
  if (bx > cap) bx = cap;
  keystats_kernel<<<dim3((unsigned)bx, (unsigned)n_seg), 256, 0, stream>>>(values, score_kind, seg_len,
                                                                            (unsigned long long*)seg_stats);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

int dml_ood_keygen(const float* values, const float* minmax, int32_t minmax_slot, float* conf_out, const uint8_t* gt_u8,
                   const int64_t* gt_i64, uint64_t out_label_mask, const uint8_t* pos_u8, int32_t score_kind,
                   uint32_t key_base, int32_t n_seg, int64_t seg_len, uint32_t* keys, long long* seg_stats,
                   const float* msp, float* msp_norm_out, float* mix_out, float lambda, float thr, void* sort_workspace,
                   size_t sort_workspace_bytes, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!values || !keys || !seg_stats || n_seg < 0 || seg_len < 0 || n_seg > 65535) return DML_ERR_INVALID_ARG;
  const int nsrc = (gt_u8 != nullptr) + (gt_i64 != nullptr) + (pos_u8 != nullptr);
  if (nsrc != 1) return DML_ERR_INVALID_ARG;
  if (minmax && (minmax_slot < 0 || minmax_slot > 1)) return DML_ERR_INVALID_ARG;
  if (score_kind != 0 && score_kind != 1) return DML_ERR_INVALID_ARG;
  if ((msp_norm_out || mix_out) && (!msp || !minmax || minmax_slot != 0)) return DML_ERR_INVALID_ARG;
  if (n_seg == 0) return DML_OK;
  DML_CUDA_TRY(cudaMemsetAsync(seg_stats, 0, (size_t)n_seg * 4 * sizeof(long long), stream));
  if (seg_len == 0) return DML_OK;
  KeygenFuse fz;
  fz.msp = (msp_norm_out || mix_out) ? msp : nullptr;
  fz.msp_norm = msp_norm_out; fz.mix = mix_out; fz.lambda = lambda; fz.thr = thr;
  auto al = [](const void* q, size_t a) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % a) == 0; };
  const bool vec4 = (seg_len % 4 == 0) && al(values, 16) && al(keys, 16) && al(conf_out, 16) && al(fz.msp, 16) &&
                    al(msp_norm_out, 16) && al(mix_out, 16) && al(gt_u8, 4) && al(pos_u8, 4);
  const long long per_thread = vec4 ? 4 : 1;
  long long bx = (seg_len / per_thread + 256 * 2 - 1) / (256 * 2);
  const long long cap = n_seg >= 148 * 8 ? 16 : (148 * 32) / n_seg + 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)n_seg);
  unsigned long long* st = (unsigned long long*)seg_stats;
  const long long* g64 = (const long long*)gt_i64;
  HistArgs hz = {};
  const bool hist = sort_workspace != nullptr;
  if (hist) {
    // fused digit histograms for the full-key sort dml_ood_eval_segments(..., hist_precomputed = 1) will run
    const SortPlan plan = make_sort_plan(n_seg, seg_len, 0, 32);
    if (sort_workspace_bytes < plan.off_end) return DML_ERR_WORKSPACE;
    hz.ghist = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(sort_workspace) + plan.off_hist);
    hz.n_passes = plan.n_passes;
    if (plan.n_passes != MAX_PASSES) return DML_ERR_INVALID_ARG;   // the kernel hard-codes the 4 x 8-bit places
    for (int p = 0; p < MAX_PASSES; ++p) hz.shifts[p] = plan.shifts[p];
    // measured on B200 (profiles/r1e_*): ATOMS.POPC.INC merges same-address lanes in hardware, so even the heavily
    // skewed top place is cheaper with shared atomics (key-gen 266 us / 46 M keys) than with ballot groups (306 us);
    // DML_KEYGEN_TOP_ATOMIC=0 selects the ballot form (A-B knob, read per call)
    const char* ta = getenv("DML_KEYGEN_TOP_ATOMIC");
    hz.top_atomic = (ta && ta[0] == '0') ? 0 : 1;
    DML_CUDA_TRY(cudaMemsetAsync(hz.ghist, 0, plan.off_lookback - plan.off_hist, stream));
  }
#define DML_KEYGEN_LAUNCH(GT, V, gtp, posp)                                                                          \
  do {                                                                                                                 \
    if (hist) keygen_kernel<GT, V, true><<<grid, 256, 0, stream>>>(values, minmax, minmax_slot, conf_out, gtp, out_label_mask, posp, score_kind, seg_len, key_base, keys, st, fz, hz); \
    else keygen_kernel<GT, V, false><<<grid, 256, 0, stream>>>(values, minmax, minmax_slot, conf_out, gtp, out_label_mask, posp, score_kind, seg_len, key_base, keys, st, fz, hz); \
  } while (0)
  if (gt_i64) {
    if (vec4) DML_KEYGEN_LAUNCH(long long, 4, g64, nullptr);
    else DML_KEYGEN_LAUNCH(long long, 1, g64, nullptr);
  } else {
    if (vec4) DML_KEYGEN_LAUNCH(uint8_t, 4, gt_u8, pos_u8);
    else DML_KEYGEN_LAUNCH(uint8_t, 1, gt_u8, pos_u8);
  }
#undef DML_KEYGEN_LAUNCH
  DML_LAUNCH_CHECK();
  return DML_OK;
}

size_t dml_ood_workspace_bytes(int32_t n_seg, int64_t seg_len) {
  if (n_seg <= 0 || seg_len <= 0) return 256;
  return make_metrics_plan(n_seg, seg_len).off_end;
}

int dml_ood_eval_segments(uint32_t* keys, const long long* seg_stats, int32_t n_seg, int64_t seg_len, double recall_level,
                          void* workspace, size_t workspace_bytes, int32_t hist_precomputed, dml_ood_result* results,
                          dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!keys || !seg_stats || !workspace || !results || n_seg < 0 || seg_len < 0 || n_seg > 65535) return DML_ERR_INVALID_ARG;
  if (seg_len >= (1ll << 32)) return DML_ERR_INVALID_ARG;
  if (n_seg == 0) return DML_OK;
  const MetricsPlan m = make_metrics_plan(n_seg, seg_len);
  if (workspace_bytes < m.off_end) return DML_ERR_WORKSPACE;
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  uint32_t* sorted = keys;
  if (seg_len > 0) {
    int rc = radix_sort_segments(keys, m.sort, workspace, &sorted, stream, hist_precomputed != 0);
    if (rc != DML_OK) return rc;
  }
  RangeInfo* info = reinterpret_cast<RangeInfo*>(ws + m.off_info);
  range_info_from_stats_kernel<<<ceil_div_i(n_seg, 256), 256, 0, stream>>>((const unsigned long long*)seg_stats, seg_len, n_seg, info);
  DML_LAUNCH_CHECK();
  return run_scan(sorted, m, ws, info, recall_level, (const unsigned long long*)seg_stats, results, nullptr, stream);
}

int dml_ood_pool_histograms(const void* seg_workspace, size_t seg_workspace_bytes, const long long* seg_stats, int32_t n_seg,
                            int64_t seg_len, void* pooled_workspace, size_t pooled_workspace_bytes, int64_t pooled_len,
                            long long* pooled_stats, int32_t reset, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_seg < 0 || n_seg > 65535 || seg_len < 0 || pooled_len < 0) return DML_ERR_INVALID_ARG;
  if (seg_len >= (1ll << 32) || pooled_len >= (1ll << 32)) return DML_ERR_INVALID_ARG;
  if ((pooled_workspace && !seg_workspace) || (pooled_stats && !seg_stats)) return DML_ERR_INVALID_ARG;
  const uint32_t* sh = nullptr;
  uint32_t* ph = nullptr;
  if (pooled_workspace) {
    const SortPlan seg_plan = make_sort_plan(n_seg > 0 ? n_seg : 1, seg_len, 0, 32);
    const SortPlan pool_plan = make_sort_plan(1, pooled_len, 0, 32);
    if (seg_workspace_bytes < seg_plan.off_end || pooled_workspace_bytes < pool_plan.off_end) return DML_ERR_WORKSPACE;
    sh = reinterpret_cast<const uint32_t*>(reinterpret_cast<const unsigned char*>(seg_workspace) + seg_plan.off_hist);
    ph = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(pooled_workspace) + pool_plan.off_hist);
  }
  if (n_seg == 0 || seg_len == 0) {
    if (reset && ph) DML_CUDA_TRY(cudaMemsetAsync(ph, 0, (size_t)MAX_PASSES * RADIX * sizeof(uint32_t), stream));
    if (reset && pooled_stats) DML_CUDA_TRY(cudaMemsetAsync(pooled_stats, 0, 4 * sizeof(long long), stream));
    return DML_OK;
  }
  pool_hist_kernel<<<MAX_PASSES + 1, RADIX, 0, stream>>>(sh, n_seg, ph, reset, (const unsigned long long*)seg_stats,
                                                         (unsigned long long*)pooled_stats);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

int dml_ood_roc_fpr(const uint32_t* keys, const long long* seg_stats, int32_t n_seg, int64_t seg_len, double recall_level,
                    const void* workspace, size_t workspace_bytes, double* fpr_out, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!keys || !seg_stats || !workspace || !fpr_out || n_seg < 0 || seg_len < 1 || n_seg > 65535) return DML_ERR_INVALID_ARG;
  if (seg_len >= (1ll << 32)) return DML_ERR_INVALID_ARG;
  if (n_seg == 0) return DML_OK;
  const MetricsPlan m = make_metrics_plan(n_seg, seg_len);
  if (workspace_bytes < m.off_end) return DML_ERR_WORKSPACE;
  const unsigned char* ws = reinterpret_cast<const unsigned char*>(workspace);
  // where dml_ood_eval_segments left the sorted keys: the ping-pong ends in `keys` after an even number of passes
  const uint32_t* sorted = (m.sort.n_passes % 2 == 0) ? keys : reinterpret_cast<const uint32_t*>(ws + m.sort.off_alt);
  roc_fpr_kernel<<<n_seg, SCAN_THREADS, 0, stream>>>(sorted, seg_len, m.tiles_per_seg, reinterpret_cast<const Carry*>(ws + m.off_carry),
                                                     (const unsigned long long*)seg_stats, recall_level, fpr_out);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

int dml_ood_sort(uint32_t* keys, int32_t n_seg, int64_t seg_len, int32_t begin_bit, int32_t end_bit, void* workspace,
                 size_t workspace_bytes, uint32_t** sorted_out, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!keys || !workspace || !sorted_out || n_seg < 0 || seg_len < 0 || n_seg > 65535) return DML_ERR_INVALID_ARG;
  if (begin_bit < 0 || end_bit > 32 || begin_bit > end_bit || seg_len >= (1ll << 32)) return DML_ERR_INVALID_ARG;
  const SortPlan plan = make_sort_plan(n_seg, seg_len, begin_bit, end_bit);
  if (workspace_bytes < plan.off_end) return DML_ERR_WORKSPACE;
  return radix_sort_segments(keys, plan, workspace, sorted_out, stream);
}

int dml_ood_lower_bound(const uint32_t* sorted_keys, int64_t n, const uint32_t* queries, int32_t n_queries,
                        long long* positions, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || n_queries < 0 || (n > 0 && !sorted_keys) || (n_queries > 0 && (!queries || !positions))) return DML_ERR_INVALID_ARG;
  if (n_queries == 0) return DML_OK;
  lower_bound_kernel<<<ceil_div_i(n_queries, 128), 128, 0, stream>>>(sorted_keys, n, queries, n_queries, positions);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

int dml_ood_count_positive(const uint32_t* keys, int64_t n, long long* count, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || !count || (n > 0 && !keys)) return DML_ERR_INVALID_ARG;
  DML_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(long long), stream));
  if (n == 0) return DML_OK;
  long long blocks = (n + 256 * 16 - 1) / (256 * 16);
  if (blocks > 148 * 8) blocks = 148 * 8;
  count_pos_kernel<<<(int)blocks, 256, 0, stream>>>(keys, n, (unsigned long long*)count);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

int dml_ood_scan_range(const uint32_t* sorted_keys, int64_t n, const long long* range_info, double recall_level,
                       void* workspace, size_t workspace_bytes, void* partial_out, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!sorted_keys || !range_info || !workspace || !partial_out || n < 0 || n >= (1ll << 32)) return DML_ERR_INVALID_ARG;
  const MetricsPlan m = make_metrics_plan(1, n);
  if (workspace_bytes < m.off_end) return DML_ERR_WORKSPACE;
  return run_scan(sorted_keys, m, reinterpret_cast<unsigned char*>(workspace), reinterpret_cast<const RangeInfo*>(range_info),
                  recall_level, nullptr, nullptr, reinterpret_cast<TilePartial*>(partial_out), stream);
}

#pragma GCC visibility pop
}  // extern "C"
