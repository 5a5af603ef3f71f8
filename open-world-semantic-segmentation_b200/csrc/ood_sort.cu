// Segmented one-read/one-write-per-pass LSD radix sort (see ood_sort.cuh).
#include "ood_sort.cuh"

namespace dml {

namespace {

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

template <typename T>
__device__ __forceinline__ T* opaque_ptr(T* p) {
  unsigned long long v = (unsigned long long)p;
  asm volatile("" : "+l"(v));
  return (T*)v;
}

// look-back words: gpu-scope relaxed accesses (served by L2), not the system-scope ones `volatile` emits
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- digit histograms of every pass in one read of the keys --------------------------------
// Warp-private shared histograms updated with shared atomics for every digit place: ATOMS.POPC.INC merges the lanes of
// a warp that hit the same bin in hardware, so the near-uniform lower places of float keys and the heavily skewed top
// place (1-3 distinct exponents per warp) cost the same single instruction (measured in the key-generation kernel, which
// fuses the same counting: 266 us per 46 M keys with atomics for the top place against 306 us with ballot groups).
constexpr int HIST_TILES_PER_BLOCK = 16;

__global__ void __launch_bounds__(SORT_THREADS) hist_kernel(const uint32_t* __restrict__ keys, long long seg_len,
                                                            int n_passes, int s0, int s1, int s2, int s3,
                                                            uint32_t* __restrict__ ghist) {
  __shared__ uint32_t s_h[SORT_WARPS][MAX_PASSES][RADIX];  // 32 KB
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  for (int i = tid; i < SORT_WARPS * MAX_PASSES * RADIX; i += SORT_THREADS) (&s_h[0][0][0])[i] = 0u;
  __syncthreads();
  const int seg = blockIdx.y;
  const uint32_t* k = keys + (size_t)seg * seg_len;
  const long long chunk = (long long)HIST_TILES_PER_BLOCK * SORT_TILE;
  const long long begin = (long long)blockIdx.x * chunk;
  long long end = begin + chunk;
  if (end > seg_len) end = seg_len;
  // each warp walks its own contiguous slice, 32 keys per step
  const long long per_warp = chunk / SORT_WARPS;
  const long long wbeg = begin + (long long)w * per_warp;
  constexpr int U = 8;  // independent 128-byte warp loads in flight per step
  uint32_t* h0 = s_h[w][0];
  uint32_t* h1 = s_h[w][1];
  uint32_t* h2 = s_h[w][2];
  uint32_t* h3 = s_h[w][3];
  auto count = [&](uint32_t key) {
    atomicAdd(h0 + ((key >> s0) & (RADIX - 1)), 1u);
    if (n_passes > 1) atomicAdd(h1 + ((key >> s1) & (RADIX - 1)), 1u);
    if (n_passes > 2) atomicAdd(h2 + ((key >> s2) & (RADIX - 1)), 1u);
    if (n_passes > 3) atomicAdd(h3 + ((key >> s3) & (RADIX - 1)), 1u);
  };
  for (long long off = 0; off < per_warp; off += 32 * U) {
    const long long base = wbeg + off;
    if (base >= end) break;  // warp-uniform
    const uint32_t* src = k + base + lane;
    if (base + 32 * U <= end) {
      uint32_t key[U];
#pragma unroll
      for (int u = 0; u < U; ++u) key[u] = __ldg(src + u * 32);
#pragma unroll
      for (int u = 0; u < U; ++u) count(key[u]);
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (base + u * 32 + lane < end) count(__ldg(src + u * 32));
    }
  }
  __syncthreads();
  for (int i = tid; i < n_passes * RADIX; i += SORT_THREADS) {
    const int p = i >> RADIX_BITS, d = i & (RADIX - 1);
    uint32_t c = 0;
#pragma unroll
    for (int ww = 0; ww < SORT_WARPS; ++ww) c += s_h[ww][p][d];
    if (c) atomicAdd(ghist + ((size_t)seg * MAX_PASSES + p) * RADIX + d, c);
  }
}

// exclusive scan of each 256-bin histogram (one block per (segment, pass))
__global__ void __launch_bounds__(RADIX) hist_scan_kernel(uint32_t* ghist) {
  __shared__ uint32_t s_w[RADIX / 32];
  uint32_t* h = ghist + (size_t)blockIdx.x * RADIX;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const uint32_t c = h[t];
  uint32_t incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) s_w[w] = incl;
  __syncthreads();
  uint32_t base = 0;
  for (int i = 0; i < w; ++i) base += s_w[i];
  h[t] = base + incl - c;
}

template <typename LB>
struct LbTraits;
template <>
struct LbTraits<uint32_t> {
  static constexpr uint32_t LOCAL = 1u << 30, INCL = 2u << 30, MASK = (1u << 30) - 1;
  static constexpr int FLAG_SHIFT = 30;
};
template <>
struct LbTraits<unsigned long long> {
  static constexpr unsigned long long LOCAL = LB_FLAG_LOCAL, INCL = LB_FLAG_INCL, MASK = LB_VALUE_MASK;
  static constexpr int FLAG_SHIFT = 62;
};

// byte `sel` (0..3) of k: the digit when the shift is a multiple of 8 (PRMT, one instruction)
__device__ __forceinline__ uint32_t digit_of(uint32_t k, int shift, uint32_t prmt_sel) {
  (void)shift;
  return __byte_perm(k, 0u, prmt_sel);
}

template <typename LB>
__global__ void __launch_bounds__(SORT_THREADS, 4) onesweep_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                                   long long seg_len, int n_seg, int tiles_per_seg, int shift, int pass,
                                                                   const uint32_t* __restrict__ ghist_excl, LB* lookback,
                                                                   uint32_t* tickets) {
  using T = LbTraits<LB>;
  __shared__ uint32_t s_whist[SORT_WARPS][RADIX];
  __shared__ uint32_t s_keys[SORT_TILE];
  __shared__ uint32_t s_gbase[RADIX];   // global index of the tile's first key of each digit, minus its tile position
  __shared__ uint32_t s_scan[SORT_WARPS];
  __shared__ int s_tile;

  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  // 1-D grid, blocks dealt round-robin to the segments: co-resident tiles come from many segments, which
  // keeps every segment's look-back chain short
  const int seg = (int)(blockIdx.x % (unsigned)n_seg);
  if (tid == 0) s_tile = (int)atomicAdd(tickets + seg, 1u);
  const uint32_t gh = ghist_excl[((size_t)seg * MAX_PASSES + pass) * RADIX + tid];  // prefetch: used after the look-back
#pragma unroll
  for (int i = 0; i < SORT_WARPS; ++i) s_whist[i][tid] = 0u;
  __syncthreads();
  const int tile = s_tile;
  const size_t seg_base = (size_t)seg * (size_t)seg_len;
  const long long tile_off = (long long)tile * SORT_TILE;
  const long long rem = seg_len - tile_off;
  const int nvalid = rem >= SORT_TILE ? SORT_TILE : (int)rem;
  const bool full = nvalid == SORT_TILE;
  const uint32_t sel = 0x4440u + (uint32_t)(shift >> 3);  // shift is a multiple of 8

  // ---- load (warp-striped inside the warp's contiguous slice => LSD-stable order) ----------
  uint32_t key[SORT_ITEMS];
  unsigned peers[SORT_ITEMS];
  const int wbase = w * 32 * SORT_ITEMS;
  const uint32_t* src = in + seg_base + tile_off + wbase + lane;
  if (full) {
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) key[i] = src[i * 32];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) peers[i] = match8_full(digit_of(key[i], shift, sel));
  } else {
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) key[i] = (wbase + i * 32 + lane) < nvalid ? src[i * 32] : 0xffffffffu;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
      const bool valid = (wbase + i * 32 + lane) < nvalid;
      const unsigned vm = __ballot_sync(0xffffffffu, valid);
      const unsigned pm = match8_full(digit_of(key[i], shift, sel));
      peers[i] = valid ? (pm & vm) : 0u;
    }
  }
  // ---- per-warp digit counts, order-free: the lowest lane of every digit group adds the group size to its
  //      warp's row (shared atomic, no serial chain) ------------------------------------------------------
  const unsigned lt = lanemask_lt();
  uint32_t* myhist = s_whist[w];
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; ++i)
    if (peers[i] != 0u && (peers[i] & lt) == 0u) atomicAdd(&myhist[digit_of(key[i], shift, sel)], (uint32_t)__popc(peers[i]));
  __syncthreads();

  // ---- thread t owns digit t: prefix over warps -> tile count; publish the LOCAL look-back entry early ------
  uint32_t run = 0;
#pragma unroll
  for (int ww = 0; ww < SORT_WARPS; ++ww) {
    const uint32_t c = s_whist[ww][tid];
    s_whist[ww][tid] = run;
    run += c;
  }
  LB* lb = lookback + ((size_t)seg * tiles_per_seg + tile) * RADIX;
  st_relaxed(lb + tid, (LB)((LB)run | (tile == 0 ? T::INCL : T::LOCAL)));

  // exclusive scan of the tile's digit counts -> start of every digit inside the sorted tile
  uint32_t incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) s_scan[w] = incl;
  __syncthreads();
  uint32_t dbase = incl - run;
#pragma unroll
  for (int i = 0; i < SORT_WARPS; ++i)
    if (i < w) dbase += s_scan[i];
#pragma unroll
  for (int ww = 0; ww < SORT_WARPS; ++ww) s_whist[ww][tid] += dbase;   // running write position of (warp, digit)
  __syncthreads();

  // ---- rank + local scatter in one sweep: every member reads the running position of its (warp, digit), the
  //      group's lowest lane advances it by the group size; keys land in digit order in shared memory --------
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; ++i) {
    const uint32_t d = digit_of(key[i], shift, sel);
    const uint32_t below = (uint32_t)__popc(peers[i] & lt);
    const uint32_t c = myhist[d];
    __syncwarp();
    if (peers[i] != 0u) {
      if (below == 0) myhist[d] = c + (uint32_t)__popc(peers[i]);
      s_keys[c + below] = key[i];
    }
    __syncwarp();
  }

  // ---- decoupled look-back with a window: LB_WIN predecessor entries are requested together (independent
  //      L2 round trips overlap), then consumed nearest-first until an INCLUSIVE entry closes the prefix ----
  unsigned long long excl = 0;
  if (tile > 0) {
    constexpr int LB_WIN = 8;
    const LB* base = lookback + (size_t)seg * tiles_per_seg * RADIX + tid;
    int p = tile - 1;
    bool done = false;
    while (!done) {
      LB v[LB_WIN];
#pragma unroll
      for (int j = 0; j < LB_WIN; ++j) v[j] = (p - j >= 0) ? ld_relaxed(base + (size_t)(p - j) * RADIX) : (LB)T::INCL;
#pragma unroll
      for (int j = 0; j < LB_WIN; ++j) {
        if (!done) {
          const unsigned flag = (unsigned)(v[j] >> T::FLAG_SHIFT);
          if (flag == 0) break;            // not published yet: re-poll from this predecessor
          excl += (unsigned long long)(v[j] & T::MASK);
          --p;
          if (flag == 2) done = true;
        }
      }
    }
    st_relaxed(lb + tid, (LB)((LB)(excl + run) | T::INCL));
  }
  // segment-relative index (seg_len < 2^32): wraps correctly in 32-bit arithmetic
  s_gbase[tid] = gh + (uint32_t)excl - dbase;
  __syncthreads();

  // ---- coalesced per-digit runs to global -------------------------------------------------------------------
  // the segment base goes through an opaque register pair: nvcc otherwise re-derives seg * seg_len (two wide
  // multiplies) for every key; with it a key costs LDS, PRMT, LDS, IADD, IMAD.WIDE, STG
  uint32_t* dst = opaque_ptr(out + seg_base);
  if (full) {
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
      const int pos = j * SORT_THREADS + tid;
      const uint32_t k = s_keys[pos];
      dst[(uint32_t)(s_gbase[digit_of(k, shift, sel)] + (uint32_t)pos)] = k;
    }
  } else {
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
      const int pos = j * SORT_THREADS + tid;
      if (pos < nvalid) {
        const uint32_t k = s_keys[pos];
        dst[(uint32_t)(s_gbase[digit_of(k, shift, sel)] + (uint32_t)pos)] = k;
      }
    }
  }
}

}  // namespace

SortPlan make_sort_plan(int n_seg, long long seg_len, int begin_bit, int end_bit) {
  SortPlan p;
  p.n_seg = n_seg;
  p.seg_len = seg_len;
  p.tiles_per_seg = (int)((seg_len + SORT_TILE - 1) / SORT_TILE);
  if (p.tiles_per_seg < 1) p.tiles_per_seg = 1;
  p.n_passes = 0;
  for (int b = begin_bit; b < end_bit && p.n_passes < MAX_PASSES; b += RADIX_BITS) p.shifts[p.n_passes++] = b;
  for (int i = p.n_passes; i < MAX_PASSES; ++i) p.shifts[i] = 0;
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t n = (size_t)n_seg * (size_t)seg_len;
  const size_t lb_entry = seg_len < (1ll << 30) ? sizeof(uint32_t) : sizeof(unsigned long long);
  p.off_alt = 0;
  p.off_hist = align(p.off_alt + n * sizeof(uint32_t));
  p.off_lookback = align(p.off_hist + (size_t)n_seg * MAX_PASSES * RADIX * sizeof(uint32_t));
  p.off_ticket = align(p.off_lookback + (size_t)n_seg * p.tiles_per_seg * RADIX * lb_entry);
  p.off_end = align(p.off_ticket + (size_t)n_seg * sizeof(uint32_t));
  return p;
}

int radix_sort_segments(uint32_t* keys, const SortPlan& plan, void* workspace, uint32_t** sorted, cudaStream_t stream,
                        bool hist_done) {
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  uint32_t* alt = reinterpret_cast<uint32_t*>(ws + plan.off_alt);
  uint32_t* ghist = reinterpret_cast<uint32_t*>(ws + plan.off_hist);
  void* lookback = ws + plan.off_lookback;
  uint32_t* tickets = reinterpret_cast<uint32_t*>(ws + plan.off_ticket);
  const bool wide = plan.seg_len >= (1ll << 30);
  if (plan.n_seg == 0 || plan.seg_len == 0 || plan.n_passes == 0) {
    *sorted = keys;
    return DML_OK;
  }
  if (!hist_done) {
    DML_CUDA_TRY(cudaMemsetAsync(ghist, 0, plan.off_lookback - plan.off_hist, stream));
    const long long chunk = (long long)HIST_TILES_PER_BLOCK * SORT_TILE;
    dim3 grid((unsigned)((plan.seg_len + chunk - 1) / chunk), (unsigned)plan.n_seg);
    hist_kernel<<<grid, SORT_THREADS, 0, stream>>>(keys, plan.seg_len, plan.n_passes, plan.shifts[0], plan.shifts[1],
                                                   plan.shifts[2], plan.shifts[3], ghist);
    DML_LAUNCH_CHECK();
  }
  {
    hist_scan_kernel<<<plan.n_seg * MAX_PASSES, RADIX, 0, stream>>>(ghist);
    DML_LAUNCH_CHECK();
  }
  uint32_t* in = keys;
  uint32_t* out = alt;
  const unsigned grid = (unsigned)((size_t)plan.tiles_per_seg * plan.n_seg);
  for (int p = 0; p < plan.n_passes; ++p) {
    DML_CUDA_TRY(cudaMemsetAsync(lookback, 0, plan.off_end - plan.off_lookback, stream));  // look-back + tickets
    if (wide)
      onesweep_kernel<unsigned long long><<<grid, SORT_THREADS, 0, stream>>>(
          in, out, plan.seg_len, plan.n_seg, plan.tiles_per_seg, plan.shifts[p], p, ghist, (unsigned long long*)lookback, tickets);
    else
      onesweep_kernel<uint32_t><<<grid, SORT_THREADS, 0, stream>>>(in, out, plan.seg_len, plan.n_seg, plan.tiles_per_seg,
                                                                  plan.shifts[p], p, ghist, (uint32_t*)lookback, tickets);
    DML_LAUNCH_CHECK();
    uint32_t* t = in; in = out; out = t;
  }
  *sorted = in;
  return DML_OK;
}

}  // namespace dml
