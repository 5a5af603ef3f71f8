// Exact AUROC / AUPR / FPR@recall WITHOUT sorting the negatives ("minority rank" path of kernel (d)).
//
// The three metrics of anomaly/anom_utils.py:25-78 (+ scikit-learn's roc_auc_score / average_precision_score) depend on
// the negatives only through HOW MANY of them fall between / onto consecutive distinct POSITIVE scores:
//   groups in ranking order = ..., [negatives strictly between S[g-1] and S[g]], [score S[g]: pc[g] positives, eq[g]
//   negatives], ...;  AUROC numerator = sum_g neg_g (2 tps_g - pos_g), AP = sum over positive groups, FPR@recall is
//   decided at positive groups and the negative-only groups right after them.
// OOD pixels are rare (~1 % of an image), so instead of radix-sorting every (score, label) pair (4 passes x 8 B / pair,
// issue-bound at ~35 % of HBM) this path
//   1. gathers the positives' pixel indices per segment                    (pos_gather_kernel: reads gt, 1 B / pair)
//   2. fetches their values, normalises + packs the keys, sorts + de-duplicates them, one CTA per segment in shared
//      memory                                                              (pos_sort_kernel: S[g], pc[g])
//   3. streams over all pairs ONCE: normalisation, conf / MMSP / mix maps, packed keys for the pooled metric, and for
//      every negative a lower_bound in the segment's S (value-linear LUT + a few probes, all in shared memory) followed
//      by one shared-memory counter increment                              (rank_kernel: replaces key-gen + 4 sort
//                                                                            passes + 2 scan reads)
//   4. scans the <= pos_capacity groups of every segment                   (rank_scan_kernel -> dml_ood_result)
// All counting is integer and the AP terms are the ones the sort path evaluates, so AUROC / FPR are bit-identical to the
// sort path and AUPR differs by float64 summation order only.  Segments whose positives exceed `pos_capacity` are flagged
// (seg_stats[seg][3] = 1, NaN result): the caller re-evaluates those with the sort path (dml_ood_keygen +
// dml_ood_eval_segments), which has no such limit.
#include "ood_rank.cuh"

namespace dml {
namespace {

constexpr int RANK_SORT_MAX = 32768;    // positives per segment the shared-memory sort handles (128 KB)
constexpr int BUCKET_SORT_MAX = 16384;  // ... of which the bucket sort handles this many (else: bitonic network)
constexpr int BUCKET_SORT_NB = 8192;    // value-linear buckets
constexpr int BUCKET_SORT_RUN = 48;     // largest bucket the per-bucket insertion sort accepts (else: bitonic network)

struct RankWs {  // carve-up of the workspace (byte offsets); per segment `cap` entries
  size_t off_plist, off_S, off_pc, off_cnt, off_G, off_cursor, off_end;
};
RankWs make_rank_ws(int n_seg, int cap) {
  RankWs w;
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t per = (size_t)n_seg * (size_t)cap * sizeof(uint32_t);
  w.off_plist = 0;
  w.off_S = align(w.off_plist + per);
  w.off_pc = align(w.off_S + per);
  w.off_cnt = align(w.off_pc + per);
  w.off_G = align(w.off_cnt + (size_t)n_seg * (2 * (size_t)cap + 2) * sizeof(uint32_t));
  w.off_cursor = align(w.off_G + (size_t)n_seg * sizeof(uint32_t));
  w.off_end = align(w.off_cursor + (size_t)n_seg * sizeof(uint32_t));
  return w;
}

// ---------------------------------------------------------------------------------------------
// 1. positives -> per-segment list of their pixel indices, any order
// ---------------------------------------------------------------------------------------------
// The list holds pixel INDICES (within the segment); pos_sort_kernel turns them into keys.  Keeping the sparse value loads
// out of this kernel leaves a pure scan of the label map: 32 labels per thread and step (two 16-byte loads in flight),
// one ballot, and -- for the few warps that see a positive -- one cursor reservation.
template <typename GT>
__global__ void __launch_bounds__(256) pos_gather_kernel(const GT* __restrict__ gt, uint64_t out_mask,
                                                         const uint8_t* __restrict__ pos_u8, long long seg_len, int cap,
                                                         uint32_t* __restrict__ plist, uint32_t* __restrict__ cursor) {
  const int seg = blockIdx.y;
  const size_t base = (size_t)seg * (size_t)seg_len;
  const int lane = threadIdx.x & 31;
  constexpr int PER = 32;   // consecutive pixels per thread and step
  const bool bytes = pos_u8 != nullptr || sizeof(GT) == 1;
  const uint8_t* gb = pos_u8 ? pos_u8 : reinterpret_cast<const uint8_t*>(gt);
  const bool vec_ok = bytes && (seg_len % PER == 0) && ((reinterpret_cast<uintptr_t>(gb) & 15) == 0);
  const long long nstep = (seg_len + PER - 1) / PER;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long q_end = ((nstep + stride - 1) / stride) * stride;   // block-uniform trip count (warp collectives below)
  uint32_t* out = plist + (size_t)seg * cap;
  const uint32_t single4 = positive_label4(out_mask);
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < q_end; q += stride) {
    unsigned flags = 0u;
    const long long p0 = q * PER;
    if (q < nstep) {
      if (vec_ok) {
        const uint4 w0 = *reinterpret_cast<const uint4*>(gb + base + p0);
        const uint4 w1 = *reinterpret_cast<const uint4*>(gb + base + p0 + 16);
        const uint32_t ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) flags |= byte_mask_to_bits(positive_bytes(ww[j], out_mask, pos_u8 != nullptr, single4)) << (4 * j);
      } else {
        for (int j = 0; j < PER; ++j) {
          const long long p = p0 + j;
          if (p >= seg_len) break;
          bool pos;
          if (pos_u8) pos = pos_u8[base + p] != 0;
          else {
            const long long g = (long long)gt[base + p];
            pos = g >= 0 && g < 64 && ((out_mask >> g) & 1ull);
          }
          flags |= (pos ? 1u : 0u) << j;
        }
      }
    }
    const int c = __popc(flags);
    if (__ballot_sync(0xffffffffu, c > 0) == 0u) continue;   // warp-uniform: most warps see no positive at all
    // warp exclusive prefix of the per-thread counts, one cursor reservation per warp
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t wbase = 0;
    if (lane == 31) wbase = atomicAdd(cursor + seg, (uint32_t)total);
    wbase = __shfl_sync(0xffffffffu, wbase, 31);
    uint32_t dst = wbase + (uint32_t)(incl - c);
    while (flags) {
      const int j = __ffs(flags) - 1;
      flags &= flags - 1u;
      if (dst < (uint32_t)cap) out[dst] = (uint32_t)(p0 + j);
      ++dst;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// 2. per segment: sort in shared memory, distinct scores S[g] + multiplicities pc[g], zeroed counters.
//    Usual case (<= BUCKET_SORT_MAX positives, no long run of equal scores): bucket sort over BUCKET_SORT_NB
//    value-linear buckets (LinIndex is monotone, so bucket order == key order) -- two coalesced reads of the list,
//    one shared atomic per key, then every bucket (a handful of keys) is insertion-sorted by one thread.
//    Otherwise: bitonic network over the padded list.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RANK_THREADS) pos_sort_kernel(uint32_t* __restrict__ plist, const uint32_t* __restrict__ cursor,
                                                                int cap, uint32_t key_base, uint32_t* __restrict__ S,
                                                                uint32_t* __restrict__ pc, uint32_t* __restrict__ cnt,
                                                                uint32_t* __restrict__ Gout, unsigned long long* __restrict__ seg_stats,
                                                                const float* __restrict__ values, const float* __restrict__ minmax,
                                                                int slot, int kind, long long seg_len) {
  extern __shared__ uint32_t s_k[];                        // [n2(cap)] sorted keys | [NB + 1] bucket offsets | u16 [BUCKET_SORT_MAX] ranks
  __shared__ uint32_t s_w[RANK_THREADS / 32];
  __shared__ uint32_t s_mn, s_mx, s_big;
  const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t np_all = cursor[seg];
  const bool overflow = np_all > (uint32_t)cap;
  const int P = overflow ? 0 : (int)np_all;
  if (tid == 0) {
    seg_stats[(size_t)seg * 4 + 0] = np_all;
    seg_stats[(size_t)seg * 4 + 3] = overflow ? 1ull : 0ull;
    s_mn = 0xffffffffu; s_mx = 0u; s_big = 0u;
  }
  int n2 = 2;
  while (n2 < P) n2 <<= 1;
  int n2cap = 2;
  while (n2cap < cap) n2cap <<= 1;
  uint32_t* src = plist + (size_t)seg * cap;
  {
    // pixel indices -> normalised, packed score keys, in place (the sparse value loads of a thread are all in flight at once;
    // the list then also is what dml_ood_rank_export_positives hands to the pooled metric)
    const Norm nm = load_norm(minmax, seg, slot);
    const float* vseg = values + (size_t)seg * (size_t)seg_len;
    for (int i0 = tid; i0 < P; i0 += 8 * RANK_THREADS) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { const int i = i0 + u * RANK_THREADS; v[u] = i < P ? vseg[src[i]] : 0.f; }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * RANK_THREADS;
        unsigned d0 = 0, d1 = 0;
        if (i < P) src[i] = pack_key(apply_norm(nm, v[u]), kind, true, key_base, d0, d1) >> 1;
      }
    }
    __syncthreads();
  }
  bool sorted = false;
  if (P > 64 && P <= BUCKET_SORT_MAX) {
    uint32_t* s_off = s_k + n2cap;                                              // [NB + 1]
    unsigned short* s_r = reinterpret_cast<unsigned short*>(s_off + BUCKET_SORT_NB + 1);   // [BUCKET_SORT_MAX]
    uint32_t mn = 0xffffffffu, mx = 0u;
    for (int i = tid; i < P; i += RANK_THREADS) { const uint32_t k = src[i]; mn = min(mn, k); mx = max(mx, k); }
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    for (int i = tid; i <= BUCKET_SORT_NB; i += RANK_THREADS) s_off[i] = 0u;
    __syncthreads();
    if (lane == 0) { atomicMin(&s_mn, mn); atomicMax(&s_mx, mx); }
    __syncthreads();
    LinIndex li;
    li.init(s_mn, s_mx, BUCKET_SORT_NB, key_base);
    for (int i = tid; i < P; i += RANK_THREADS) s_r[i] = (unsigned short)atomicAdd(&s_off[li(src[i])], 1u);
    __syncthreads();
    // exclusive scan of the bucket sizes (thread t: buckets [8t, 8t + 8)), largest bucket
    constexpr int BPT = BUCKET_SORT_NB / RANK_THREADS;
    uint32_t c[BPT], sum = 0, big = 0;
#pragma unroll
    for (int j = 0; j < BPT; ++j) { c[j] = s_off[tid * BPT + j]; sum += c[j]; big = max(big, c[j]); }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    big = __reduce_max_sync(0xffffffffu, big);
    if (lane == 31) s_w[w] = incl;
    if (lane == 0) atomicMax(&s_big, big);
    __syncthreads();
    uint32_t off = incl - sum;
    for (int i = 0; i < w; ++i) off += s_w[i];
#pragma unroll
    for (int j = 0; j < BPT; ++j) { s_off[tid * BPT + j] = off; off += c[j]; }
    if (tid == RANK_THREADS - 1) s_off[BUCKET_SORT_NB] = off;
    __syncthreads();
    for (int i = tid; i < P; i += RANK_THREADS) { const uint32_t k = src[i]; s_k[s_off[li(k)] + s_r[i]] = k; }
    __syncthreads();
    if (s_big <= (uint32_t)BUCKET_SORT_RUN) {
#pragma unroll 1
      for (int j = 0; j < BPT; ++j) {
        const int b0 = (int)s_off[tid * BPT + j], b1 = (int)s_off[tid * BPT + j + 1];
        for (int i = b0 + 1; i < b1; ++i) {
          const uint32_t k = s_k[i];
          int m = i - 1;
          while (m >= b0 && s_k[m] > k) { s_k[m + 1] = s_k[m]; --m; }
          s_k[m + 1] = k;
        }
      }
      sorted = true;
    } else {
      for (int i = P + tid; i < n2; i += RANK_THREADS) s_k[i] = 0xffffffffu;   // pad for the network below
    }
    __syncthreads();
  } else {
    for (int i = tid; i < n2; i += RANK_THREADS) s_k[i] = i < P ? src[i] : 0xffffffffu;
    __syncthreads();
  }
  if (!sorted) {
    for (int k = 2; k <= n2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = tid; t < (n2 >> 1); t += RANK_THREADS) {
          const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          const uint32_t a = s_k[i], b = s_k[i + j];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { s_k[i] = b; s_k[i + j] = a; }
        }
        __syncthreads();
      }
    }
  }
  // distinct values: thread t owns the contiguous slice [t * per, (t + 1) * per) of the sorted list
  const int per = (n2 + RANK_THREADS - 1) / RANK_THREADS;
  const int b0 = min(tid * per, P), b1 = min(b0 + per, P);
  int heads = 0;
  for (int i = b0; i < b1; ++i) heads += (i == 0 || s_k[i] != s_k[i - 1]) ? 1 : 0;
  int incl = heads;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) s_w[w] = (uint32_t)incl;
  __syncthreads();
  int g = incl - heads;
  int G = 0;
  for (int i = 0; i < RANK_THREADS / 32; ++i) {
    if (i < w) g += (int)s_w[i];
    G += (int)s_w[i];
  }
  uint32_t* Sg = S + (size_t)seg * cap;
  uint32_t* pcg = pc + (size_t)seg * cap;
  uint32_t* c = cnt + (size_t)seg * (2 * (size_t)cap + 2);
  // group g: score, and (in the counter area, cleared below) the index of its first key; the multiplicity is the
  // distance to the next group's first key -- no serial walk over a run of equal scores (clamp plateaus)
  for (int i = b0; i < b1; ++i) {
    if (i == 0 || s_k[i] != s_k[i - 1]) {
      Sg[g] = s_k[i];
      c[g] = (uint32_t)i;
      ++g;
    }
  }
  if (tid == 0) Gout[seg] = (uint32_t)G;
  __syncthreads();
  for (int j = tid; j < G; j += RANK_THREADS) pcg[j] = (j + 1 < G ? c[j + 1] : (uint32_t)P) - c[j];
  __syncthreads();
  for (int i = tid; i < 2 * G + 2; i += RANK_THREADS) c[i] = 0u;
}

// ---------------------------------------------------------------------------------------------
// 3. one pass over all pairs: maps, pooled keys, and the rank of every negative among the positives
// ---------------------------------------------------------------------------------------------
struct RankFuse {
  const float* msp;
  float* msp_norm;
  float* mix;
  float lambda, thr;
};

template <typename GT, int VEC>
__global__ void __launch_bounds__(RANK_THREADS, 1) rank_kernel(const float* __restrict__ values, const float* __restrict__ minmax,
                                                               int slot, float* conf_out, const GT* __restrict__ gt,
                                                               uint64_t out_mask, const uint8_t* __restrict__ pos_u8, int kind,
                                                               long long seg_len, uint32_t key_base, uint32_t* keys_out,
                                                               unsigned long long* seg_stats, const RankFuse fz, int cap,
                                                               int pass_cap, const uint32_t* __restrict__ S,
                                                               const uint32_t* __restrict__ Gin, uint32_t* __restrict__ cnt) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  uint32_t* s_S = reinterpret_cast<uint32_t*>(s_raw);                       // [pass_cap]
  uint32_t* s_cnt = s_S + pass_cap + 1;                                      // [2 * pass_cap + 2] (after the sentinel slot)
  uint32_t* s_lut = s_cnt + 2 * pass_cap + 2;                                // [RANK_LUT]
  __shared__ unsigned s_c[2][RANK_THREADS / 32];
  const int seg = blockIdx.y, tid = threadIdx.x;
  const size_t base = (size_t)seg * (size_t)seg_len;
  const Norm nm = load_norm(minmax, seg, slot);
  const Norm nmm = load_norm(fz.msp ? minmax : nullptr, seg, 1);
  const bool fuse = fz.msp != nullptr;
  const bool overflow = seg_stats[(size_t)seg * 4 + 3] != 0ull;
  const uint32_t single4 = positive_label4(out_mask);
  const int G = overflow ? 0 : (int)Gin[seg];
  const uint32_t* Sg = S + (size_t)seg * cap;
  uint32_t* gcnt = cnt + (size_t)seg * (2 * (size_t)cap + 2);
  // my slice of the segment
  const long long nvec = seg_len / VEC;
  const long long per_block = (nvec + gridDim.x - 1) / gridDim.x;
  const long long v0 = (long long)blockIdx.x * per_block;
  const long long v1 = min(v0 + per_block, nvec);
  unsigned n_nan = 0, n_oow = 0;

  int g0 = 0;
  bool first_pass = true;
  do {
    const int gn = min(pass_cap, G - g0);
    const bool last_pass = g0 + gn >= G;
    // ---- stage this pass's positives, zero its counters, build the index table ------------------------------
    for (int i = tid; i < gn; i += RANK_THREADS) s_S[i] = Sg[g0 + i];
    if (tid == 0) s_S[gn] = 0xffffffffu;                                    // sentinel behind the last positive score
    for (int i = tid; i < 2 * gn + 2; i += RANK_THREADS) s_cnt[i] = 0u;
    const uint32_t s_prev = g0 > 0 ? Sg[g0 - 1] : 0u;
    const bool one_pass = g0 == 0 && last_pass;                             // the usual case: no pass bookkeeping per pixel
    __syncthreads();
    SmemTable tab;
    tab.build(s_S, s_lut, gn, key_base);

    for (long long q = v0 + tid; q < v1; q += RANK_THREADS) {
      const size_t i = base + (size_t)q * VEC;
      float v[VEC];
      bool pos[VEC];
      if constexpr (VEC == 4) {
        const float4 t = *reinterpret_cast<const float4*>(values + i);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        if (pos_u8 || sizeof(GT) == 1) {
          const uint32_t g = *reinterpret_cast<const uint32_t*>((pos_u8 ? pos_u8 : reinterpret_cast<const uint8_t*>(gt)) + i);
          const uint32_t pm = positive_bytes(g, out_mask, pos_u8 != nullptr, single4);
#pragma unroll
          for (int j = 0; j < 4; ++j) pos[j] = ((pm >> (8 * j)) & 1u) != 0u;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const long long g = (long long)gt[i + j];
            pos[j] = g >= 0 && g < 64 && ((out_mask >> g) & 1ull);
          }
        }
      } else {
        v[0] = values[i];
        if (pos_u8) pos[0] = pos_u8[i] != 0;
        else {
          const long long g = (long long)gt[i];
          pos[0] = g >= 0 && g < 64 && ((out_mask >> g) & 1ull);
        }
      }
      uint32_t key[VEC];
      bool bad[VEC];
      apply_norm_vec<VEC>(nm, v);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        unsigned c_nan = 0, c_oow = 0;
        key[j] = pack_key(v[j], kind, pos[j], key_base, c_nan, c_oow);
        bad[j] = c_nan != 0;
        if (first_pass) { n_nan += c_nan; n_oow += c_oow; }
      }
      if (first_pass) {
        if constexpr (VEC == 4) {
          if (keys_out) *reinterpret_cast<uint4*>(keys_out + i) = make_uint4(key[0], key[1], key[2], key[3]);
          if (conf_out) *reinterpret_cast<float4*>(conf_out + i) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
          if (keys_out) keys_out[i] = key[0];
          if (conf_out) conf_out[i] = v[0];
        }
        if (fuse) {
          float m[VEC], mn[VEC], mx[VEC];
          if constexpr (VEC == 4) {
            const float4 t = *reinterpret_cast<const float4*>(fz.msp + i);
            m[0] = t.x; m[1] = t.y; m[2] = t.z; m[3] = t.w;
          } else {
            m[0] = fz.msp[i];
          }
#pragma unroll
          for (int j = 0; j < VEC; ++j) mn[j] = m[j];
          apply_norm_vec<VEC>(nmm, mn);
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            // NumPy: c = 1 / (1 + exp(lamda * (e - thre))); mix = c*e + (1-c)*mmsp   (float32)
            const float c = mix_coefficient_fast(v[j], fz.lambda, fz.thr);
            mx[j] = __fadd_rn(__fmul_rn(c, v[j]), __fmul_rn(__fsub_rn(1.0f, c), mn[j]));
          }
          if constexpr (VEC == 4) {
            if (fz.msp_norm) *reinterpret_cast<float4*>(fz.msp_norm + i) = make_float4(mn[0], mn[1], mn[2], mn[3]);
            if (fz.mix) *reinterpret_cast<float4*>(fz.mix + i) = make_float4(mx[0], mx[1], mx[2], mx[3]);
          } else {
            if (fz.msp_norm) fz.msp_norm[i] = mn[0];
            if (fz.mix) fz.mix[i] = mx[0];
          }
        }
      }
      if (gn == 0) continue;
      // ---- negatives: lower bound among this pass's positive scores, one counter per (interval | tie) ---------
      // (a register-counted short cut for keys above every positive was measured and dropped: the extra branch cost
      //  more than the searches it saved -- most negatives of an image are NOT above its largest positive score)
      int lo[VEC], hi[VEC];
      uint32_t sk[VEC];
      bool act[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        sk[j] = key[j] >> 1;
        act[j] = !pos[j] && !bad[j];
        tab.range(sk[j], lo[j], hi[j]);
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j) lo[j] = tab.finish(sk[j], lo[j], hi[j]);
      // (s_S[gn] holds 0xffffffff, which no 31-bit score key equals: no bound test on l)
      if (one_pass) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const int l = lo[j];
          const int slot = 2 * l + (s_S[l] == sk[j] ? 1 : 0);      // for every pixel: only the atomic is predicated
          if (act[j]) atomicAdd(&s_cnt[slot], 1u);
        }
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const int l = lo[j];
          const bool eq = s_S[l] == sk[j];
          // a negative belongs to the pass whose positives bracket it from above: (S[g0-1], S[g0+gn-1]] -- and to the
          // last pass when it lies above every positive
          const bool mine = (l > 0 || g0 == 0 || sk[j] > s_prev) && (l < gn || last_pass);
          if (act[j] && mine) atomicAdd(&s_cnt[2 * l + (eq ? 1 : 0)], 1u);
        }
      }
    }
    __syncthreads();
    for (int c = tid; c < 2 * gn + 1; c += RANK_THREADS) {
      const uint32_t n = s_cnt[c];
      if (n) atomicAdd(gcnt + 2 * (size_t)g0 + c, n);
    }
    __syncthreads();
    g0 += pass_cap;
    first_pass = false;
  } while (g0 < G);

  n_nan = __reduce_add_sync(0xffffffffu, n_nan);
  n_oow = __reduce_add_sync(0xffffffffu, n_oow);
  if ((tid & 31) == 0) { s_c[0][tid >> 5] = n_nan; s_c[1][tid >> 5] = n_oow; }
  __syncthreads();
  if (tid < 2) {
    unsigned t = 0;
    for (int wv = 0; wv < RANK_THREADS / 32; ++wv) t += s_c[tid][wv];
    if (t) atomicAdd(seg_stats + (size_t)seg * 4 + 1 + tid, (unsigned long long)t);
  }
}

// ---------------------------------------------------------------------------------------------
// 4. group scan on the compressed form: pc[g] positives and eq[g] negatives at score S[g], bt[g] negatives strictly
//    between S[g-1] and S[g] (bt[G]: above every positive).  One CTA per segment.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RANK_THREADS) rank_scan_kernel(const uint32_t* __restrict__ pc, const uint32_t* __restrict__ cnt,
                                                                 const uint32_t* __restrict__ Gin, int cap, long long seg_len,
                                                                 const unsigned long long* __restrict__ seg_stats,
                                                                 double recall_level, dml_ood_result* __restrict__ results) {
  __shared__ unsigned long long s_T[RANK_THREADS], s_F[RANK_THREADS];   // exclusive prefixes per thread
  __shared__ unsigned long long s_wT[RANK_THREADS / 32], s_wF[RANK_THREADS / 32];
  __shared__ unsigned long long s_au[RANK_THREADS / 32];
  __shared__ double s_ap[RANK_THREADS / 32];
  __shared__ int s_gstar;
  const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const long long P = (long long)seg_stats[(size_t)seg * 4 + 0];
  const long long N = seg_len - P;
  const bool overflow = seg_stats[(size_t)seg * 4 + 3] != 0ull;
  dml_ood_result o;
  o.n_pos = P; o.n_neg = N;
  o.n_nan = (long long)seg_stats[(size_t)seg * 4 + 1];
  o.n_groups = -1;   // the number of distinct NEGATIVE-only scores is not known on this path
  if (P <= 0 || N <= 0 || overflow) {
    if (tid == 0) {
      const double nan = __longlong_as_double(0x7ff8000000000000ll);
      o.auroc = o.aupr = o.fpr = nan;
      results[seg] = o;
    }
    return;
  }
  const int G = (int)Gin[seg];
  const uint32_t* pcg = pc + (size_t)seg * cap;
  const uint32_t* c = cnt + (size_t)seg * (2 * (size_t)cap + 2);   // c[2g] = bt[g], c[2g+1] = eq[g]
  const int per = (G + RANK_THREADS - 1) / RANK_THREADS;
  const int b0 = min(tid * per, G), b1 = min(b0 + per, G);
  unsigned long long tp = 0, fn = 0;
  for (int g = b0; g < b1; ++g) { tp += pcg[g]; fn += (unsigned long long)c[2 * g] + c[2 * g + 1]; }
  unsigned long long it = tp, in_ = fn;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned long long a = __shfl_up_sync(0xffffffffu, it, off), b = __shfl_up_sync(0xffffffffu, in_, off);
    if (lane >= off) { it += a; in_ += b; }
  }
  if (lane == 31) { s_wT[w] = it; s_wF[w] = in_; }
  if (tid == 0) s_gstar = -1;
  __syncthreads();
  unsigned long long T = it - tp, F = in_ - fn;
  for (int i = 0; i < w; ++i) { T += s_wT[i]; F += s_wF[i]; }
  s_T[tid] = T; s_F[tid] = F;
  const long long tstar = recall_threshold(P, recall_level);
  unsigned long long au = 0ull;
  double ap = 0.0;
  int gs_mine = -1;
  for (int g = b0; g < b1; ++g) {
    const unsigned long long bt = c[2 * g], eq = c[2 * g + 1], pg = pcg[g];
    au += bt * 2ull * T;                    // negative-only groups before S[g]: tps = T, no positives
    T += pg;
    F += bt + eq;
    au += eq * (2ull * T - pg);             // the group of S[g]
    ap += (double)pg * ((double)T / (double)(T + F));
    if ((long long)T <= tstar) gs_mine = g; // tps is strictly increasing: the last group with tps <= T*
  }
  if (gs_mine >= 0) atomicMax(&s_gstar, gs_mine);
  au = warp_reduce_sum_u64(au);
  // fixed-order float64 reduction: lanes by xor-shuffle tree, warps sequentially
  ap = warp_reduce_sum_d(ap);
  if (lane == 0) { s_au[w] = au; s_ap[w] = ap; }
  __syncthreads();
  if (tid != 0) return;
  unsigned long long au_t = 0ull;
  double ap_t = 0.0;
  for (int i = 0; i < RANK_THREADS / 32; ++i) { au_t += s_au[i]; ap_t += s_ap[i]; }
  au_t += 2ull * (unsigned long long)P * (unsigned long long)c[2 * G];   // negatives above every positive
  // cumulative (tps, fps) through group g, g in [-1, G)
  auto prefix = [&](int g, unsigned long long& Tg, unsigned long long& Fg) {
    if (g < 0) { Tg = 0; Fg = 0; return; }
    const int owner = g / per;
    Tg = s_T[owner]; Fg = s_F[owner];
    for (int k = owner * per; k <= g; ++k) { Tg += pcg[k]; Fg += (unsigned long long)c[2 * k] + c[2 * k + 1]; }
  };
  // FPR candidates (see TilePartial in ood_scan_thread.cuh): a = the LAST group with tps <= T*, b = the smallest
  // tps > T* and the latest group having it.  Groups count up to the first one with full recall (= group G-1); the
  // negative-only groups that follow a positive group share its tps and come later in the ranking.
  const int gs = s_gstar;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  double da = inf, db = inf;
  unsigned long long a_fps = 0, b_fps = 0;
  {
    unsigned long long Tg, Fg;
    prefix(gs, Tg, Fg);
    const unsigned long long trail = (gs + 1 <= G - 1) ? c[2 * (gs + 1)] : 0ull;
    if (gs >= 0 || trail > 0) {
      da = fabs((double)Tg / (double)P - recall_level);
      a_fps = Fg + trail;
    }
  }
  if (gs + 1 <= G - 1) {
    unsigned long long Tg, Fg;
    prefix(gs + 1, Tg, Fg);
    const unsigned long long trail = (gs + 2 <= G - 1) ? c[2 * (gs + 2)] : 0ull;
    db = fabs((double)Tg / (double)P - recall_level);
    b_fps = Fg + trail;
  }
  o.auroc = (double)au_t / (2.0 * (double)P * (double)N);
  o.aupr = ap_t / (double)P;
  o.fpr = (double)(db <= da ? b_fps : a_fps) / (double)N;
  results[seg] = o;
}

// positives of every segment of the last dml_ood_rank_segments call -> appended to one list (pooled metric)
constexpr int EXPORT_SPLIT = 8;
__global__ void __launch_bounds__(256) export_positives_kernel(const uint32_t* __restrict__ plist, const uint32_t* __restrict__ cursor,
                                                               int cap, uint32_t* __restrict__ out, long long out_capacity,
                                                               unsigned long long* __restrict__ count) {
  // EXPORT_SPLIT CTAs per segment (one CTA copied its ~9000 keys with a single load in flight per thread: 18 us per batch);
  // each reserves its own piece of the output list -- the order of the list is irrelevant
  __shared__ unsigned long long s_base;
  const int seg = blockIdx.x;
  const uint32_t np_all = cursor[seg];
  const uint32_t P = np_all > (uint32_t)cap ? 0u : np_all;   // an overflowed segment exports nothing: the count then falls short
  const uint32_t piece = (P + EXPORT_SPLIT - 1) / EXPORT_SPLIT;
  const uint32_t i0 = min(blockIdx.y * piece, P), i1 = min(i0 + piece, P);
  if (i0 == i1) return;
  if (threadIdx.x == 0) s_base = atomicAdd(count, (unsigned long long)(i1 - i0));
  __syncthreads();
  const unsigned long long base = s_base;
  if (base + (i1 - i0) > (unsigned long long)out_capacity) return;
  const uint32_t* src = plist + (size_t)seg * cap + i0;
  for (uint32_t i = threadIdx.x; i < i1 - i0; i += blockDim.x) out[base + i] = src[i];
}

}  // namespace
}  // namespace dml

using namespace dml;

extern "C" {
#pragma GCC visibility push(default)

size_t dml_ood_rank_workspace_bytes(int32_t n_seg, int32_t pos_capacity) {
  if (n_seg <= 0 || pos_capacity <= 0) return 256;
  return make_rank_ws(n_seg, pos_capacity).off_end;
}

int dml_ood_rank_segments(const float* values, const float* minmax, int32_t minmax_slot, float* conf_out,
                          const uint8_t* gt_u8, const int64_t* gt_i64, uint64_t out_label_mask, const uint8_t* pos_u8,
                          int32_t score_kind, uint32_t key_base, int32_t n_seg, int64_t seg_len, uint32_t* keys_out,
                          long long* seg_stats, const float* msp, float* msp_norm_out, float* mix_out, float lambda,
                          float thr, int32_t pos_capacity, double recall_level, void* workspace, size_t workspace_bytes,
                          dml_ood_result* results, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!values || !seg_stats || !workspace || !results || n_seg < 0 || seg_len < 0 || n_seg > 65535) return DML_ERR_INVALID_ARG;
  if (seg_len >= (1ll << 32)) return DML_ERR_INVALID_ARG;
  const int nsrc = (gt_u8 != nullptr) + (gt_i64 != nullptr) + (pos_u8 != nullptr);
  if (nsrc != 1) return DML_ERR_INVALID_ARG;
  if (minmax && (minmax_slot < 0 || minmax_slot > 1)) return DML_ERR_INVALID_ARG;
  if (score_kind != 0 && score_kind != 1) return DML_ERR_INVALID_ARG;
  if ((msp_norm_out || mix_out) && (!msp || !minmax || minmax_slot != 0)) return DML_ERR_INVALID_ARG;
  if (pos_capacity < 1 || pos_capacity > RANK_SORT_MAX) return DML_ERR_INVALID_ARG;
  if (n_seg == 0) return DML_OK;
  const RankWs w = make_rank_ws(n_seg, pos_capacity);
  if (workspace_bytes < w.off_end) return DML_ERR_WORKSPACE;
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  uint32_t* plist = reinterpret_cast<uint32_t*>(ws + w.off_plist);
  uint32_t* S = reinterpret_cast<uint32_t*>(ws + w.off_S);
  uint32_t* pc = reinterpret_cast<uint32_t*>(ws + w.off_pc);
  uint32_t* cnt = reinterpret_cast<uint32_t*>(ws + w.off_cnt);
  uint32_t* G = reinterpret_cast<uint32_t*>(ws + w.off_G);
  uint32_t* cursor = reinterpret_cast<uint32_t*>(ws + w.off_cursor);
  unsigned long long* st = (unsigned long long*)seg_stats;
  DML_CUDA_TRY(cudaMemsetAsync(seg_stats, 0, (size_t)n_seg * 4 * sizeof(long long), stream));
  DML_CUDA_TRY(cudaMemsetAsync(cursor, 0, (size_t)n_seg * sizeof(uint32_t), stream));
  const long long* g64 = (const long long*)gt_i64;

  // 1. positives
  if (seg_len > 0) {
    long long bx = (seg_len / 32 + 255) / 256;
    // one resident wave (4 CTAs of 256 threads per SM): a partial second wave would cost a whole wave's time, and 8 or 16
    // CTAs per SM measured slower (per-image stage 13.08 / 13.18 / 13.22 ms per 1500 images)
    const long long capb = n_seg >= 148 * 4 ? 1 : (148 * 4) / n_seg;
    if (bx > capb) bx = capb;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)n_seg);
    if (gt_i64)
      pos_gather_kernel<long long><<<grid, 256, 0, stream>>>(g64, out_label_mask, nullptr, seg_len, pos_capacity, plist, cursor);
    else
      pos_gather_kernel<uint8_t><<<grid, 256, 0, stream>>>(gt_u8, out_label_mask, pos_u8, seg_len, pos_capacity, plist, cursor);
    DML_LAUNCH_CHECK();
  }
  // 2. sort + distinct
  {
    int n2 = 2;
    while (n2 < pos_capacity) n2 <<= 1;
    // sorted keys | bucket offsets | per-key ranks of the bucket sort (<= 128 KB + 32 KB + 32 KB)
    const size_t smem = (size_t)n2 * sizeof(uint32_t) + (size_t)(BUCKET_SORT_NB + 1) * sizeof(uint32_t) +
                        (size_t)BUCKET_SORT_MAX * sizeof(unsigned short) + 16;
    DML_CUDA_TRY(cudaFuncSetAttribute(pos_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pos_sort_kernel<<<n_seg, RANK_THREADS, smem, stream>>>(plist, cursor, pos_capacity, key_base, S, pc, cnt, G, st, values, minmax,
                                                           minmax_slot, score_kind, seg_len);
    DML_LAUNCH_CHECK();
  }
  // 3. rank
  {
    // positives of one pass: as many as fit next to their counters and the index table in 227 KB of shared memory
    int pass_cap = pos_capacity < 12288 ? pos_capacity : 12288;
    pass_cap = (pass_cap + 3) & ~3;
    const size_t smem = ((size_t)pass_cap + 1) * 4 + ((size_t)2 * pass_cap + 2) * 4 + SmemTable::lut_bytes() + 16;
    auto al = [](const void* q, size_t a) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % a) == 0; };
    const bool vec4 = (seg_len % 4 == 0) && al(values, 16) && al(keys_out, 16) && al(conf_out, 16) && al(msp, 16) &&
                      al(msp_norm_out, 16) && al(mix_out, 16) && al(gt_u8, 4) && al(pos_u8, 4);
    RankFuse fz;
    fz.msp = (msp_norm_out || mix_out) ? msp : nullptr;
    fz.msp_norm = msp_norm_out; fz.mix = mix_out; fz.lambda = lambda; fz.thr = thr;
    // one CTA per SM; a segment is split over floor(148 / n_seg) CTAs when there are fewer segments than SMs
    int bpi = n_seg >= 148 ? 1 : 148 / n_seg;
    const long long nvec = seg_len / (vec4 ? 4 : 1);
    if ((long long)bpi * RANK_THREADS > nvec) bpi = (int)((nvec + RANK_THREADS - 1) / RANK_THREADS);
    if (bpi < 1) bpi = 1;
    dim3 grid((unsigned)bpi, (unsigned)n_seg);
#define DML_RANK_LAUNCH(GT, V, gtp, posp)                                                                                   \
  do {                                                                                                                        \
    DML_CUDA_TRY(cudaFuncSetAttribute(rank_kernel<GT, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    rank_kernel<GT, V><<<grid, RANK_THREADS, smem, stream>>>(values, minmax, minmax_slot, conf_out, gtp, out_label_mask, posp, \
                                                             score_kind, seg_len, key_base, keys_out, st, fz, pos_capacity,   \
                                                             pass_cap, S, G, cnt);                                             \
  } while (0)
    if (gt_i64) {
      if (vec4) DML_RANK_LAUNCH(long long, 4, g64, nullptr);
      else DML_RANK_LAUNCH(long long, 1, g64, nullptr);
    } else {
      if (vec4) DML_RANK_LAUNCH(uint8_t, 4, gt_u8, pos_u8);
      else DML_RANK_LAUNCH(uint8_t, 1, gt_u8, pos_u8);
    }
#undef DML_RANK_LAUNCH
    DML_LAUNCH_CHECK();
  }
  // 4. scan
  rank_scan_kernel<<<n_seg, RANK_THREADS, 0, stream>>>(pc, cnt, G, pos_capacity, seg_len, st, recall_level, results);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

int dml_ood_rank_export_positives(const void* rank_workspace, size_t workspace_bytes, int32_t n_seg, int32_t pos_capacity,
                                  uint32_t* out, int64_t out_capacity, long long* count, dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!rank_workspace || !count || n_seg < 0 || pos_capacity < 1 || out_capacity < 0 || (out_capacity > 0 && !out)) return DML_ERR_INVALID_ARG;
  if (n_seg == 0) return DML_OK;
  const RankWs w = make_rank_ws(n_seg, pos_capacity);
  if (workspace_bytes < w.off_end) return DML_ERR_WORKSPACE;
  const unsigned char* ws = reinterpret_cast<const unsigned char*>(rank_workspace);
  export_positives_kernel<<<dim3((unsigned)n_seg, EXPORT_SPLIT), 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(ws + w.off_plist),
                                                     reinterpret_cast<const uint32_t*>(ws + w.off_cursor), pos_capacity, out,
                                                     out_capacity, (unsigned long long*)count);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

#pragma GCC visibility pop
}  // extern "C"
