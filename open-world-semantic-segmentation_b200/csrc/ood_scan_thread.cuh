// Per-thread arithmetic of the tie-aware group scan (scan_agg / scan_apply kernels of ood_metrics.cu), written
// on BIT MASKS: a thread owns SCAN_ITEMS = 16 consecutive sorted keys; one cheap sweep turns them into three 16-bit
// masks (positive flag, "starts a score group", "ends a score group") and everything the scan needs -- the thread's
// aggregate, its AUROC / AP contributions, the FPR@recall candidates -- follows from popcounts / find-first-set on
// those masks plus a short loop over the group ends that actually carry positives (rare: positives are ~1 % of
// the pixels).  The per-key instruction count drops from ~130 (one branchy 64-bit state machine step per key) to
// ~25.
//
// The functions are __host__ __device__ and free of CUDA intrinsics so that tests/host/scan_emulation.cpp can run
// the exact same code on the CPU (whole tiles emulated thread by thread) against the Python oracle.
//
// Semantics (anomaly/anom_utils.py:25-78 + scikit-learn's _binary_clf_curve, see the header of ood_metrics.cu):
// keys are sorted ascending, key = (rank key << 1) | positive, so inside a group of equal score the negatives come
// first and the positives last.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define DML_HD __host__ __device__ __forceinline__
#else
#define DML_HD inline
#endif

namespace dml {

constexpr int SCAN_ITEMS = 16;

struct Agg {  // tile-local counts (<= SCAN_TILE)
  unsigned pos, spos, slen, head;
};
struct Carry {  // running state across tiles (64-bit)
  unsigned long long pos, spos, slen;
};
// Per-tile / per-range partial result (10 x 8 bytes; the layout is part of the C ABI, see
// dml_ood_scan_range).  FPR candidates are kept in integers: with T* = the largest tps whose
// float64 recall tps/P is <= recall_level, |tps/P - recall_level| is non-increasing up to T* and
// non-decreasing after it, so the reference's argmin (ties -> later group) is one of
//   a = the LAST group with tps <= T*,   b = the smallest tps > T*, latest group having it.
struct TilePartial {
  unsigned long long auroc_num;
  double ap_sum;
  long long a_idx, a_tps, a_fps;   // a_idx = -1: none
  long long b_tps, b_idx, b_fps;   // b_tps = LLONG_MAX: none
  long long n_groups;
  long long reserved;
};

// ---- ranking keys ------------------------------------------------------------------------------------------
DML_HD uint32_t float_bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, sizeof(u));
  return u;
#endif
}
// Order-preserving packed key of one (value, label) pair: kind 0 ranks a conf map (score = -conf, positives expected at
// LOW conf: anomaly/eval_ood_traditional.py:139-141), kind 1 a plain score (higher = more positive).  Ascending key order
// == descending score order; -0 and +0 are one threshold; the key is taken relative to `key_base` and must fit 31 bits
// (violations are counted and clamped, NaNs counted and ranked last); bit 0 carries the positive flag, so inside a group of
// equal score the negatives sort first.
DML_HD uint32_t pack_key(float v, int kind, bool pos, uint32_t key_base, unsigned& n_nan,
                                             unsigned& n_oow) {
  float f = kind == 0 ? v : -v;
  if (f != f) {
    ++n_nan;
    return 0xfffffffeu | (pos ? 1u : 0u);
  }
  if (f == 0.f) f = 0.f;  // -0 and +0 are one threshold
  const uint32_t u = float_bits(f);
  const uint32_t srt = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  uint32_t rel = srt - key_base;
  if (srt < key_base || rel >= 0x80000000u) {
    ++n_oow;
    rel = srt < key_base ? 0u : 0x7fffffffu;
  }
  return (rel << 1) | (pos ? 1u : 0u);
}


// ---- operators shared by the kernels and the host emulation ------------------------------------------------
DML_HD Agg agg_combine(const Agg& a, const Agg& b) {
  Agg r;
  r.pos = a.pos + b.pos;
  r.spos = b.head ? b.spos : a.spos + b.spos;
  r.slen = b.head ? b.slen : a.slen + b.slen;
  r.head = a.head | b.head;
  return r;
}

constexpr long long NO_B = 0x7fffffffffffffffll;

DML_HD void partial_init(TilePartial& t) {
  t.auroc_num = 0ull; t.ap_sum = 0.0;
  t.a_idx = -1; t.a_tps = 0; t.a_fps = 0;
  t.b_tps = NO_B; t.b_idx = -1; t.b_fps = 0;
  t.n_groups = 0; t.reserved = 0;
}
DML_HD void partial_merge(TilePartial& a, const TilePartial& b) {
  a.auroc_num += b.auroc_num;
  a.ap_sum += b.ap_sum;
  a.n_groups += b.n_groups;
  if (b.a_idx > a.a_idx) { a.a_idx = b.a_idx; a.a_tps = b.a_tps; a.a_fps = b.a_fps; }
  if (b.b_tps < a.b_tps || (b.b_tps == a.b_tps && b.b_idx > a.b_idx)) { a.b_tps = b.b_tps; a.b_idx = b.b_idx; a.b_fps = b.b_fps; }
}
// largest integer t in [0, P] with (double)t / (double)P <= r  (float64 division, like NumPy's recall)
DML_HD long long recall_threshold(long long P, double r) {
  if (P <= 0) return 0;
  const double dP = (double)P;
  double g = floor(r * dP);
  long long t = g < 0.0 ? 0 : (g > dP ? P : (long long)g);
  while (t < P && (double)(t + 1) / dP <= r) ++t;
  while (t > 0 && (double)t / dP > r) --t;
  return t;
}

DML_HD void carry_apply(Carry& c, unsigned& chead, const Agg& b) {
  c.pos += b.pos;
  if (b.head) { c.spos = b.spos; c.slen = b.slen; chead = 1; }
  else { c.spos += b.spos; c.slen += b.slen; }
}

// ---- portable bit helpers ---------------------------------------------------------------------------------
DML_HD int bit_popc(unsigned x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
DML_HD int bit_lowest(unsigned x) {  // index of the lowest set bit (x != 0)
#if defined(__CUDA_ARCH__)
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}
DML_HD int bit_highest(unsigned x) {  // index of the highest set bit (x != 0)
#if defined(__CUDA_ARCH__)
  return 31 - __clz((int)x);
#else
  return 31 - __builtin_clz(x);
#endif
}
DML_HD unsigned mask_upto(int j) { return (2u << j) - 1u; }   // bits 0..j (j in [0, 30])
DML_HD unsigned mask_below(int j) { return (1u << j) - 1u; }  // bits 0..j-1 (j in [0, 31])
// position of the n-th (1-based) set bit of x; needs 1 <= n <= popc(x).  Only reached by the few threads that sit
// on the recall crossing / the last positive, so a short loop is fine.
DML_HD int bit_nth(unsigned x, int n) {
  for (int i = 1; i < n; ++i) x &= x - 1u;
  return bit_lowest(x);
}

// The three masks of one thread's run.  nv = number of valid keys (16 except at the end of a segment).
struct RunMasks {
  unsigned pos;    // bit j: key j is a positive
  unsigned head;   // bit j: key j starts a score group (differs from its predecessor, or is the segment's first key)
  unsigned end;    // bit j: key j ends a score group (its successor differs, or it is the segment's last key)
  int nv;
};

// k[0..15]: the thread's keys; prev / next: the neighbouring keys (has_prev = false at the segment's first key);
// first = index of k[0] inside the segment of seg_len keys.
DML_HD RunMasks run_masks(const uint32_t (&k)[SCAN_ITEMS], uint32_t prev, bool has_prev, uint32_t next, long long first,
                          long long seg_len) {
  RunMasks m;
  const long long left = seg_len - first;
  m.nv = left >= SCAN_ITEMS ? SCAN_ITEMS : (left > 0 ? (int)left : 0);
  unsigned pos = 0u, head = 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    const uint32_t p = j == 0 ? prev : k[j - 1];
    pos |= (k[j] & 1u) << j;
    head |= (((k[j] ^ p) > 1u) ? 1u : 0u) << j;   // group ids differ <=> the XOR has a bit above the positive flag
  }
  if (!has_prev) head |= 1u;
  const unsigned valid = m.nv > 0 ? mask_below(m.nv) : 0u;   // nv = 16 -> 0xffff
  m.pos = pos & valid;
  m.head = head & valid;
  unsigned end = (head >> 1) & valid;
  if (m.nv > 0) {
    // the run's last valid key ends a group if it is the segment's last key (always true for a partial run) or its
    // successor differs (static register index: no local-memory array)
    bool last_ends = true;
    if (m.nv == SCAN_ITEMS) last_ends = (first + SCAN_ITEMS >= seg_len) || ((k[SCAN_ITEMS - 1] ^ next) > 1u);
    end = (end & mask_below(m.nv - 1)) | ((last_ends ? 1u : 0u) << (m.nv - 1));
  }
  m.end = end;
  return m;
}

// Aggregate of the run: positives, and the state of the group still open at its end (its positives / length counted
// from the last head inside the run, or over the whole run when no group starts here).
DML_HD Agg run_aggregate(const RunMasks& m) {
  Agg a;
  a.pos = (unsigned)bit_popc(m.pos);
  if (m.head) {
    const int h = bit_highest(m.head);
    a.head = 1u;
    a.slen = (unsigned)(m.nv - h);
    a.spos = (unsigned)bit_popc(m.pos >> h);
  } else {
    a.head = 0u;
    a.slen = (unsigned)m.nv;
    a.spos = a.pos;
  }
  return a;
}

// What one run contributes, in tile-local terms (the caller widens / reduces across the block).
struct RunContribution {
  unsigned long long auroc;   // 2 * sum over negatives of (positives ranked before) + sum over mixed groups of neg_g * pos_g
  double ap_sum;              // sum over groups ending here of pos_g * tps_g / (tps_g + fps_g)
  int n_groups;               // groups ending in this run
  int a_j, a_Pl;              // last valid group end with tps <= T*   (j = -1: none; Pl = positives of the run up to j)
  int b_j, b_Pl;              // first valid group end with tps > T*, moved to the latest end with the same tps
};

// base_P: positives ranked before the run; open_pos / open_len: positives / length of the group open at the run's
// start (0 / 0 when k[0] starts a group); total_pos = P; first_idx = global rank of k[0];
// t_local = clamp(T* - base_P), rem_local = clamp(P - base_P) to [-1, SCAN_ITEMS + 1].
DML_HD RunContribution run_contribution(const RunMasks& m, long long base_P, long long open_pos, long long open_len,
                                        long long total_pos, long long first_idx, int t_local, int rem_local) {
  RunContribution c;
  c.auroc = 0ull; c.ap_sum = 0.0; c.n_groups = 0;
  c.a_j = -1; c.a_Pl = 0; c.b_j = -1; c.b_Pl = 0;
  if (m.nv == 0) return c;
  const int pos = bit_popc(m.pos);
  const int n_neg = m.nv - pos;
  // sum over the run's negatives of the run's positives ranked before them
  //   = sum over positives i of (negatives above i) = pos * (nv - 1) - sum_i position_i - pos (pos - 1) / 2
  const int sum_pos_idx = bit_popc(m.pos & 0xAAAAu) + 2 * bit_popc(m.pos & 0xCCCCu) + 4 * bit_popc(m.pos & 0xF0F0u) +
                          8 * bit_popc(m.pos & 0xFF00u);
  const int neg_P = pos * (m.nv - 1) - sum_pos_idx - (pos * (pos - 1)) / 2;
  unsigned long long tie = 0ull;
  c.n_groups = bit_popc(m.end);
  const bool carried = (m.head & 1u) == 0u;   // the first group of the run started earlier
  // ---- groups that END here and contain positives (their last key is a positive): tie term + AP term -------
  unsigned pe = m.end & m.pos;
  while (pe) {
    const int e = bit_lowest(pe);
    pe &= pe - 1u;
    const unsigned hb = m.head & mask_upto(e);
    long long pos_g, len_g;
    if (hb) {
      const int h = bit_highest(hb);
      pos_g = bit_popc((m.pos & mask_upto(e)) >> h);
      len_g = e - h + 1;
    } else {
      pos_g = open_pos + bit_popc(m.pos & mask_upto(e));
      len_g = open_len + e + 1;
    }
    const long long neg_g = len_g - pos_g;
    if (neg_g) tie += (unsigned long long)(neg_g * pos_g);
    const int Pl = bit_popc(m.pos & mask_upto(e));
    c.ap_sum += (double)pos_g * ((double)(base_P + Pl) / (double)(first_idx + e + 1));
  }
  c.auroc = 2ull * ((unsigned long long)n_neg * (unsigned long long)base_P + (unsigned long long)neg_P) + tie;

  // ---- FPR candidates: group ends whose group starts before full recall (tps_{g-1} < P) ---------------------
  unsigned vmask;
  {
    int e0 = -1;
    unsigned in_ends = m.end;
    if (carried && m.end) {
      e0 = bit_lowest(m.end);
      in_ends &= ~(1u << e0);
    }
    // groups starting here at h are valid iff (positives of the run below h) < rem_local
    unsigned vin;
    if (rem_local <= 0) vin = 0u;
    else if (rem_local > pos) vin = in_ends;
    else {
      const int q = bit_nth(m.pos, rem_local);   // the group holding the rem_local-th positive is the last valid one
      const unsigned rest = m.end >> q;
      vin = rest ? (in_ends & mask_upto(q + bit_lowest(rest))) : in_ends;
    }
    const bool open_valid = (base_P - open_pos) < total_pos;
    vmask = vin | ((e0 >= 0 && open_valid) ? (1u << e0) : 0u);
  }
  if (vmask) {
    // ends with (positives of the run up to and including j) <= t_local
    unsigned low;
    if (t_local < 0) low = 0u;
    else if (t_local >= pos) low = 0xffffu;
    else low = mask_below(bit_nth(m.pos, t_local + 1));
    const unsigned A = vmask & low;
    if (A) {
      c.a_j = bit_highest(A);
      c.a_Pl = bit_popc(m.pos & mask_upto(c.a_j));
    }
    const unsigned B = vmask & ~low;
    if (B) {
      const int j1 = bit_lowest(B);
      c.b_Pl = bit_popc(m.pos & mask_upto(j1));
      const unsigned above = m.pos >> (j1 + 1);                 // next positive after j1 raises tps
      const unsigned same = above ? mask_below(j1 + 1 + bit_lowest(above)) : 0xffffu;
      c.b_j = bit_highest(B & same);
    }
  }
  return c;
}

}  // namespace dml
