// Host emulation of the group-scan kernels (scan_agg / scan_carry / scan_apply / scan_finalize of
// csrc/ood_metrics.cu), thread by thread and tile by tile, running the SAME per-thread code the kernels run
// (csrc/ood_scan_thread.cuh is __host__ __device__).  Test infrastructure: built by tests/test_scan_emulation.py with
// g++, never linked into the product.  It checks two things for every thread of every tile:
//   * the bit-mask form (run_masks / run_aggregate / run_contribution) against a straightforward one-key-at-a-time
//     state machine (the kernels' original formulation, restated below) -- exact equality, AP bits included;
//   * the assembled (auroc, aupr, fpr) -- compared with the CPU oracle by the Python side.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../open-world-semantic-segmentation_b200/csrc/ood_scan_thread.cuh"

using namespace dml;

namespace {
constexpr int THREADS = 256;
constexpr int TILE = THREADS * SCAN_ITEMS;

struct Run {
  uint32_t k[SCAN_ITEMS];
  uint32_t prev, next;
  bool has_prev;
  long long first;
};

Run load_run(const uint32_t* keys, long long n, long long first) {
  Run r;
  for (int j = 0; j < SCAN_ITEMS; ++j) r.k[j] = first + j < n ? keys[first + j] : 0u;
  r.has_prev = first > 0;
  r.prev = first > 0 && first - 1 < n ? keys[first - 1] : 0u;
  r.next = first + SCAN_ITEMS < n ? keys[first + SCAN_ITEMS] : 0u;
  r.first = first;
  return r;
}

// ---- the one-key-at-a-time formulation ----------------------------------------------------------------------
Agg aggregate_v1(const Run& r, long long n) {
  Agg a = {0u, 0u, 0u, 0u};
  uint32_t prev = r.prev;
  bool have_prev = r.has_prev;
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    if (r.first + j < n) {
      const uint32_t k = r.k[j];
      const bool head = !have_prev || ((k >> 1) != (prev >> 1));
      const unsigned p = k & 1u;
      if (head) { a.spos = 0; a.slen = 0; a.head = 1; }
      a.pos += p; a.spos += p; a.slen += 1;
      prev = k; have_prev = true;
    }
  }
  return a;
}

RunContribution contribution_v1(const Run& r, long long n, long long base_P, long long open_pos, long long open_len,
                                long long total_pos, long long first_idx, int t_local, int rem_local) {
  RunContribution c;
  c.auroc = 0; c.ap_sum = 0.0; c.n_groups = 0; c.a_j = -1; c.a_Pl = 0; c.b_j = -1; c.b_Pl = 0;
  const bool open_valid = (base_P - open_pos) < total_pos;
  unsigned long long neg_P = 0, tie = 0;
  int n_neg = 0, Pl = 0, ls = 0, ll = 0, Pl_start = 0;
  bool in_thread = false;
  bool head = !r.has_prev || ((r.k[0] >> 1) != (r.prev >> 1));
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    const long long gi = r.first + j;
    if (gi >= n) break;
    const uint32_t key = r.k[j];
    const int p = (int)(key & 1u);
    if (head) { ls = 0; ll = 0; Pl_start = Pl; in_thread = true; }
    if (!p) { neg_P += (unsigned)Pl; ++n_neg; }
    Pl += p; ls += p; ll += 1;
    bool endg;
    if (gi + 1 >= n) endg = true;
    else {
      const uint32_t nk = (j + 1 < SCAN_ITEMS) ? r.k[j + 1] : r.next;
      endg = (nk >> 1) != (key >> 1);
    }
    if (endg) {
      const long long pos_g = in_thread ? (long long)ls : open_pos + ls;
      const long long len_g = in_thread ? (long long)ll : open_len + ll;
      if (pos_g) {
        const long long neg_g = len_g - pos_g;
        if (neg_g) tie += (unsigned long long)(neg_g * pos_g);
        c.ap_sum += (double)pos_g * ((double)(base_P + Pl) / (double)(first_idx + j + 1));
      }
      const bool valid = in_thread ? (Pl_start < rem_local) : open_valid;
      if (valid) {
        if (Pl <= t_local) { c.a_j = j; c.a_Pl = Pl; }
        else if (c.b_j < 0 || Pl == c.b_Pl) { c.b_j = j; c.b_Pl = Pl; }
      }
      ++c.n_groups;
    }
    head = endg;
  }
  c.auroc = 2ull * ((unsigned long long)n_neg * (unsigned long long)base_P + neg_P) + tie;
  return c;
}

int clamp_local(long long v) { return (int)(v < -1 ? -1 : (v > SCAN_ITEMS + 1 ? SCAN_ITEMS + 1 : v)); }
}  // namespace

// keys: one sorted segment (or a sorted range of a larger ranking that starts on a group boundary, described by
// pos_before / idx_before / total_pos / total_n like dml_ood_scan_range).  out3 = (auroc, aupr, fpr); partial10 = the
// 80-byte range partial; returns the number of threads whose mask form differs from the one-key-at-a-time form.
extern "C" long long scan_emulate(const uint32_t* keys, long long n, long long pos_before, long long idx_before,
                                  long long total_pos, long long total_n, double recall_level, double* out3,
                                  long long* partial10) {
  long long mismatches = 0;
  const long long tiles = n > 0 ? (n + TILE - 1) / TILE : 1;
  // phase 1: per-tile aggregates
  std::vector<Agg> tile_agg(tiles);
  for (long long t = 0; t < tiles; ++t) {
    Agg acc = {0u, 0u, 0u, 0u};
    for (int tid = 0; tid < THREADS; ++tid) {
      const long long first = t * TILE + (long long)tid * SCAN_ITEMS;
      const Run r = load_run(keys, n, first);
      const Agg a = run_aggregate(run_masks(r.k, r.prev, r.has_prev, r.next, first, n));
      const Agg b = aggregate_v1(r, n);
      if (std::memcmp(&a, &b, sizeof(Agg)) != 0) ++mismatches;
      acc = agg_combine(acc, a);
    }
    tile_agg[t] = acc;
  }
  // phase 2: carries
  std::vector<Carry> carry(tiles);
  {
    Carry run = {(unsigned long long)pos_before, 0ull, 0ull};
    unsigned head = 0;
    for (long long t = 0; t < tiles; ++t) {
      carry[t] = run;
      carry_apply(run, head, tile_agg[t]);
    }
  }
  // phase 3: contributions
  const long long tstar = recall_threshold(total_pos, recall_level);
  TilePartial total;
  partial_init(total);
  for (long long t = 0; t < tiles; ++t) {
    Agg excl = {0u, 0u, 0u, 0u};
    const Carry tc = carry[t];
    for (int tid = 0; tid < THREADS; ++tid) {
      const long long first = t * TILE + (long long)tid * SCAN_ITEMS;
      const Run r = load_run(keys, n, first);
      const RunMasks rm = run_masks(r.k, r.prev, r.has_prev, r.next, first, n);
      const long long base_P = (long long)(tc.pos + excl.pos);
      const long long open_pos = (long long)(excl.head ? excl.spos : tc.spos + excl.spos);
      const long long open_len = (long long)(excl.head ? excl.slen : tc.slen + excl.slen);
      const long long first_idx = idx_before + first;
      const int t_local = clamp_local(tstar - base_P), rem_local = clamp_local(total_pos - base_P);
      const RunContribution c = run_contribution(rm, base_P, open_pos, open_len, total_pos, first_idx, t_local, rem_local);
      const RunContribution d = contribution_v1(r, n, base_P, open_pos, open_len, total_pos, first_idx, t_local, rem_local);
      if (c.auroc != d.auroc || std::memcmp(&c.ap_sum, &d.ap_sum, sizeof(double)) != 0 || c.n_groups != d.n_groups ||
          c.a_j != d.a_j || c.b_j != d.b_j || (c.a_j >= 0 && c.a_Pl != d.a_Pl) || (c.b_j >= 0 && c.b_Pl != d.b_Pl))
        ++mismatches;
      TilePartial p;
      partial_init(p);
      p.auroc_num = c.auroc; p.ap_sum = c.ap_sum; p.n_groups = c.n_groups;
      if (c.a_j >= 0) { p.a_idx = first_idx + c.a_j; p.a_tps = base_P + c.a_Pl; p.a_fps = first_idx + c.a_j + 1 - p.a_tps; }
      if (c.b_j >= 0) { p.b_idx = first_idx + c.b_j; p.b_tps = base_P + c.b_Pl; p.b_fps = first_idx + c.b_j + 1 - p.b_tps; }
      partial_merge(total, p);
      excl = agg_combine(excl, run_aggregate(rm));
    }
  }
  if (partial10) std::memcpy(partial10, &total, sizeof(TilePartial));
  // phase 4: finalize (scan_finalize_kernel)
  const double P = (double)total_pos, N = (double)(total_n - total_pos);
  if (total_pos > 0 && total_n - total_pos > 0) {
    out3[0] = (double)total.auroc_num / (2.0 * P * N);
    out3[1] = total.ap_sum / P;
    const double inf = __builtin_inf();
    const double da = total.a_idx >= 0 ? __builtin_fabs((double)total.a_tps / P - recall_level) : inf;
    const double db = total.b_tps != NO_B ? __builtin_fabs((double)total.b_tps / P - recall_level) : inf;
    out3[2] = (double)(db <= da ? total.b_fps : total.a_fps) / N;
  } else {
    out3[0] = out3[1] = out3[2] = __builtin_nan("");
  }
  return mismatches;
}

// dml_ood_keygen's packing (without the fused normalisation) for n (value, positive) pairs; counts[0] = NaNs,
// counts[1] = keys outside the 31-bit window starting at key_base.
extern "C" void pack_keys(const float* values, const uint8_t* positive, long long n, int kind, uint32_t key_base,
                          uint32_t* keys, long long* counts) {
  unsigned n_nan = 0, n_oow = 0;
  for (long long i = 0; i < n; ++i) keys[i] = pack_key(values[i], kind, positive[i] != 0, key_base, n_nan, n_oow);
  counts[0] = n_nan;
  counts[1] = n_oow;
}

