// Segmented LSD radix sort of packed 32-bit keys (8-bit digits, one read + one write per pass,
// decoupled look-back across tiles) -- the sorting stage of the exact OOD metrics (kernel (d)).
//
// Replaces the three independent host sorts the reference runs per image
// (anomaly/anom_utils.py:41 np.argsort(kind="mergesort"), and the two inside sklearn's
// roc_auc_score / average_precision_score called at anomaly/anom_utils.py:74-75).
//
// Ranking is atomic-free: per warp, __match_any_sync groups equal digits, the group leader bumps a
// warp-private shared-memory counter with plain LDS/STS, ranks follow from the peer mask.  Keys of
// a tile are staged through shared memory so that global writes form contiguous per-digit runs.
#pragma once
#include "dml_common.cuh"

namespace dml {

constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 4096 keys
constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int MAX_PASSES = 4;

constexpr unsigned long long LB_FLAG_LOCAL = 1ull << 62;
constexpr unsigned long long LB_FLAG_INCL = 2ull << 62;
constexpr unsigned long long LB_VALUE_MASK = (1ull << 62) - 1;

struct SortPlan {
  int n_seg;
  long long seg_len;
  int tiles_per_seg;
  int n_passes;
  int shifts[MAX_PASSES];
  // workspace carve-up (byte offsets)
  size_t off_alt, off_hist, off_lookback, off_ticket, off_end;
};

// Peer mask of the lanes holding the same 8-bit digit from 8 warp ballots (full-rate VOTE + LOP3; MATCH.ANY is far slower
// on sm_100): per bit a predicate test, VOTE and a predicated NOT; the eight terms are folded with
// three-input LOP3s (a & b & c), four instead of seven ANDs.  nvcc's own code for a C++ loop over __ballot_sync spends 6
// per bit, and ~30 where it cannot prove the warp converged (WARPSYNC / ENDCOLLECTIVE around every vote).
__device__ __forceinline__ unsigned match8_full(uint32_t d) {
  unsigned peers;
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " .reg .b32 a, b, t;\n"
      " and.b32 t, %1, 1;   setp.ne.u32 p, t, 0; vote.sync.ballot.b32 %0, p, 0xffffffff; @!p not.b32 %0, %0;\n"
      " and.b32 t, %1, 2;   setp.ne.u32 p, t, 0; vote.sync.ballot.b32 a, p, 0xffffffff; @!p not.b32 a, a;\n"
      " and.b32 t, %1, 4;   setp.ne.u32 p, t, 0; vote.sync.ballot.b32 b, p, 0xffffffff; @!p not.b32 b, b;\n"
      " lop3.b32 %0, %0, a, b, 0x80;\n"
      " and.b32 t, %1, 8;   setp.ne.u32 p, t, 0; vote.sync.ballot.b32 a, p, 0xffffffff; @!p not.b32 a, a;\n"
      " and.b32 t, %1, 16;  setp.ne.u32 p, t, 0; vote.sync.ballot.b32 b, p, 0xffffffff; @!p not.b32 b, b;\n"
      " lop3.b32 %0, %0, a, b, 0x80;\n"
      " and.b32 t, %1, 32;  setp.ne.u32 p, t, 0; vote.sync.ballot.b32 a, p, 0xffffffff; @!p not.b32 a, a;\n"
      " and.b32 t, %1, 64;  setp.ne.u32 p, t, 0; vote.sync.ballot.b32 b, p, 0xffffffff; @!p not.b32 b, b;\n"
      " lop3.b32 %0, %0, a, b, 0x80;\n"
      " and.b32 t, %1, 128; setp.ne.u32 p, t, 0; vote.sync.ballot.b32 a, p, 0xffffffff; @!p not.b32 a, a;\n"
      " and.b32 %0, %0, a;\n"
      "}\n"
      : "=r"(peers)
      : "r"(d));
  return peers;
}

// ballot of a per-lane flag as a bare VOTE: inside loops whose convergence nvcc cannot prove, the C++ __ballot_sync is
// wrapped in a WARPSYNC / ENDCOLLECTIVE protocol that costs ~10 instructions per call.  Every lane of the warp must
// reach the call (the callers iterate to warp-uniform bounds).
__device__ __forceinline__ unsigned ballot_all(bool flag) {
  unsigned r;
  asm volatile(
      "{\n"
      " .reg .pred q;\n"
      " setp.ne.u32 q, %1, 0;\n"
      " vote.sync.ballot.b32 %0, q, 0xffffffff;\n"
      "}\n"
      : "=r"(r)
      : "r"((unsigned)flag));
  return r;
}

SortPlan make_sort_plan(int n_seg, long long seg_len, int begin_bit, int end_bit);
// sort; returns pointer to the buffer that holds the result (keys or the alt buffer in workspace)
// `hist_done`: the digit histograms in the workspace were already accumulated by the key-generation kernel
int radix_sort_segments(uint32_t* keys, const SortPlan& plan, void* workspace, uint32_t** sorted, cudaStream_t stream,
                        bool hist_done = false);

}  // namespace dml
