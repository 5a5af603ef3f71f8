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

SortPlan make_sort_plan(int n_seg, long long seg_len, int begin_bit, int end_bit);
// sort; returns pointer to the buffer that holds the result (keys or the alt buffer in workspace)
// `hist_done`: the digit histograms in the workspace were already accumulated by the key-generation kernel
int radix_sort_segments(uint32_t* keys, const SortPlan& plan, void* workspace, uint32_t** sorted, cudaStream_t stream,
                        bool hist_done = false);

}  // namespace dml
