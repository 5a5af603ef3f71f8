// Shared device helpers of the minority-rank metric kernels (ood_rank.cu: per segment; ood_pool_rank.cu: pooled).
#pragma once
#include "ood_sort.cuh"
#include "ood_scan_thread.cuh"

namespace dml {

constexpr int RANK_THREADS = 1024;
constexpr int RANK_LUT = 8192;          // value-linear index table over [f(first), f(last)] of a sorted key table

// float whose order is the key order (inverse of pack_key's sortable image): kind 0 -> conf, kind 1 -> -score
__device__ __forceinline__ float key_float(uint32_t skey, uint32_t key_base) {
  const uint32_t srt = skey + key_base;
  const uint32_t u = (srt & 0x80000000u) ? (srt ^ 0x80000000u) : ~srt;
  return __uint_as_float(u);
}

// EDS / MMSP mix coefficient 1 / (1 + exp(lambda (conf - thr))) (anomaly/eval_ood_traditional.py:101-106) for the one-pass
// rank kernel, which is issue-bound: exp as ex2.approx of the scaled argument, reciprocal as rcp.approx -- 4 instructions
// instead of ~20 for expf + IEEE reciprocal.  The coefficient is within 7e-7 (absolute) of the correctly rounded one
// (argument rounding at |lambda (conf - thr)| <= 40 dominates: 2.6e-6 relative on exp, times c (1 - c) <= 1/4), the mix map
// within the same bound: inside the 1e-5 bar, but no longer bit-identical to dml_ood_keygen / dml_scores_finalize, which keep
// the IEEE forms.  exp = inf gives 0 and exp = 0 gives 1 like the reference's float32 arithmetic.
__device__ __forceinline__ float mix_coefficient_fast(float v, float lambda, float thr) {
  const float e = ex2_approx(__fmul_rn(__fmul_rn(lambda, __fsub_rn(v, thr)), 1.4426950408889634f));
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fadd_rn(1.0f, e)));
  return r;
}

// per-segment normalisation constants, as dml_ood_keygen applies them: NumPy's fp32 (x - min) / (max - min).
// The divisor is constant per segment, so the IEEE division is evaluated as Markstein's sequence on the correctly
// rounded reciprocal r = RN(1 / den):  q0 = x r;  rem = fma(-q0, den, x) (exact);  q = fma(rem, r, q0)  ==  RN(x / den)
// for 0 <= x <= den -- 3 instructions instead of the ~9 of the division subroutine.  tools/check_div.cu compares the
// two bit for bit over all 2^23 significands of x at 6 exponents for 512 divisors incl. the adversarial ones
// (2.3e10 cases on the B200, 0 mismatches: profiles/r2_check_div.json); results below 1e-30 (where the residual could
// leave the normal range) and a zero / non-finite divisor take the division itself.
struct Norm { float lo, den, r; bool on; };
__device__ __forceinline__ Norm load_norm(const float* __restrict__ minmax, int seg, int slot) {
  Norm n;
  n.on = minmax != nullptr;
  n.lo = 0.f; n.den = 1.f; n.r = 1.f;
  if (n.on) {
    n.lo = minmax[seg * 4 + slot * 2];
    n.den = __fsub_rn(minmax[seg * 4 + slot * 2 + 1], n.lo);
    n.r = __frcp_rn(n.den);
  }
  return n;
}
__device__ __forceinline__ float apply_norm(const Norm& n, float v) {
  if (!n.on) return v;
  const float x = __fsub_rn(v, n.lo);
  const float q0 = __fmul_rn(x, n.r);
  const float q = __fmaf_rn(__fmaf_rn(-q0, n.den, x), n.r, q0);
  // (q0 >= 1e-30 is false for NaN / tiny / negative q0 and for r = inf or NaN, i.e. den = 0)
  return (q0 >= 1.0e-30f && n.r < 3.0e38f) ? q : __fdiv_rn(x, n.den);
}

// The same for the VEC values of a thread with ONE branch: the guard of apply_norm compiled to a compare + branch + reconvergence
// pair per value (5 instructions x 8 sites per pixel quad in rank_kernel); the division fallback is taken only by the
// pixel that sits on the segment's minimum.  Bit-identical to apply_norm element by element.
template <int VEC>
__device__ __forceinline__ void apply_norm_vec(const Norm& n, float (&v)[VEC]) {
  if (!n.on) return;
  float x[VEC], q0[VEC], q[VEC];
  bool ok = n.r < 3.0e38f;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    x[j] = __fsub_rn(v[j], n.lo);
    q0[j] = __fmul_rn(x[j], n.r);
    q[j] = __fmaf_rn(__fmaf_rn(-q0[j], n.den, x[j]), n.r, q0[j]);
    ok = ok && (q0[j] >= 1.0e-30f);
  }
  if (ok) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) v[j] = q[j];
  } else {
#pragma unroll
    for (int j = 0; j < VEC; ++j) v[j] = (q0[j] >= 1.0e-30f && n.r < 3.0e38f) ? q[j] : __fdiv_rn(x[j], n.den);
  }
}

// Positive flags of four ground-truth bytes at once: 0xFF in every byte of the result whose label is in `out_mask`
// (labels 0..63) -- or, with `nonzero`, whose byte is non-zero (a positive mask).  One SIMD byte compare per set label.
// `single4`: the label replicated into four bytes when out_mask holds exactly one label (the usual case: `seg_label == 13`),
// else 0xffffffff -- computed once per kernel (positive_label4), so that the common case is ONE byte compare per word
// instead of a 64-bit bit-scan loop (12 instructions per pixel in the gather kernel before).
__device__ __forceinline__ uint32_t positive_label4(uint64_t out_mask) {
  if (out_mask == 0ull || (out_mask & (out_mask - 1ull)) != 0ull) return 0xffffffffu;
  const int l = __ffsll((long long)out_mask) - 1;          // < 64: never 0xff in a byte
  return (uint32_t)l * 0x01010101u;
}
__device__ __forceinline__ uint32_t positive_bytes(uint32_t w, uint64_t out_mask, bool nonzero, uint32_t single4 = 0xffffffffu) {
  if (nonzero) return __vcmpne4(w, 0u);
  if (single4 != 0xffffffffu) return __vcmpeq4(w, single4);
  uint32_t m = 0u;
  while (out_mask) {                       // warp-uniform: usually one label
    const int l = __ffsll((long long)out_mask) - 1;
    out_mask &= out_mask - 1;
    m |= __vcmpeq4(w, (uint32_t)l * 0x01010101u);
  }
  return m;
}
// bit j = byte j of a byte mask (0x00 / 0xFF per byte)
__device__ __forceinline__ uint32_t byte_mask_to_bits(uint32_t m) { return ((m & 0x01010101u) * 0x01020408u) >> 24; }

// Monotone value-linear index of a score key: q(k) = clamp(int((f(k) - f_lo) * scale), 0, n - 1) with f the float the
// key was packed from.  Subtraction, multiplication by a non-negative constant, truncation and clamping are all
// monotone (non-decreasing) under round-to-nearest, so k1 <= k2  =>  q(k1) <= q(k2): the property every use below
// relies on (bucket sort, lower-bound tables); how evenly q spreads the keys only affects speed.
struct LinIndex {
  float f_lo, scale;
  int n;
  uint32_t key_base;
  __device__ __forceinline__ void init(uint32_t k_lo, uint32_t k_hi, int n_, uint32_t key_base_) {
    n = n_; key_base = key_base_;
    f_lo = key_float(k_lo, key_base);
    const float f_hi = key_float(k_hi, key_base);
    scale = (f_hi > f_lo) ? __fdiv_rn((float)n_, __fsub_rn(f_hi, f_lo)) : 0.f;
    if (!(scale == scale) || scale > 3.0e38f) scale = 0.f;   // degenerate spans: everything in slot 0 (still monotone)
  }
  __device__ __forceinline__ int operator()(uint32_t skey) const {
    const int q = __float2int_rz(__fmul_rn(__fsub_rn(key_float(skey, key_base), f_lo), scale));
    return min(max(q, 0), n - 1);
  }
};

// lower_bound in a sorted shared-memory table of `gn` score keys through a RANK_LUT-entry table:
// lut[q] = first entry whose index is >= q (q = 0 .. RANK_LUT), so the lower bound of a key with index q lies in
// [lut[q], lut[q + 1]] and a few probes finish the search.
struct SmemTable {
  const uint32_t* s;        // [gn] ascending, gn < 65536
  uint32_t* lut;            // [RANK_LUT]: first | (end << 16) of the search range of index q
  int gn;
  LinIndex li;
  static constexpr size_t lut_bytes() { return (size_t)RANK_LUT * sizeof(uint32_t); }
  // all NT threads of the CTA; `s` must be visible (synchronised) before the call; ends with a __syncthreads()
  template <int NT = RANK_THREADS>
  __device__ __forceinline__ void build(const uint32_t* s_, uint32_t* lut_, int gn_, uint32_t key_base) {
    s = s_; lut = lut_; gn = gn_;
    li.init(gn > 0 ? s[0] : 0u, gn > 0 ? s[gn - 1] : 0u, RANK_LUT, key_base);
    // lut[q] = first entry whose index is >= q; the end of the range (first entry with index >= q + 1) is the next
    // slot's start, packed into the upper half afterwards
    for (int q = threadIdx.x; q < RANK_LUT; q += blockDim.x) {
      int lo = 0, hi = gn;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (li(s[mid]) < q) lo = mid + 1; else hi = mid;
      }
      lut[q] = (uint32_t)lo;
    }
    __syncthreads();
    constexpr int PER = RANK_LUT / NT;
    uint32_t nxt[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int q = threadIdx.x + j * NT;
      nxt[j] = q + 1 < RANK_LUT ? lut[q + 1] : (uint32_t)gn;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PER; ++j) lut[threadIdx.x + j * NT] |= nxt[j] << 16;
    __syncthreads();
  }
  __device__ __forceinline__ void range(uint32_t sk, int& lo, int& hi) const {
    const uint32_t e = lut[li(sk)];
    lo = (int)(e & 0xffffu);
    hi = (int)(e >> 16);
  }
  __device__ __forceinline__ int finish(uint32_t sk, int lo, int hi) const {
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s[mid] < sk) lo = mid + 1; else hi = mid;
    }
    return lo;
  }
};

}  // namespace dml
