// Host emulation of the fused multi-scale head's gather (head_kernel<K, HEAD_MS>: ms_gather in csrc/dml_head.cuh)
// pixel by pixel, calling the SAME scalar functions the kernel calls (csrc/dml_bilinear.cuh is __host__ __device__).
// Test infrastructure: built by tests/test_ms_emulation.py with g++ -ffp-contract=off, never linked into the product.
#include <cstdint>

#include "../../open-world-semantic-segmentation_b200/csrc/dml_bilinear.cuh"

using namespace dml;

// z[s]: [B, K, hs[s], ws[s]] fp32; out: [B, K, H, W] = sum_s bilinear_s(z[s]) / n accumulated in scale order
extern "C" void ms_emulate(const float* const* z, const int* hs, const int* ws, int n, int B, int K, int H, int W,
                           int reciprocal_average, float* out) {
  const float div = (float)n, inv = 1.0f / (float)n;
  for (long long i = 0; i < (long long)B * K * H * W; ++i) out[i] = 0.f;
  for (int s = 0; s < n; ++s) {
    // torch area_pixel_compute_scale<float>(in, out, align_corners=false) = float(in) / out  (dml_head.cu)
    const float rh = (float)hs[s] / (float)H, rw = (float)ws[s] / (float)W;
    const long long plane = (long long)hs[s] * ws[s];
    for (int b = 0; b < B; ++b)
      for (int y = 0; y < H; ++y) {
        const float h1r = bilinear_src(rh, y);
        const int h1 = (int)h1r;
        const int dy = (h1 < hs[s] - 1) ? ws[s] : 0;
        const float h1l = h1r - h1, h0l = 1.0f - h1l;
        for (int x = 0; x < W; ++x) {
          const float w1r = bilinear_src(rw, x);
          const int w1 = (int)w1r;
          const int dx = (w1 < ws[s] - 1) ? 1 : 0;
          const float w1l = w1r - w1, w0l = 1.0f - w1l;
          for (int k = 0; k < K; ++k) {
            const float* r = z[s] + ((long long)b * K + k) * plane + (long long)h1 * ws[s] + w1;
            const float val = bilinear_blend(w0l, w1l, h0l, h1l, r[0], r[dx], r[dy], r[dy + dx]);
            float* o = out + (((long long)b * K + k) * H + y) * W + x;
            *o = bl_add(*o, scale_share(val, inv, div, reciprocal_average != 0));
          }
        }
      }
  }
}
