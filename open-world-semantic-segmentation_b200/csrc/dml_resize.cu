// Input side of the multi-scale evaluation: the resize + normalisation loop of
//   anomaly/dataset.py:281-297   for this_short_size in imgSizes: imresize(img, (w, h), 'bilinear') -> img_transform
//   anomaly/dataset.py:11-21     imresize = PIL.Image.resize(size, Image.BILINEAR)
//   anomaly/dataset.py:65-70     img_transform: float32(img) / 255, HWC -> CHW, Normalize(mean, std)
// on the GPU: one launch per scale reads the decoded uint8 RGB image (2.8 MB at 720 x 1280, L2-resident across the five
// scales) and writes the normalised float32 CHW tensor the backbone consumes -- bit for bit what PIL + NumPy + torchvision
// produce on the CPU (10 - 20 ms per image there).
//
// The resampling arithmetic is Pillow's (a dependency of the reference, not part of /root/reference; algorithm of
// src/libImaging/Resample.c, stable since Pillow 3.x, checked here against Pillow 12.2): two separable passes over 8-bit
// data with integer coefficients.  Per axis and output index xx (scale = in / out, support = max(scale, 1) for the
// triangle filter):  center = (xx + 0.5) scale;  taps [xmin, xmax) = [int(center - support + 0.5), int(center + support +
// 0.5)) clamped to the image;  w(x) = max(0, 1 - |x + 0.5 - center| / max(scale, 1)), normalised to sum 1 in double, then
// k = int(0.5 + w 2^22).  A pass computes clip8((2^21 + sum_x pixel(x) k(x)) >> 22); the horizontal pass runs first and its
// uint8 result feeds the vertical pass.  The tables are built on the host in double (same operation order as Pillow);
// the kernel does the two passes through shared memory for one 8 x 64 output tile (the horizontal blends of the tile's
// input rows are computed once, not once per output row) and applies the normalisation with IEEE divisions in the
// reference's order:  ((v / 255) - mean) / std.
#include <cmath>
#include <vector>
#include "dml_common.cuh"

namespace dml {
namespace {

constexpr int RS_PRECISION_BITS = 32 - 8 - 2;
constexpr int RS_TX = 64, RS_TY = 8, RS_THREADS = 256;

int resize_ksize(int in_size, int out_size) {
  double filterscale = (double)in_size / (double)out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 1.0 * filterscale;          // triangle filter: support 1
  return (int)std::ceil(support) * 2 + 1;
}

// bounds[2 xx] = first tap, bounds[2 xx + 1] = number of taps; kk[xx * ksize + x] = integer weight of tap x
void resize_coeffs(int in_size, int out_size, int ksize, int32_t* bounds, int32_t* kk) {
  const double scale = (double)in_size / (double)out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  const double ss = 1.0 / filterscale;
  std::vector<double> k((size_t)ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      double t = (x + xmin - center + 0.5) * ss;
      if (t < 0.0) t = -t;
      const double w = t < 1.0 ? 1.0 - t : 0.0;
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (int x = xmax; x < ksize; ++x) k[x] = 0.0;
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
    for (int x = 0; x < ksize; ++x)
      kk[(size_t)xx * ksize + x] = k[x] < 0 ? (int)(-0.5 + k[x] * (1 << RS_PRECISION_BITS)) : (int)(0.5 + k[x] * (1 << RS_PRECISION_BITS));
  }
}

__device__ __forceinline__ int clip8(int v) {
  v >>= RS_PRECISION_BITS;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

struct ResizeArgs {
  const uint8_t* img;      // [B, H, W, 3]
  float* out;              // [B, 3, oh, ow]
  const int32_t *bx, *kx, *by, *ky;
  int H, W, oh, ow, ksx, ksy, rows_cap;
  float mean[3], stdv[3];
};

__global__ void __launch_bounds__(RS_THREADS) resize_norm_kernel(const ResizeArgs a) {
  extern __shared__ __align__(16) uint8_t s_h[];     // [rows][RS_TX] uchar4: the horizontally resampled input rows of the tile
  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int tx0 = blockIdx.x * RS_TX, ty0 = blockIdx.y * RS_TY;
  const int ty1 = min(ty0 + RS_TY, a.oh) - 1;
  const int y_lo = a.by[2 * ty0];
  const int n_rows = a.by[2 * ty1] + a.by[2 * ty1 + 1] - y_lo;      // the tap windows move monotonically with the output row
  const uint8_t* img = a.img + (size_t)b * a.H * a.W * 3;
  for (int i = tid; i < n_rows * RS_TX; i += RS_THREADS) {
    const int r = i / RS_TX, c = i - r * RS_TX;
    const int xx = tx0 + c;
    if (xx >= a.ow) continue;
    const int xmin = a.bx[2 * xx], cnt = a.bx[2 * xx + 1];
    const int32_t* k = a.kx + (size_t)xx * a.ksx;
    const uint8_t* src = img + ((size_t)(y_lo + r) * a.W + xmin) * 3;
    int s0 = 1 << (RS_PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < cnt; ++x) {
      const int kv = __ldg(k + x);
      s0 += (int)__ldg(src + 3 * x) * kv;
      s1 += (int)__ldg(src + 3 * x + 1) * kv;
      s2 += (int)__ldg(src + 3 * x + 2) * kv;
    }
    reinterpret_cast<uchar4*>(s_h)[i] = make_uchar4((unsigned char)clip8(s0), (unsigned char)clip8(s1), (unsigned char)clip8(s2), 0);
  }
  __syncthreads();
  const size_t plane = (size_t)a.oh * a.ow;
  float* out = a.out + (size_t)b * 3 * plane;
  for (int i = tid; i < RS_TY * RS_TX; i += RS_THREADS) {
    const int ly = i / RS_TX, c = i - ly * RS_TX;
    const int yy = ty0 + ly, xx = tx0 + c;
    if (yy >= a.oh || xx >= a.ow) continue;
    const int ymin = a.by[2 * yy] - y_lo, cnt = a.by[2 * yy + 1];
    const int32_t* k = a.ky + (size_t)yy * a.ksy;
    int s0 = 1 << (RS_PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int y = 0; y < cnt; ++y) {
      const int kv = __ldg(k + y);
      const uchar4 p = reinterpret_cast<const uchar4*>(s_h)[(ymin + y) * RS_TX + c];
      s0 += (int)p.x * kv;
      s1 += (int)p.y * kv;
      s2 += (int)p.z * kv;
    }
    const int v[3] = {clip8(s0), clip8(s1), clip8(s2)};
    float* o = out + (size_t)yy * a.ow + xx;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
      o[ch * plane] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v[ch], 255.f), a.mean[ch]), a.stdv[ch]);
  }
}

}  // namespace
}  // namespace dml

using namespace dml;

extern "C" {
#pragma GCC visibility push(default)

int32_t dml_resize_ksize(int32_t in_size, int32_t out_size) {
  if (in_size < 1 || out_size < 1) return 0;
  return resize_ksize(in_size, out_size);
}

int dml_resize_coeffs(int32_t in_size, int32_t out_size, int32_t* bounds, int32_t* coeffs, int32_t ksize) {
  if (in_size < 1 || out_size < 1 || !bounds || !coeffs || ksize != resize_ksize(in_size, out_size)) return DML_ERR_INVALID_ARG;
  resize_coeffs(in_size, out_size, ksize, bounds, coeffs);
  return DML_OK;
}

int dml_resize_bilinear_normalize(const uint8_t* image, int32_t B, int32_t H, int32_t W, const int32_t* bounds_x, const int32_t* coeffs_x,
                                  int32_t ksize_x, const int32_t* bounds_y, const int32_t* coeffs_y, int32_t ksize_y, int32_t out_h,
                                  int32_t out_w, int32_t max_rows_per_tile, const float* mean, const float* stdv, float* out,
                                  dml_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!image || !bounds_x || !coeffs_x || !bounds_y || !coeffs_y || !mean || !stdv || !out || B < 0 || H < 1 || W < 1 || out_h < 1 ||
      out_w < 1 || ksize_x < 1 || ksize_y < 1 || max_rows_per_tile < 1 || B > 65535)
    return DML_ERR_INVALID_ARG;
  const size_t smem = (size_t)max_rows_per_tile * RS_TX * 4;
  if (smem > 200 * 1024) return DML_ERR_INVALID_ARG;          // a reduction this strong needs Pillow's reducing_gap route first
  if (B == 0) return DML_OK;
  ResizeArgs a;
  a.img = image; a.out = out; a.bx = bounds_x; a.kx = coeffs_x; a.by = bounds_y; a.ky = coeffs_y;
  a.H = H; a.W = W; a.oh = out_h; a.ow = out_w; a.ksx = ksize_x; a.ksy = ksize_y; a.rows_cap = max_rows_per_tile;
  for (int c = 0; c < 3; ++c) { a.mean[c] = mean[c]; a.stdv[c] = stdv[c]; }
  if (smem > 48 * 1024) DML_CUDA_TRY(cudaFuncSetAttribute(resize_norm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((out_w + RS_TX - 1) / RS_TX), (unsigned)((out_h + RS_TY - 1) / RS_TY), (unsigned)B);
  resize_norm_kernel<<<grid, RS_THREADS, smem, stream>>>(a);
  DML_LAUNCH_CHECK();
  return DML_OK;
}

#pragma GCC visibility pop
}  // extern "C"
