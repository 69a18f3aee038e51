// VNG4 demosaic and the dual-demosaic blend (sharp demosaicer where there is detail, VNG4 in flat areas).
//
// Reference: iop/demosaic/basic.c lin_interpolate :20-125; iop/demosaic/vng.c vng_interpolate :33-202; iop/demosaic/dual.c
// dual_demosaic :39-112; iop/demosaic.c intp :250-257; develop/masks/detail.c dt_masks_calc_rawdetail_mask :282-317,
// dt_masks_calc_detail_mask :327-337, dt_masks_blur_9x9 :214-234, dt_masks_extend_border :91-120.
//
// The reference runs VNG in place over the bilinear image behind a three-row ring buffer, which only serves to make every
// pixel read bilinear values: VNG is a pure function of the bilinear image within +-2 pixels.  Here: one kernel writes the
// four-colour bilinear image (a pure function of the mosaic within +-1), one kernel per pixel walks the 64 gradient terms
// of the pattern and averages the neighbours under the threshold.  Every border fill of the mask code ("extend the
// border by copying the nearest interior value") is the stencil evaluated at clamped coordinates, so the detail mask is
// three pointwise/stencil kernels without separate border passes.  Operation order is the reference's: bit-identical.
#ifndef B200_KERNELS_ON_CPU // tests/emul compiles the kernels of this file with g++ to check them against the oracle without a GPU
#include "runtime.h"
#endif
#include <limits.h>
#include <math.h>
#include <string.h>

namespace
{
constexpr int VNT = 128;

// the CFA seen by fcol (develop/imageop_math.h:207-214): a Bayer word (the four-colour one, second green = 3), or 9 and the
// sensor's 6x6 X-Trans table
struct cfa_t
{
  unsigned filters;
  unsigned char xtrans[36];
};
__device__ __forceinline__ int vfc(int row, int col, const cfa_t &F)
{
  if(F.filters == 9u) return F.xtrans[((row + 600) % 6) * 6 + (col + 600) % 6];
  return (F.filters >> ((((row << 1) & 14) + (col & 1)) << 1)) & 3;
}
__device__ __forceinline__ float lane(const float4 &p, int c) { return c == 0 ? p.x : (c == 1 ? p.y : (c == 2 ? p.z : p.w)); }
__device__ __forceinline__ void set_lane(float4 &p, int c, float v)
{
  if(c == 0) p.x = v;
  else if(c == 1) p.y = v;
  else if(c == 2) p.z = v;
  else p.w = v;
}

// bilinear interpolation with four colours (the second green is colour 3), basic.c:20-125
__global__ void __launch_bounds__(VNT) lin_interpolate_kernel(const float *__restrict__ in, float4 *__restrict__ out, int width, int height, int x0, int y0,
                                                              const cfa_t filters4)
{
  const int col = blockIdx.x * VNT + threadIdx.x, row = blockIdx.y;
  if(col >= width) return;
  const int colors = filters4.filters == 9u ? 3 : 4; // X-Trans: lane 3 is not a colour and is not written (kept by vng_kernel)
  const int f = vfc(row + y0, col + x0, filters4);
  const float own = in[(size_t)row * width + col];
  float4 o;
  if(row == 0 || col == 0 || row == height - 1 || col == width - 1)
  { // the outermost pixels: average of the neighbours of each colour inside the frame, raster order (:28-55)
    float sum[4] = { 0.f, 0.f, 0.f, 0.f };
    int count[4] = { 0, 0, 0, 0 };
    for(int y = row - 1; y != row + 2; y++)
      for(int x = col - 1; x != col + 2; x++)
        if(y >= 0 && x >= 0 && y < height && x < width)
        {
          const int c = vfc(y + y0, x + x0, filters4);
          const float v = in[(size_t)y * width + x];
#pragma unroll
          for(int q = 0; q < 4; q++)
            if(q == c)
            {
              sum[q] += v;
              count[q]++;
            }
        }
#pragma unroll
    for(int c = 0; c < 4; c++) set_lane(o, c, c >= colors ? 0.0f : ((c != f && count[c] != 0) ? sum[c] / (float)count[c] : own));
  }
  else
  { // weights 1 (diagonal), 2 (edge), 4 (never: the centre is the pixel's own colour); :68-124
    float sum[4] = { 0.f, 0.f, 0.f, 0.f };
    int tot[4] = { 0, 0, 0, 0 };
    for(int y = -1; y <= 1; y++)
      for(int x = -1; x <= 1; x++)
      {
        const int weight = 1 << ((y == 0) + (x == 0));
        const int c = vfc(row + y + y0, col + x + x0, filters4);
        if(c == f) continue;
        const float v = in[(size_t)(row + y) * width + col + x] * (float)weight;
#pragma unroll
        for(int q = 0; q < 4; q++)
          if(q == c)
          {
            sum[q] += v;
            tot[q] += weight;
          }
      }
#pragma unroll
    for(int c = 0; c < 4; c++) set_lane(o, c, c >= colors ? 0.0f : (c == f ? own : sum[c] / (float)tot[c]));
  }
  out[(size_t)row * width + col] = o;
}

// the gradient terms of dcraw's VNG: y1, x1, y2, x2, weight, gradient mask (vng.c:38-54), and the eight neighbours (:55-56)
__constant__ signed char c_terms[64 * 6] = {
  -2, -2, +0, -1, 1, 0x01, -2, -2, +0, +0, 2, 0x01, -2, -1, -1, +0, 1, 0x01, -2, -1, +0, -1, 1, 0x02, -2, -1, +0, +0, 1, 0x03, -2, -1, +0, +1, 2, 0x01,
  -2, +0, +0, -1, 1, 0x06, -2, +0, +0, +0, 2, 0x02, -2, +0, +0, +1, 1, 0x03, -2, +1, -1, +0, 1, 0x04, -2, +1, +0, -1, 2, 0x04, -2, +1, +0, +0, 1, 0x06,
  -2, +1, +0, +1, 1, 0x02, -2, +2, +0, +0, 2, 0x04, -2, +2, +0, +1, 1, 0x04, -1, -2, -1, +0, 1, (signed char)0x80, -1, -2, +0, -1, 1, 0x01, -1, -2, +1, -1, 1, 0x01,
  -1, -2, +1, +0, 2, 0x01, -1, -1, -1, +1, 1, (signed char)0x88, -1, -1, +1, -2, 1, 0x40, -1, -1, +1, -1, 1, 0x22, -1, -1, +1, +0, 1, 0x33, -1, -1, +1, +1, 2, 0x11,
  -1, +0, -1, +2, 1, 0x08, -1, +0, +0, -1, 1, 0x44, -1, +0, +0, +1, 1, 0x11, -1, +0, +1, -2, 2, 0x40, -1, +0, +1, -1, 1, 0x66, -1, +0, +1, +0, 2, 0x22,
  -1, +0, +1, +1, 1, 0x33, -1, +0, +1, +2, 2, 0x10, -1, +1, +1, -1, 2, 0x44, -1, +1, +1, +0, 1, 0x66, -1, +1, +1, +1, 1, 0x22, -1, +1, +1, +2, 1, 0x10,
  -1, +2, +0, +1, 1, 0x04, -1, +2, +1, +0, 2, 0x04, -1, +2, +1, +1, 1, 0x04, +0, -2, +0, +0, 2, (signed char)0x80, +0, -1, +0, +1, 2, (signed char)0x88, +0, -1, +1, -2, 1, 0x40,
  +0, -1, +1, +0, 1, 0x11, +0, -1, +2, -2, 1, 0x40, +0, -1, +2, -1, 1, 0x20, +0, -1, +2, +0, 1, 0x30, +0, -1, +2, +1, 2, 0x10, +0, +0, +0, +2, 2, 0x08,
  +0, +0, +2, -2, 2, 0x40, +0, +0, +2, -1, 1, 0x60, +0, +0, +2, +0, 2, 0x20, +0, +0, +2, +1, 1, 0x30, +0, +0, +2, +2, 2, 0x10, +0, +1, +1, +0, 1, 0x44,
  +0, +1, +1, +2, 1, 0x10, +0, +1, +2, -1, 2, 0x40, +0, +1, +2, +0, 1, 0x60, +0, +1, +2, +1, 1, 0x20, +0, +1, +2, +2, 1, 0x10, +1, -2, +1, +0, 1, (signed char)0x80,
  +1, -1, +1, +1, 1, (signed char)0x88, +1, +0, +1, +2, 1, 0x08, +1, +0, +2, -1, 1, 0x40, +1, +0, +2, +1, 1, 0x10
};
__constant__ signed char c_chood[16] = { -1, -1, -1, 0, -1, +1, 0, +1, +1, +1, +1, 0, +1, -1, 0, -1 };

// vng.c:77-186 as a function of the bilinear image; the two greens are averaged on the way out (:193-197), the fourth
// lane keeps the second green like the reference's buffer does
__global__ void __launch_bounds__(VNT) vng_kernel(const float4 *__restrict__ lin, float4 *__restrict__ out, int width, int height, int x0, int y0,
                                                  const cfa_t filters4)
{
  const int col = blockIdx.x * VNT + threadIdx.x, row = blockIdx.y;
  if(col >= width) return;
  const bool xtrans = filters4.filters == 9u;
  const int colors = xtrans ? 3 : 4, period_row = xtrans ? 6 : 8, period_col = xtrans ? 6 : 2;
  const float4 *pix = lin + (size_t)row * width + col;
  float4 o = pix[0];
  if(row >= 2 && col >= 2 && row < height - 2 && col < width - 2)
  {
    const int prow = (row + y0) % period_row, pcol = (col + x0) % period_col;
    float gval[8] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
    for(int t = 0; t < 64; t++)
    {
      const int y1 = c_terms[6 * t], x1 = c_terms[6 * t + 1], y2 = c_terms[6 * t + 2], x2 = c_terms[6 * t + 3], weight = c_terms[6 * t + 4];
      const int grads = c_terms[6 * t + 5] & 0xff;
      const int color = vfc(prow + y1, pcol + x1, filters4);
      if(vfc(prow + y2, pcol + x2, filters4) != color) continue;
      const int diag = (vfc(prow, pcol + 1, filters4) == color && vfc(prow + 1, pcol, filters4) == color) ? 2 : 1;
      if(abs(y1 - y2) == diag && abs(x1 - x2) == diag) continue;
      const float diff = fabsf(lane(pix[y1 * width + x1], color) - lane(pix[y2 * width + x2], color)) * (float)weight;
#pragma unroll
      for(int g = 0; g < 8; g++)
        if(grads & (1 << g)) gval[g] += diff;
    }
    float gmin = gval[0], gmax = gval[0];
#pragma unroll
    for(int g = 1; g < 8; g++)
    {
      if(gmin > gval[g]) gmin = gval[g];
      if(gmax < gval[g]) gmax = gval[g];
    }
    if(gmax != 0)
    {
      const float thold = gmin + (gmax * 0.5f);
      float sum[4] = { 0.f, 0.f, 0.f, 0.f };
      const int color = vfc(row + y0, col + x0, filters4);
      const float own = lane(o, color);
      int num = 0;
#pragma unroll
      for(int g = 0; g < 8; g++)
      {
        const int y = c_chood[2 * g], x = c_chood[2 * g + 1];
        if(gval[g] <= thold)
        {
          const float4 near = pix[y * width + x];
          const bool far = vfc(prow + y, pcol + x, filters4) != color && vfc(prow + y * 2, pcol + x * 2, filters4) == color;
          const float mid = far ? (own + lane(pix[(y * width + x) * 2], color)) * 0.5f : 0.f;
#pragma unroll
          for(int c = 0; c < 4; c++)
            if(c < colors) sum[c] += (c == color && far) ? mid : lane(near, c);
          num++;
        }
      }
      float base = 0.f;
#pragma unroll
      for(int c = 0; c < 4; c++)
        if(c == color) base = sum[c];
#pragma unroll
      for(int c = 0; c < 4; c++)
        if(c < colors)
        {
          float tot = own;
          if(c != color) tot += (sum[c] - base) / (float)num;
          set_lane(o, c, tot);
        }
    }
  }
  if(xtrans)
    o.w = out[(size_t)row * width + col].w; // the reference leaves lane 3 to whatever its buffers held: kept as found
  else
    o.y = (o.y + o.w) / 2.0f;
  out[(size_t)row * width + col] = o;
}

// ---- the detail mask of the dual demosaic -----------------------------------------------------------------------------------
// dt_masks_calc_rawdetail_mask, first loop :288-293: the white-balance-neutral luminance with a square-root gamma
__global__ void __launch_bounds__(256) detail_luma_kernel(const float4 *__restrict__ rgb, float *__restrict__ tmp, size_t n, float wb0, float wb1, float wb2)
{
  const size_t k = (size_t)blockIdx.x * 256 + threadIdx.x;
  if(k >= n) return;
  const float4 p = rgb[k];
  const float val = 0.333333333f * (fmaxf(p.x, 0.0f) / wb0 + fmaxf(p.y, 0.0f) / wb1 + fmaxf(p.z, 0.0f) / wb2);
  tmp[k] = sqrtf(val);
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ float fast_expf(float x)
{ // dt_fast_expf, math/math.h:254-267: int + float * int is float arithmetic; (int) of it truncates, INT_MIN out of range
  const float f = (float)0x3f800000 + x * (float)(0x402DF854 - 0x3f800000);
  const int k0 = (f >= -2147483648.0f && f < 2147483648.0f) ? (int)f : INT_MIN;
  return __uint_as_float((unsigned)(k0 > 0 ? k0 : 0));
}
// Scharr gradient magnitude :296-315 with its one-pixel border extension, then calcBlendFactor :319-325 (:331-335)
__global__ void __launch_bounds__(VNT) detail_sigmoid_kernel(const float *__restrict__ tmp, float *__restrict__ out, int width, int height, float threshold)
{
  const int col = blockIdx.x * VNT + threadIdx.x, row = blockIdx.y;
  if(col >= width) return;
  const int r = clampi(row, 1, height - 2), c = clampi(col, 1, width - 2);
  const int idx = r * width + c;
  const float gx = 47.0f * (tmp[idx - width - 1] - tmp[idx - width + 1]) + 162.0f * (tmp[idx - 1] - tmp[idx + 1])
                   + 47.0f * (tmp[idx + width - 1] - tmp[idx + width + 1]);
  const float gy = 47.0f * (tmp[idx - width - 1] - tmp[idx + width - 1]) + 162.0f * (tmp[idx - width] - tmp[idx + width])
                   + 47.0f * (tmp[idx - width + 1] - tmp[idx + width + 1]);
  const float gxs = gx / 256.0f, gys = gy / 256.0f;
  const float mask = (1.0f / 16.0f) * sqrtf(gxs * gxs + gys * gys);
  out[(size_t)row * width + col] = 1.0f / (1.0f + fast_expf(16.0f - (16.0f / threshold) * mask));
}
struct blur9_t
{
  float c[13];
};
// dt_masks_blur_9x9 :214-234 with its four-pixel border extension; the sum in the order of FAST_BLUR_9 :194-208
__global__ void __launch_bounds__(VNT) detail_blur_kernel(const float *__restrict__ src, float *__restrict__ out, int width, int height, const blur9_t B)
{
  const int col = blockIdx.x * VNT + threadIdx.x, row = blockIdx.y;
  if(col >= width) return;
  const int i = clampi(row, 4, height - 5) * width + clampi(col, 4, width - 5);
  const int w1 = width, w2 = 2 * width, w3 = 3 * width, w4 = 4 * width;
  const float v =
      B.c[12] * (src[i - w4 - 2] + src[i - w4 + 2] + src[i - w2 - 4] + src[i - w2 + 4] + src[i + w2 - 4] + src[i + w2 + 4] + src[i + w4 - 2] + src[i + w4 + 2])
      + B.c[11] * (src[i - w4 - 1] + src[i - w4 + 1] + src[i - w1 - 4] + src[i - w1 + 4] + src[i + w1 - 4] + src[i + w1 + 4] + src[i + w4 - 1] + src[i + w4 + 1])
      + B.c[10] * (src[i - w4] + src[i - 4] + src[i + 4] + src[i + w4])
      + B.c[9] * (src[i - w3 - 3] + src[i - w3 + 3] + src[i + w3 - 3] + src[i + w3 + 3])
      + B.c[8] * (src[i - w3 - 2] + src[i - w3 + 2] + src[i - w2 - 3] + src[i - w2 + 3] + src[i + w2 - 3] + src[i + w2 + 3] + src[i + w3 - 2] + src[i + w3 + 2])
      + B.c[7] * (src[i - w3 - 1] + src[i - w3 + 1] + src[i - w1 - 3] + src[i - w1 + 3] + src[i + w1 - 3] + src[i + w1 + 3] + src[i + w3 - 1] + src[i + w3 + 1])
      + B.c[6] * (src[i - w3] + src[i - 3] + src[i + 3] + src[i + w3])
      + B.c[5] * (src[i - w2 - 2] + src[i - w2 + 2] + src[i + w2 - 2] + src[i + w2 + 2])
      + B.c[4] * (src[i - w2 - 1] + src[i - w2 + 1] + src[i - w1 - 2] + src[i - w1 + 2] + src[i + w1 - 2] + src[i + w1 + 2] + src[i + w2 - 1] + src[i + w2 + 1])
      + B.c[3] * (src[i - w2] + src[i - 2] + src[i + 2] + src[i + w2])
      + B.c[2] * (src[i - w1 - 1] + src[i - w1 + 1] + src[i + w1 - 1] + src[i + w1 + 1])
      + B.c[1] * (src[i - w1] + src[i - 1] + src[i + 1] + src[i + w1])
      + B.c[0] * src[i];
  out[(size_t)row * width + col] = fminf(1.0f, fmaxf(0.0f, v));
}
// dual.c:97-103: intp(blend, sharp, vng) = blend * (sharp - vng) + vng on all four lanes
__global__ void __launch_bounds__(256) dual_blend_kernel(float4 *__restrict__ rgb, const float4 *__restrict__ vng, const float *__restrict__ blend, size_t n)
{
  const size_t k = (size_t)blockIdx.x * 256 + threadIdx.x;
  if(k >= n) return;
  const float a = blend[k];
  const float4 s = rgb[k], v = vng[k];
  rgb[k] = make_float4(a * (s.x - v.x) + v.x, a * (s.y - v.y) + v.y, a * (s.z - v.z) + v.z, a * (s.w - v.w) + v.w);
}

// ---- host-side set-up shared with tests/emul ----------------------------------------------------------------------------------
inline unsigned four_colour_word(unsigned filters) { return (filters & 3) == 1 ? (filters | 0x03030303u) : (filters | 0x0c0c0c0cu); } // vng.c:68-73
inline cfa_t make_cfa(unsigned filters, const unsigned char *xtrans36)
{
  cfa_t F;
  memset(&F, 0, sizeof(F));
  F.filters = filters == 9u ? 9u : four_colour_word(filters);
  if(filters == 9u && xtrans36) memcpy(F.xtrans, xtrans36, 36);
  return F;
}
// dt_masks_blur_9x9_coeff :159-192 (host libm expf, as in the reference)
inline void blur9_coeff(blur9_t *B, float sigma)
{
  float kernel[9][9];
  const float temp = -2.0f * (sigma * sigma), range = (3.0f * 1.5f) * (3.0f * 1.5f);
  float sum = 0.0f;
  for(int k = -4; k <= 4; k++)
    for(int j = -4; j <= 4; j++)
    {
      const float d = (float)k * (float)k + (float)j * (float)j;
      kernel[k + 4][j + 4] = d <= range ? expf(d / temp) : 0.0f;
      if(d <= range) sum += kernel[k + 4][j + 4];
    }
  for(int i = 0; i < 9; i++)
    for(int j = 0; j < 9; j++) kernel[i][j] /= sum;
  const float pick[13] = { kernel[4][4], kernel[3][4], kernel[3][3], kernel[2][4], kernel[2][3], kernel[2][2], kernel[1][4],
                           kernel[1][3], kernel[1][2], kernel[1][1], kernel[0][4], kernel[0][3], kernel[0][2] };
  memcpy(B->c, pick, sizeof(pick));
}
inline float slider2contrast(float slider) { return 0.005f * powf(slider, 1.1f); } // dual.c:34-37
} // namespace

#ifndef B200_KERNELS_ON_CPU
namespace b200
{
int demosaic_color_smoothing_dev(float *d_out, int width, int height, int passes, cudaStream_t s);

// vng_interpolate(out, in, roo, roi, piece->dsc_in.filters, ..., FALSE), demosaic.c:1172-1175: filters is the sensor word,
// the ROI origin enters through x0 / y0.  lin_slot: the scratch slot for the bilinear image.
int vng_demosaic_dev(const float *d_in, float *d_out, int width, int height, int x0, int y0, uint32_t filters, const uint8_t *xtrans36, int lin_slot,
                     cudaStream_t s)
{
  if(width < 5 || height < 5 || height > 65535) return fail(B200_ERR_UNSUPPORTED, "demosaic: VNG4 on a %dx%d frame", width, height);
  void *lin = nullptr;
  int rc = scratch(lin_slot, (size_t)width * height * 16, &lin);
  if(rc) return rc;
  const dim3 grid((unsigned)((width + VNT - 1) / VNT), (unsigned)height);
  const cfa_t f4 = make_cfa(filters, xtrans36);
  lin_interpolate_kernel<<<grid, VNT, 0, s>>>(d_in, (float4 *)lin, width, height, x0, y0, f4);
  B200_CUDA_TRY(cudaGetLastError());
  vng_kernel<<<grid, VNT, 0, s>>>((const float4 *)lin, (float4 *)d_out, width, height, x0, y0, f4);
  B200_CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

// dual_demosaic(), dual.c:39-112 (without the GUI's mask display): d_rgb holds the sharp demosaicer's frame and is blended in place
int dual_demosaic_dev(float *d_rgb, const float *d_raw, int width, int height, int x0, int y0, uint32_t filters, const float wb[4], float dual_threshold,
                      cudaStream_t s)
{
  if(width < 16 || height < 16 || !(dual_threshold > 0.0f)) return B200_OK; // :49, :52
  if(height > 65535) return fail(B200_ERR_ARG, "demosaic: frame height %d", height);
  const size_t n = (size_t)width * height;
  void *vng = nullptr, *tmp = nullptr, *blend = nullptr;
  int rc;
  if((rc = scratch(SLOT_TMP1, n * 16, &vng))) return rc;
  if((rc = scratch(SLOT_TMP2, n * 4, &tmp))) return rc;
  if((rc = scratch(SLOT_TMP3, n * 4, &blend))) return rc;
  if((rc = vng_demosaic_dev(d_raw, (float *)vng, width, height, x0, y0, filters, nullptr, SLOT_TMP0, s))) return rc;
  if((rc = demosaic_color_smoothing_dev((float *)vng, width, height, 2, s))) return rc;
  const dim3 grid((unsigned)((width + VNT - 1) / VNT), (unsigned)height);
  detail_luma_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const float4 *)d_rgb, (float *)tmp, n, wb[0], wb[1], wb[2]);
  detail_sigmoid_kernel<<<grid, VNT, 0, s>>>((const float *)tmp, (float *)blend, width, height, slider2contrast(dual_threshold));
  blur9_t B;
  blur9_coeff(&B, 2.0f);
  detail_blur_kernel<<<grid, VNT, 0, s>>>((const float *)blend, (float *)tmp, width, height, B);
  dual_blend_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((float4 *)d_rgb, (const float4 *)vng, (const float *)tmp, n);
  B200_CUDA_TRY(cudaGetLastError());
  return B200_OK;
}
} // namespace b200
#endif
