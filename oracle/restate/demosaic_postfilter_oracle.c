/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the guided-Laplacian post-filter of the half-size demosaic.
 * Never linked into, loaded by or called from the product (ansel_b200/); tests/ uses it as the checker.
 *
 * Follows iop/demosaic.c: _downsample_guided_laplacian_fit :681-759, _apply :770-796, _postfilter :810-926 (with
 * DOWNSAMPLE_GUIDED_SCALES :117 = 1); pixel/bspline.h: sparse_scalar_product :83-117, _bspline_vertical_pass :118-133,
 * _bspline_horizontal :136-151, blur_2D_Bspline :330-350, decompose_2D_Bspline :351-377; system/simd.h dt_simd_max_zero :107-114.
 * Pinned: bit-identical to those lines compiled in place (oracle/_ref, ref_demosaic_downsample_postfilter), strict build,
 * tests/test_cpu_ppg.py.
 *
 * One iteration, per a-trous scale s (mult = 2^s): LF = clipped B-spline blur of the current image, HF = (image - LF) / max(LF, 1e-8)
 * per colour; around every pixel a 5x5 patch (clamped at the frame) of HF gives the least-squares line channel = slope * guide +
 * intercept over the guide (R+G+B)/3; slopes and intercepts are blurred (mult 1, unclipped), and the filtered band
 * (slope * guide + intercept) * LF is accumulated.  The iteration's result is max(sum of bands + last LF, 0), non-finite -> 0.
 */
#include "oracle_common.h"
#include <stdlib.h>
#include <string.h>

#define GUIDED_SCALES 1 /* demosaic.c:117 */

static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int iclamp(int v, int lo, int hi) { return v > hi ? hi : (v < lo ? lo : v); } /* glib CLAMP */
static inline float max_zero(float v) { return isfinite(v) ? (v > 0.0f ? v : 0.0f) : 0.f; }

/* blur_2D_Bspline / the LF half of decompose_2D_Bspline: vertical pass into a row buffer, then horizontal; both clip when asked */
static void bspline_blur(const float *in, float *out, int width, int height, int mult, int clip)
{
  static const float f[5] = { 1.0f / 16.0f, 4.0f / 16.0f, 6.0f / 16.0f, 4.0f / 16.0f, 1.0f / 16.0f };
  float *temp = malloc(sizeof(float) * 4 * (size_t)width);
  for(int i = 0; i < height; i++)
  {
    const size_t r[5] = { (size_t)4 * width * imax(i - 2 * mult, 0), (size_t)4 * width * imax(i - mult, 0), (size_t)4 * width * i,
                          (size_t)4 * width * imin(i + mult, height - 1), (size_t)4 * width * imin(i + 2 * mult, height - 1) };
    for(int j = 0; j < width; j++)
      for(int c = 0; c < 4; c++)
      {
        const float *b = in + 4 * (size_t)j + c;
        const float v = f[0] * b[r[0]] + f[1] * b[r[1]] + f[2] * b[r[2]] + f[3] * b[r[3]] + f[4] * b[r[4]];
        temp[4 * j + c] = clip ? (0.0f > v ? 0.0f : v) : v;
      }
    for(int j = 0; j < width; j++)
    {
      const int x[5] = { 4 * imax(j - 2 * mult, 0), 4 * imax(j - mult, 0), 4 * j, 4 * imin(j + mult, width - 1), 4 * imin(j + 2 * mult, width - 1) };
      float *o = out + 4 * ((size_t)i * width + j);
      for(int c = 0; c < 4; c++)
      {
        const float v = f[0] * temp[x[0] + c] + f[1] * temp[x[1] + c] + f[2] * temp[x[2] + c] + f[3] * temp[x[3] + c] + f[4] * temp[x[4] + c];
        o[c] = clip ? (0.0f > v ? 0.0f : v) : v;
      }
    }
  }
  free(temp);
}

/* :681-759 */
static void fit(const float *HF, float *coeff, float *bias, int width, int height)
{
  const float inv_patch = 1.f / 25.f;
  for(int row = 0; row < height; row++)
    for(int col = 0; col < width; col++)
    {
      float sum_rgb[4] = { 0.f }, sum_rgb_guide[4] = { 0.f }, sum_guide = 0.f, sum_guide_sq = 0.f;
      for(int jj = -2; jj <= 2; jj++)
        for(int ii = -2; ii <= 2; ii++)
        {
          const float *s = HF + 4 * ((size_t)iclamp(row + jj, 0, height - 1) * width + iclamp(col + ii, 0, width - 1));
          const float guide = (s[0] + s[1] + s[2]) / 3.f;
          for(int c = 0; c < 4; c++) sum_rgb[c] += s[c];
          sum_guide += guide;
          sum_guide_sq += guide * guide;
          for(int c = 0; c < 4; c++) sum_rgb_guide[c] += s[c] * guide;
        }
      const float guide_mean = sum_guide * inv_patch;
      float variance = sum_guide_sq * inv_patch - guide_mean * guide_mean;
      if(variance < 0.f) variance = 0.f;
      float *k = coeff + 4 * ((size_t)row * width + col), *b = bias + 4 * ((size_t)row * width + col);
      for(int c = 0; c < 3; c++)
      {
        const float mean = sum_rgb[c] * inv_patch;
        const float covariance = sum_rgb_guide[c] * inv_patch - mean * guide_mean;
        const float slope = variance > 1e-12f ? covariance / variance : 0.f;
        k[c] = slope;
        b[c] = mean - slope * guide_mean;
      }
      k[3] = b[3] = 0.f;
    }
}

/* rgba: width * height * 4 floats, filtered in place; iterations = data->color_smoothing (demosaic.c:1108) */
int orc_demosaic_downsample_postfilter(float *out, int width, int height, int iterations)
{
  if(iterations <= 0) return 0;
  orc_fp_fast_mode(); /* the pipe's threads run with FTZ|DAZ (darktable.c:877) */
  const size_t px = (size_t)width * height;
  float *LF[2] = { malloc(16 * px), malloc(16 * px) }, *HF = malloc(16 * px), *rec = malloc(16 * px), *coeff = malloc(16 * px), *bias = malloc(16 * px),
        *tmp = malloc(16 * px);
  for(int it = 0; it < iterations; it++)
  {
    const float *residual = out;
    for(int s = 0; s < GUIDED_SCALES; s++)
    {
      const float *bin = s == 0 ? out : (s % 2 ? LF[1] : LF[0]); /* LF[1] = LF_odd */
      float *bout = (s == 0 || s % 2 == 0) ? LF[1] : LF[0];
      bspline_blur(bin, bout, width, height, 1 << s, 1);
      for(size_t k = 0; k < px; k++)
      {
        for(int c = 0; c < 3; c++) HF[4 * k + c] = (bin[4 * k + c] - bout[4 * k + c]) / fmaxf(bout[4 * k + c], 1e-8f);
        HF[4 * k + 3] = 0.f;
      }
      fit(HF, coeff, bias, width, height);
      bspline_blur(coeff, tmp, width, height, 1, 0);
      memcpy(coeff, tmp, 16 * px);
      bspline_blur(bias, tmp, width, height, 1, 0);
      memcpy(bias, tmp, 16 * px);
      for(size_t k = 0; k < px; k++)
      {
        const float *h = HF + 4 * k;
        const float guide = (h[0] + h[1] + h[2]) / 3.f;
        for(int c = 0; c < 3; c++)
        {
          const float filtered = (coeff[4 * k + c] * guide + bias[4 * k + c]) * bout[4 * k + c];
          rec[4 * k + c] = s == 0 ? filtered : filtered + rec[4 * k + c];
        }
        rec[4 * k + 3] = 0.f;
      }
      residual = bout;
    }
    for(size_t k = 0; k < px; k++)
    {
      for(int c = 0; c < 3; c++) out[4 * k + c] = max_zero(rec[4 * k + c] + residual[4 * k + c]);
      out[4 * k + 3] = 0.f;
    }
  }
  free(LF[0]), free(LF[1]), free(HF), free(rec), free(coeff), free(bias), free(tmp);
  return 0;
}
