/* CPU restatement of finalscale: the export's final resampling.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src: iop/finalscale.c process() :117-131 -> develop/imageop_math.c dt_iop_clip_and_zoom_roi
 * :146-152 -> pixel/interpolation.c: _clip :88-142 (replicate), the tap generators _maketaps_bilinear :175-194,
 * _maketaps_bicubic :200-233, _maketaps_mitchell :253-287 (four taps per run, the running position advanced by
 * 4 * interval per run), _compute_upsampling_kernel :320-342, _compute_downsampling_kernel :354-387 (ceil_fast:
 * math/math.h:324-334), _prepare_resampling_plan :710-893, _interpolation_resample_plain :897-1027 (dt_simd_max_zero:
 * system/simd.h:108-114).  Pinned bit-for-bit against those lines cut verbatim (oracle/_ref: ref_finalscale.c).
 */
#include "oracle_common.h"
#include "b200iop.h"
#include <stdlib.h>
#include <string.h>

static float ceil_fast(float x) { return x <= 0.f ? (float)(int)x : -((float)(int)-x) + 1.f; }

/* taps[0 .. 4*ceil(num/4)): lane k of run r sits at ((first + k * interval) + 4 * interval) + ... r times */
static void maketaps(int interpolator, float *taps, size_t num_taps, float first_tap, float interval)
{
  const float iter = 4.0f * interval;
  float vt[4];
  for(int k = 0; k < 4; k++) vt[k] = first_tap + (float)k * interval;
  const int runs = (int)((num_taps + 3) / 4);
  for(int r = 0; r < runs; r++)
  {
    for(int k = 0; k < 4; k++)
    {
      const float t = vt[k], a = fabsf(t);
      float v;
      if(interpolator == B200_INTERPOLATION_BILINEAR)
        v = 1.0f - a;
      else if(interpolator == B200_INTERPOLATION_BICUBIC)
      {
        const float t2 = t * t, t5 = 5.0f * a;
        const float r12 = (a * (t5 - 8.0f - t2) + 4.0f) * 0.5f;
        const float r01 = ((3.0f * t2 - t5) * a + 2.0f) * 0.5f;
        v = a <= 1.0f ? r01 : r12;
      }
      else
      {
        const float a2 = a * a, a3 = a2 * a;
        const float r01 = (7.0f / 6.0f) * a3 - 2.0f * a2 + (8.0f / 9.0f);
        const float r12 = 2.0f * a2 - (7.0f / 18.0f) * a3 - (10.0f / 3.0f) * a + (16.0f / 9.0f);
        v = a <= 1.0f ? r01 : r12;
      }
      taps[4 * r + k] = v;
      vt[k] += iter;
    }
  }
}
static int half_width(int interpolator) { return interpolator == B200_INTERPOLATION_BILINEAR ? 1 : 2; }

/* One axis of the plan.  lengths[out]; kernel/index: concatenated taps (at most max_taps).  Returns the number of taps,
 * -1 when scale == 1 (no plan), -3 when max_taps is too small. */
int orc_resampling_plan(int interpolator, int in, int in_x0, int out, int out_x0, float scale, int *lengths, float *kernel, int *index, int max_taps)
{
  if(scale == 1.f) return -1;
  const int width = half_width(interpolator);
  const int maxtapsapixel = scale > 1.f ? 2 * width : (int)ceil_fast((float)2 * (float)width / scale);
  float *scratch = malloc(sizeof(float) * (maxtapsapixel + 8));
  int n = 0;
  for(int x = 0; x < out; x++)
  {
    int first, taps;
    if(scale > 1.f)
    {
      const float fx = (float)(out_x0 + x) / scale - in_x0;
      first = (int)floorf(fx) - width + 1;
      taps = 2 * width;
      maketaps(interpolator, scratch, taps, fx - (float)first, -1.0f);
    }
    else
    {
      const float w = (float)width;
      const float xin = ceil_fast(((float)(out_x0 + x) - w) / scale);
      first = (int)xin;
      const float t = xin * scale - (float)(out_x0 + x);
      taps = (int)((w - t) / scale);
      maketaps(interpolator, scratch, taps, t, scale);
    }
    /* BORDER_REPLICATE: every tap is kept, indexes are clipped into the line */
    if(n + taps > max_taps)
    {
      free(scratch);
      return -3;
    }
    lengths[x] = taps;
    float norm = 0.f;
    for(int tap = 0; tap < taps; tap++) norm += scratch[tap];
    norm = 1.f / norm;
    for(int tap = 0; tap < taps; tap++)
    {
      kernel[n] = scratch[tap] * norm;
      const int i = first + tap;
      index[n++] = i < 0 ? 0 : (i > in - 1 ? in - 1 : i);
    }
  }
  free(scratch);
  return n;
}

int orc_clip_and_zoom(const float *in, float *out, int in_x, int in_y, int in_w, int in_h, double in_scale, int out_x, int out_y, int out_w, int out_h,
                      double out_scale, int interpolator);
/* finalscale's process(): the origins of both ROIs are zeroed, sizes and scales kept */
int orc_finalscale(const float *in, float *out, int in_w, int in_h, double in_scale, int out_w, int out_h, double out_scale, int interpolator)
{
  return orc_clip_and_zoom(in, out, 0, 0, in_w, in_h, in_scale, 0, 0, out_w, out_h, out_scale, interpolator);
}
/* dt_iop_clip_and_zoom_roi -> _interpolation_resample_plain :897-1027 with the ROIs as given (initialscale's process(),
 * iop/initialscale.c:122-129): the ROI origins enter the tap plans, and the 1:1 path copies from (roi_out - roi_in) */
int orc_clip_and_zoom(const float *in, float *out, int in_x, int in_y, int in_w, int in_h, double in_scale, int out_x, int out_y, int out_w, int out_h,
                      double out_scale, int interpolator)
{
  if(out_scale == 1.f || out_scale == in_scale)
  {
    for(int y = 0; y < out_h; y++)
      memcpy(out + (size_t)4 * out_w * y, in + (size_t)4 * in_w * (y + (out_y - in_y)) + (size_t)4 * (out_x - in_x), sizeof(float) * 4 * out_w);
    return 0;
  }
  const float resample_scale = out_scale / in_scale; /* a double division, rounded to float */
  const int width = half_width(interpolator);
  const int per = resample_scale > 1.f ? 2 * width : (int)ceil_fast((float)2 * (float)width / resample_scale) + 1;
  int *hl = malloc(sizeof(int) * out_w), *vl = malloc(sizeof(int) * out_h);
  int *hi = malloc(sizeof(int) * (size_t)per * out_w), *vi = malloc(sizeof(int) * (size_t)per * out_h);
  float *hk = malloc(sizeof(float) * (size_t)per * out_w), *vk = malloc(sizeof(float) * (size_t)per * out_h);
  const int nh = orc_resampling_plan(interpolator, in_w, in_x, out_w, out_x, resample_scale, hl, hk, hi, per * out_w);
  const int nv = orc_resampling_plan(interpolator, in_h, in_y, out_h, out_y, resample_scale, vl, vk, vi, per * out_h);
  int rc = (nh < 0 || nv < 0) ? 1 : 0;
  if(!rc)
  {
    size_t voff = 0;
    for(int oy = 0; oy < out_h; oy++)
    {
      size_t hoff = 0;
      for(int ox = 0; ox < out_w; ox++)
      {
        float vs[4] = { 0.f, 0.f, 0.f, 0.f };
        for(int iy = 0; iy < vl[oy]; iy++)
        {
          const float *line = in + (size_t)vi[voff + iy] * in_w * 4;
          float vhs[4] = { 0.f, 0.f, 0.f, 0.f };
          for(int ix = 0; ix < hl[ox]; ix++)
          {
            const float *p = line + (size_t)hi[hoff + ix] * 4;
            const float htap = hk[hoff + ix];
            for(int c = 0; c < 4; c++) vhs[c] += p[c] * htap;
          }
          const float vtap = vk[voff + iy];
          for(int c = 0; c < 4; c++) vs[c] += vhs[c] * vtap;
        }
        float *o = out + ((size_t)oy * out_w + ox) * 4;
        for(int c = 0; c < 4; c++) o[c] = isfinite(vs[c]) ? (vs[c] > 0.0f ? vs[c] : 0.0f) : 0.f;
        hoff += hl[ox];
      }
      voff += vl[oy];
    }
  }
  free(hl), free(vl), free(hi), free(vi), free(hk), free(vk);
  return rc;
}

/* flip: dt_imageio_flip_buffers, imageio/imageio_core.c:258-297 as iop/flip.c process() :388-400 calls it (bpp bytes per pixel,
 * input rows of `width` pixels; orientation: 1 flip y, 2 flip x, 4 swap x and y -- the output then has rows of `height` pixels) */
int orc_flip(const void *in, void *out, int bpp, int width, int height, int orientation)
{
  const char *src = in;
  char *dst = out;
  long ii = 0, jj = 0, si = bpp, sj = (long)width * bpp;
  if(orientation & 4)
  {
    sj = bpp;
    si = (long)height * bpp;
  }
  if(orientation & 1)
  {
    jj = height - 1;
    sj = -sj;
  }
  if(orientation & 2)
  {
    ii = width - 1;
    si = -si;
  }
  for(int j = 0; j < height; j++)
    for(int i = 0; i < width; i++)
      memcpy(dst + labs(sj) * jj + labs(si) * ii + sj * j + si * i, src + ((size_t)j * width + i) * bpp, bpp);
  return 0;
}
