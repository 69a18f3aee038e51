/* CPU restatement of the non-local-means core as the reference runs it.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src/pixel/nlmeans_core.c: scatter :95-104, define_patches :107-145,
 * pixel_difference :156-165, diff_of_pixels_diff :168-180, init_column_sums :214-264,
 * compute_slice_height/width :267-312, nlmeans_denoise :315-532; dt_fast_mexp2f math/math.h:290-301;
 * and the callers iop/denoiseprofile.c process_nlmeans_cpu :1599-1648 (nlmeans_norm :1456-1470,
 * nlmeans_scattering :1474-1499) and iop/nlmeans.c process_cpu :416-456.
 *
 * The result is order dependent by construction: patch distances are running column sums updated
 * incrementally down the rows of a ~60x72 chunk and a running sum along each row, all in float.  The
 * chunk geometry and every accumulation order are therefore restated exactly.  Pinned bit-for-bit
 * against the reference file compiled in place (oracle/_ref, ref_nlm.c).
 */
#include "oracle_common.h"
#include "flt32_math.h"
#include "b200iop.h"
#include <stdlib.h>
#include <string.h>

#define SLICE_WIDTH 72  /* nlmeans_core.c:55 */
#define SLICE_HEIGHT 60 /* :56 */
#define IMIN(a, b) ((a) < (b) ? (a) : (b))
#define IMAX(a, b) ((a) > (b) ? (a) : (b))

typedef struct
{
  short rows, cols;
  int offset;
} patch_t;

/* math/math.h:290-301 */
static inline float fast_mexp2(float x)
{
  const int i1 = 0x3f800000, i2 = 0x3f000000;
  const int k0 = i1 + (int)(x * (i2 - i1));
  const int k = k0 >= 0x800000 ? k0 : 0;
  float f;
  memcpy(&f, &k, 4);
  return f;
}
static inline int sgn(int a) { return (a > 0) - (a < 0); }
/* :95-104: evaluated in double, truncated to int */
static int scatter(float scale, float scattering, int i1, int i2)
{
  const int a1 = abs(i1), a2 = abs(i2);
  return (int)(scale * ((a1 * a1 * a1 + 7.0 * a1 * sqrt((double)a2)) * sgn(i1) * scattering / 6.0 + i1));
}
static inline float pixdiff(const float *p1, const float *p2, const float norm[4])
{ /* :156-165 */
  float s[4];
  for(int i = 0; i < 4; i++)
  {
    const float d = p1[i] - p2[i];
    s[i] = d * d * norm[i];
  }
  return s[0] + s[1] + s[2];
}
static inline float diff_of_diffs(const float *p1, const float *p2, const float *p3, const float *p4, const float norm[4])
{ /* :168-180 */
  float s[4];
  for(int i = 0; i < 4; i++)
  {
    const float d1 = p1[i] - p2[i], d2 = p3[i] - p4[i];
    s[i] = (d1 * d1 - d2 * d2) * norm[i];
  }
  return s[0] + s[1] + s[2];
}
int orc_nlm_slice_height(int height)
{ /* :267-296 */
  if(height % SLICE_HEIGHT == 0) return SLICE_HEIGHT;
  int best = height % SLICE_HEIGHT, best_incr = 0;
  for(int incr = 1; incr < 10; incr++)
  {
    const int plus_rem = height % (SLICE_HEIGHT + incr);
    if(plus_rem == 0) return SLICE_HEIGHT + incr;
    if(plus_rem > best)
    {
      best_incr = +incr;
      best = plus_rem;
    }
    const int minus_rem = height % (SLICE_HEIGHT - incr);
    if(minus_rem == 0) return SLICE_HEIGHT - incr;
    if(minus_rem > best)
    {
      best_incr = -incr;
      best = minus_rem;
    }
  }
  return SLICE_HEIGHT + best_incr;
}
int orc_nlm_slice_width(int width)
{ /* :299-312 */
  int sl = SLICE_WIDTH;
  int rem = width % sl;
  if(rem < SLICE_WIDTH / 2 && (width % (sl - 4)) > rem)
  {
    sl -= 4;
    rem = width % sl;
    if(rem < SLICE_WIDTH / 2 && (width % (sl - 4)) > rem) sl -= 4;
  }
  return sl;
}

/* nlmeans_denoise(), :315-532.  norm: four per-channel weights.  Returns 0. */
int orc_nlmeans_denoise(const float *inbuf, float *outbuf, int width, int height, float scattering, float scale, float luma,
                        float chroma, float center_weight, float sharpness, int radius, int search_radius, int decimate,
                        const float norm[4])
{
  const float weight[4] = { luma, chroma, chroma, 1.0f };
  const float invert[4] = { 1.0f - luma, 1.0f - chroma, 1.0f - chroma, 0.0f };
  const int skip_blend = (luma == 1.0 && chroma == 1.0);
  const int pw = 2 * radius + 1;
  const float cp_norm = center_weight * pw * pw; /* compute_center_pixel_norm :147-153 */
  const float center_norm[4] = { cp_norm, cp_norm, cp_norm, 1.0f };
  const int stride = 4 * width;

  int n_patches = (2 * search_radius + 1) * (2 * search_radius + 1);
  if(decimate) n_patches = (n_patches + 1) / 2;
  patch_t *patches = malloc(sizeof(patch_t) * (size_t)n_patches);
  if(!patches) return 1;
  {
    int k = 0, dec = decimate;
    for(int ri = -search_radius; ri <= search_radius; ri++)
      for(int ci = -search_radius; ci <= search_radius; ci++)
      {
        if(dec && (++dec & 1)) continue;
        const int r = scatter(scale, scattering, ri, ci), c = scatter(scale, scattering, ci, ri);
        patches[k].rows = (short)r;
        patches[k].cols = (short)c;
        patches[k].offset = r * stride + c * 4;
        k++;
      }
  }
  const int chk_h = orc_nlm_slice_height(height), chk_w = orc_nlm_slice_width(width);
  const int n_ct = (height + chk_h - 1) / chk_h, n_cl = (width + chk_w - 1) / chk_w;

#pragma omp parallel for schedule(dynamic) collapse(2)
  for(int it = 0; it < n_ct; it++)
    for(int il = 0; il < n_cl; il++)
    {
      orc_fp_fast_mode(); /* the pipe's threads run with FTZ|DAZ (darktable.c:877, common/dtpthread.c:54): a weighted sum next to FLT_MIN flushes */
      const int chunk_top = it * chk_h, chunk_left = il * chk_w;
      float scratch[SLICE_WIDTH + 2 * 8 + 1 + 48];
      float *const col_sums = scratch + (radius + 1) - chunk_left;
      const int chunk_bot = IMIN(chunk_top + chk_h, height), chunk_right = IMIN(chunk_left + chk_w, width);
      for(int i = chunk_top; i < chunk_bot; i++)
        memset(outbuf + 4 * ((size_t)i * width + chunk_left), 0, sizeof(float) * 4 * (size_t)(chunk_right - chunk_left));
      for(int p = 0; p < n_patches; p++)
      {
        const patch_t *patch = &patches[p];
        const int srow = patch->rows, scol = patch->cols, offset = patch->offset;
        const int row_min = IMAX(chunk_top, IMAX(0, -srow)), row_max = IMIN(chunk_bot, height - IMAX(0, srow));
        const int row_top = IMAX(row_min, IMAX(radius, radius - srow));
        const int row_bot = IMIN(row_max, height - 1 - IMAX(radius, radius + srow));
        const int col_min = IMAX(chunk_left, -scol), col_max = IMIN(chunk_right, width - scol);
        const int pcol_min = chunk_left - IMIN(radius, IMIN(chunk_left, chunk_left + scol));
        const int pcol_max = chunk_right + IMIN(radius, IMIN(width - chunk_right, width - (chunk_right + scol)));
        { /* init_column_sums(), :214-264, at row = row_min */
          const int row = row_min;
          const int rmin = row - IMIN(radius, IMIN(row, row + srow));
          const int rmax = row + IMIN(radius, IMIN(height - 1 - row, height - 1 - (row + srow)));
          for(int col = chunk_left - radius - 1; col < IMIN(pcol_min, chunk_right + radius); col++) col_sums[col] = 0;
          for(int col = pcol_min; col < pcol_max; col++)
          {
            float sum = 0;
            for(int r = rmin; r <= rmax; r++)
            {
              const float *px = inbuf + (size_t)r * stride + 4 * col;
              sum += pixdiff(px, px + offset, norm);
            }
            col_sums[col] = sum;
          }
          for(int col = IMAX(pcol_min, pcol_max); col < chunk_right + radius; col++) col_sums[col] = 0;
        }
        for(int row = row_min; row < row_max; row++)
        {
          float distortion = 0.0;
          for(int i = col_min - radius; i < IMIN(col_min + radius, col_max); i++) distortion += col_sums[i];
          const float *in = inbuf + (size_t)stride * row;
          float *const out = outbuf + (size_t)4 * width * row;
          if(center_weight < 0)
          { /* denoise (non-local means) iop, :389-402 */
            for(int col = col_min; col < col_max; col++)
            {
              distortion += (col_sums[col + radius] - col_sums[col - radius - 1]);
              const float wt = fast_mexp2(distortion * sharpness);
              const float *const inpx = in + 4 * col;
              const float pixel[4] = { inpx[offset], inpx[offset + 1], inpx[offset + 2], 1.0f };
              for(int c = 0; c < 4; c++) out[4 * col + c] += pixel[c] * wt;
            }
          }
          else
          { /* denoise (profiled), :404-420 */
            for(int col = col_min; col < col_max; col++)
            {
              distortion += (col_sums[col + radius] - col_sums[col - radius - 1]);
              const float dissimilarity = (distortion + pixdiff(in + 4 * col, in + 4 * col + offset, center_norm)) / (1.0f + center_weight);
              const float wt = fast_mexp2(fmaxf(0.0f, dissimilarity * sharpness - 2.0f));
              const float *const inpx = in + 4 * col;
              const float pixel[4] = { inpx[offset], inpx[offset + 1], inpx[offset + 2], 1.0f };
              for(int c = 0; c < 4; c++) out[4 * col + c] += pixel[c] * wt;
            }
          }
          if(row < IMIN(row_top, row_bot))
          { /* :424-440 */
            const float *bot_row = inbuf + (size_t)(row + 1 + radius) * stride;
            for(int col = pcol_min; col < pcol_max; col++)
            {
              const float *const b = bot_row + 4 * col;
              col_sums[col] += pixdiff(b, b + offset, norm);
            }
          }
          else if(row < row_bot)
          { /* :441-466 */
            const float *const top_row = inbuf + (size_t)(row - radius) * stride;
            const float *const bot_row = inbuf + (size_t)(row + 1 + radius) * stride;
            for(int col = pcol_min; col < pcol_max; col++)
            {
              const float *const t = top_row + 4 * col, *const b = bot_row + 4 * col;
              col_sums[col] += diff_of_diffs(b, b + offset, t, t + offset, norm);
            }
          }
          else if(row >= row_top && row + 1 < row_max)
          { /* :467-483 */
            const float *top_row = inbuf + (size_t)(row - radius) * stride;
            for(int col = pcol_min; col < pcol_max; col++)
            {
              const float *const t = top_row + 4 * col;
              col_sums[col] -= pixdiff(t, t + offset, norm);
            }
          }
        }
      }
      if(skip_blend)
      { /* :487-500 */
        for(int row = chunk_top; row < chunk_bot; row++)
        {
          float *const out = outbuf + (size_t)4 * row * width;
          for(int col = chunk_left; col < chunk_right; col++)
            for(int c = 0; c < 4; c++) out[4 * col + c] /= out[4 * col + 3];
        }
      }
      else
      { /* :502-517 */
        for(int row = chunk_top; row < chunk_bot; row++)
        {
          const float *in = inbuf + (size_t)row * stride;
          float *out = outbuf + (size_t)row * 4 * width;
          for(int col = chunk_left; col < chunk_right; col++)
            for(int c = 0; c < 4; c++) out[4 * col + c] = (in[4 * col + c] * invert[c]) + (out[4 * col + c] / out[4 * col + 3] * weight[c]);
        }
      }
    }
  free(patches);
  return 0;
}
