/* CPU restatement of the VNG4 demosaicer and of the dual-demosaic blend.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src: iop/demosaic/basic.c lin_interpolate :20-125; iop/demosaic/vng.c vng_interpolate :33-202
 * (Bayer: four colours, the second green separated as colour 3 and averaged back at the end); iop/demosaic/dual.c
 * dual_demosaic :39-112 (slider2contrast :34-37); iop/demosaic.c intp :250-257; develop/masks/detail.c
 * dt_masks_extend_border :91-120, dt_masks_blur_9x9_coeff :159-192, dt_masks_blur_9x9 :214-234,
 * dt_masks_calc_rawdetail_mask :282-317, calcBlendFactor :319-325 (dt_fast_expf: math/math.h:254-267),
 * dt_masks_calc_detail_mask :327-337; color_smoothing: demosaic_extra_oracle.c.  Pinned bit-for-bit against those lines
 * cut verbatim (oracle/_ref: ref_vng.c).
 *
 * The reference runs VNG in place over the bilinear image through a three-row ring buffer, so that every pixel reads
 * bilinear values only: here the bilinear image is kept and the result written elsewhere, which is the same function.
 */
#include "oracle_common.h"
#include "flt32_math.h"
#include <limits.h>
#include <stdlib.h>
#include <string.h>

void orc_color_smoothing(float *out, int width, int height, int passes);

static const signed char vng_terms[] = {
  -2, -2, +0, -1, 1, 0x01, -2, -2, +0, +0, 2, 0x01, -2, -1, -1, +0, 1, 0x01, -2, -1, +0, -1, 1, 0x02, -2, -1, +0, +0, 1, 0x03, -2, -1, +0, +1, 2, 0x01,
  -2, +0, +0, -1, 1, 0x06, -2, +0, +0, +0, 2, 0x02, -2, +0, +0, +1, 1, 0x03, -2, +1, -1, +0, 1, 0x04, -2, +1, +0, -1, 2, 0x04, -2, +1, +0, +0, 1, 0x06,
  -2, +1, +0, +1, 1, 0x02, -2, +2, +0, +0, 2, 0x04, -2, +2, +0, +1, 1, 0x04, -1, -2, -1, +0, 1, 0x80, -1, -2, +0, -1, 1, 0x01, -1, -2, +1, -1, 1, 0x01,
  -1, -2, +1, +0, 2, 0x01, -1, -1, -1, +1, 1, 0x88, -1, -1, +1, -2, 1, 0x40, -1, -1, +1, -1, 1, 0x22, -1, -1, +1, +0, 1, 0x33, -1, -1, +1, +1, 2, 0x11,
  -1, +0, -1, +2, 1, 0x08, -1, +0, +0, -1, 1, 0x44, -1, +0, +0, +1, 1, 0x11, -1, +0, +1, -2, 2, 0x40, -1, +0, +1, -1, 1, 0x66, -1, +0, +1, +0, 2, 0x22,
  -1, +0, +1, +1, 1, 0x33, -1, +0, +1, +2, 2, 0x10, -1, +1, +1, -1, 2, 0x44, -1, +1, +1, +0, 1, 0x66, -1, +1, +1, +1, 1, 0x22, -1, +1, +1, +2, 1, 0x10,
  -1, +2, +0, +1, 1, 0x04, -1, +2, +1, +0, 2, 0x04, -1, +2, +1, +1, 1, 0x04, +0, -2, +0, +0, 2, 0x80, +0, -1, +0, +1, 2, 0x88, +0, -1, +1, -2, 1, 0x40,
  +0, -1, +1, +0, 1, 0x11, +0, -1, +2, -2, 1, 0x40, +0, -1, +2, -1, 1, 0x20, +0, -1, +2, +0, 1, 0x30, +0, -1, +2, +1, 2, 0x10, +0, +0, +0, +2, 2, 0x08,
  +0, +0, +2, -2, 2, 0x40, +0, +0, +2, -1, 1, 0x60, +0, +0, +2, +0, 2, 0x20, +0, +0, +2, +1, 1, 0x30, +0, +0, +2, +2, 2, 0x10, +0, +1, +1, +0, 1, 0x44,
  +0, +1, +1, +2, 1, 0x10, +0, +1, +2, -1, 2, 0x40, +0, +1, +2, +0, 1, 0x60, +0, +1, +2, +1, 1, 0x20, +0, +1, +2, +2, 1, 0x10, +1, -2, +1, +0, 1, 0x80,
  +1, -1, +1, +1, 1, 0x88, +1, +0, +1, +2, 1, 0x08, +1, +0, +2, -1, 1, 0x40, +1, +0, +2, +1, 1, 0x10
};
static const signed char vng_chood[] = { -1, -1, -1, 0, -1, +1, 0, +1, +1, +1, +1, 0, +1, -1, 0, -1 };

static uint32_t four_colour_word(uint32_t filters) { return (filters & 3) == 1 ? (filters | 0x03030303u) : (filters | 0x0c0c0c0cu); }
/* fcol, develop/imageop_math.h:207-214: Bayer word or the 6x6 X-Trans table (FCxtrans :197-204 without a roi) */
static const uint8_t (*g_xtrans)[6]; /* set for the duration of a call when filters == 9; the oracle is single-threaded here */
static int fcolour(int row, int col, uint32_t filters) { return filters == 9u ? g_xtrans[(row + 600) % 6][(col + 600) % 6] : orc_fc(row, col, filters); }
#define orc_fc(r, c, f) fcolour((r), (c), (f))

/* bilinear interpolation with four colours, basic.c:20-125; (x, y): roi_in origin */
void orc_lin_interpolate(float *out, const float *in, int width, int height, int x0, int y0, uint32_t filters4)
{
  const int colors = filters4 == 9u ? 3 : 4, size = filters4 == 9u ? 6 : 16;
  for(int row = 0; row < height; row++)
    for(int col = 0; col < width; col++)
    {
      if(col == 1 && row >= 1 && row < height - 1) col = width - 1;
      float sum[4] = { 0.0f };
      uint8_t count[4] = { 0 };
      for(int y = row - 1; y != row + 2; y++)
        for(int x = col - 1; x != col + 2; x++)
          if(y >= 0 && x >= 0 && y < height && x < width)
          {
            const int f = orc_fc(y + y0, x + x0, filters4);
            sum[f] += in[y * width + x];
            count[f]++;
          }
      const int f = orc_fc(row + y0, col + x0, filters4);
      for(int c = 0; c < colors; c++)
        out[4 * (row * width + col) + c] = (c != f && count[c] != 0) ? sum[c] / count[c] : in[row * width + col];
    }
  for(int row = 1; row < height - 1; row++)
    for(int col = 1; col < width - 1; col++)
    {
      float sum[4] = { 0.0f };
      int tot[4] = { 0 };
      const int f = orc_fc(row % size + y0, col % size + x0, filters4);
      for(int y = -1; y <= 1; y++)
        for(int x = -1; x <= 1; x++)
        {
          const int weight = 1 << ((y == 0) + (x == 0));
          const int color = orc_fc(row % size + y + y0, col % size + x + x0, filters4);
          if(color == f) continue;
          sum[color] += in[(row + y) * width + col + x] * weight;
          tot[color] += weight;
        }
      float *buf = out + 4 * (row * width + col);
      /* the table lists the colours other than f in ascending order and leaves out the last one the loop counter skips */
      int written = 0;
      for(int c = 0; c < colors && written < colors - 1; c++)
        if(c != f)
        {
          buf[c] = sum[c] / tot[c];
          written++;
        }
      buf[f] = in[row * width + col];
    }
}

/* VNG proper, vng.c:77-186, as a function of the bilinear image `lin`: out = lin on the two-pixel border */
static void vng_from_linear(float *out, const float *lin, int width, int height, int x0, int y0, uint32_t filters4)
{
  const int colors = filters4 == 9u ? 3 : 4, period_row = filters4 == 9u ? 6 : 8, period_col = filters4 == 9u ? 6 : 2;
  memcpy(out, lin, sizeof(float) * 4 * width * height);
  for(int row = 2; row < height - 2; row++)
    for(int col = 2; col < width - 2; col++)
    {
      const int prow = (row + y0) % period_row, pcol = (col + x0) % period_col;
      const float *pix = lin + 4 * (row * width + col);
      float gval[8] = { 0.0f };
      const signed char *cp = vng_terms;
      for(int t = 0; t < 64; t++)
      {
        const int y1 = *cp++, x1 = *cp++, y2 = *cp++, x2 = *cp++, weight = *cp++, grads = *cp++;
        const int color = orc_fc(prow + y1, pcol + x1, filters4);
        if(orc_fc(prow + y2, pcol + x2, filters4) != color) continue;
        const int diag = (orc_fc(prow, pcol + 1, filters4) == color && orc_fc(prow + 1, pcol, filters4) == color) ? 2 : 1;
        if(abs(y1 - y2) == diag && abs(x1 - x2) == diag) continue;
        const float diff = fabsf(pix[(y1 * width + x1) * 4 + color] - pix[(y2 * width + x2) * 4 + color]) * weight;
        for(int g = 0; g < 8; g++)
          if(grads & 1 << g) gval[g] += diff;
      }
      float gmin = gval[0], gmax = gval[0];
      for(int g = 1; g < 8; g++)
      {
        if(gmin > gval[g]) gmin = gval[g];
        if(gmax < gval[g]) gmax = gval[g];
      }
      if(gmax == 0) continue; /* the bilinear pixel stays */
      const float thold = gmin + (gmax * 0.5f);
      float sum[4] = { 0.0f };
      const int color = orc_fc(row + y0, col + x0, filters4);
      int num = 0;
      cp = vng_chood;
      for(int g = 0; g < 8; g++)
      {
        const int y = *cp++, x = *cp++;
        const int near = (y * width + x) * 4;
        const int far = (orc_fc(prow + y, pcol + x, filters4) != color && orc_fc(prow + y * 2, pcol + x * 2, filters4) == color)
                            ? (y * width + x) * 8 + color : 0;
        if(gval[g] <= thold)
        {
          for(int c = 0; c < colors; c++)
            if(c == color && far)
              sum[c] += (pix[c] + pix[far]) * 0.5f;
            else
              sum[c] += pix[near + c];
          num++;
        }
      }
      float *o = out + 4 * (row * width + col);
      for(int c = 0; c < colors; c++)
      {
        float tot = pix[color];
        if(c != color) tot += (sum[c] - sum[color]) / num;
        o[c] = tot;
      }
    }
}

/* filters: the sensor word (not ROI-shifted; the origin enters through x0, y0 as in the reference) */
int orc_vng_interpolate_xtrans(float *out, const float *in, int width, int height, int x0, int y0, const uint8_t xtrans[36], int only_linear);
int orc_vng_interpolate(float *out, const float *in, int width, int height, int x0, int y0, uint32_t filters, int only_linear)
{
  const uint32_t filters4 = four_colour_word(filters);
  if(only_linear)
  {
    orc_lin_interpolate(out, in, width, height, x0, y0, filters4);
    return 0;
  }
  float *lin = malloc(sizeof(float) * 4 * width * height);
  orc_lin_interpolate(lin, in, width, height, x0, y0, filters4);
  vng_from_linear(out, lin, width, height, x0, y0, filters4);
  free(lin);
  for(int i = 0; i < height * width; i++) out[i * 4 + 1] = (out[i * 4 + 1] + out[i * 4 + 3]) / 2.0f;
  return 0;
}

/* ---- the detail mask of the dual demosaic ---------------------------------------------------------------------------- */
static float sqf(float x) { return x * x; }
static void extend_border(float *mask, int width, int height, int border)
{
  const int max_col = width - border - 1;
  for(int row = border; row < height - border; row++)
  {
    float *r = mask + (size_t)(row * width);
    for(int i = 0; i < border; i++)
    {
      r[i] = r[border];
      r[width - i - 1] = r[max_col];
    }
  }
  const float *top_row = mask + (size_t)(border * width), *bot_row = mask + (size_t)(height - border - 1) * width;
  for(int col = 0; col < width; col++)
  {
    const int c = col < border ? border : (col > max_col ? max_col : col);
    const float top = top_row[c], bot = bot_row[c];
    for(int i = 0; i < border; i++)
    {
      mask[col + i * width] = top;
      mask[col + (height - i - 1) * width] = bot;
    }
  }
}
void orc_blur_9x9_coeff(float *c, float sigma)
{
  float kernel[9][9];
  const float temp = -2.0f * sqf(sigma), range = sqf(3.0f * 1.5f);
  float sum = 0.0f;
  for(int k = -4; k <= 4; k++)
    for(int j = -4; j <= 4; j++)
    {
      if((sqf(k) + sqf(j)) <= range)
      {
        kernel[k + 4][j + 4] = f32m_expf((sqf(k) + sqf(j)) / temp);
        sum += kernel[k + 4][j + 4];
      }
      else
        kernel[k + 4][j + 4] = 0.0f;
    }
  for(int i = 0; i < 9; i++)
    for(int j = 0; j < 9; j++) kernel[i][j] /= sum;
  const float pick[13] = { kernel[4][4], kernel[3][4], kernel[3][3], kernel[2][4], kernel[2][3], kernel[2][2], kernel[1][4],
                           kernel[1][3], kernel[1][2], kernel[1][1], kernel[0][4], kernel[0][3], kernel[0][2] };
  memcpy(c, pick, sizeof(pick));
}
static float fast_expf(float x)
{ /* math/math.h:254-267: int + float * int is float arithmetic, converted back with truncation (INT_MIN when out of range) */
  const int i1 = 0x3f800000, i2 = 0x402DF854;
  const float f = (float)i1 + x * (float)(i2 - i1);
  const int k0 = (f >= -2147483648.0f && f < 2147483648.0f) ? (int)f : INT_MIN;
  const int k = k0 > 0 ? k0 : 0;
  float r;
  memcpy(&r, &k, 4);
  return r;
}
/* blend[width*height]: the mask dual_demosaic mixes with (1 = keep the sharp demosaicer) */
void orc_dual_mask(const float *rgb, float *blend, int width, int height, const float wb[3], float dual_threshold)
{
  const int msize = width * height;
  float *tmp = malloc(sizeof(float) * msize);
  const float contrastf = 0.005f * f32m_powf(dual_threshold, 1.1f);
  for(int idx = 0; idx < msize; idx++)
  {
    const float val = 0.333333333f * (fmaxf(rgb[4 * idx], 0.0f) / wb[0] + fmaxf(rgb[4 * idx + 1], 0.0f) / wb[1] + fmaxf(rgb[4 * idx + 2], 0.0f) / wb[2]);
    tmp[idx] = sqrtf(val);
  }
  const float scale = 1.0f / 16.0f;
  for(int row = 1; row < height - 1; row++)
    for(int col = 1, idx = row * width + col; col < width - 1; col++, idx++)
    {
      const float gx = 47.0f * (tmp[idx - width - 1] - tmp[idx - width + 1]) + 162.0f * (tmp[idx - 1] - tmp[idx + 1])
                       + 47.0f * (tmp[idx + width - 1] - tmp[idx + width + 1]);
      const float gy = 47.0f * (tmp[idx - width - 1] - tmp[idx + width - 1]) + 162.0f * (tmp[idx - width] - tmp[idx + width])
                       + 47.0f * (tmp[idx - width + 1] - tmp[idx + width + 1]);
      blend[idx] = scale * sqrtf(sqf(gx / 256.0f) + sqf(gy / 256.0f));
    }
  extend_border(blend, width, height, 1);
  for(int idx = 0; idx < msize; idx++) tmp[idx] = 1.0f / (1.0f + fast_expf(16.0f - (16.0f / contrastf) * blend[idx]));
  float blurmat[13];
  orc_blur_9x9_coeff(blurmat, 2.0f);
  const int w1 = width, w2 = 2 * width, w3 = 3 * width, w4 = 4 * width;
  const float *src = tmp;
  for(int row = 4; row < height - 4; row++)
    for(int col = 4; col < width - 4; col++)
    {
      const int i = row * width + col;
      const float v =
          blurmat[12] * (src[i - w4 - 2] + src[i - w4 + 2] + src[i - w2 - 4] + src[i - w2 + 4] + src[i + w2 - 4] + src[i + w2 + 4] + src[i + w4 - 2] + src[i + w4 + 2])
          + blurmat[11] * (src[i - w4 - 1] + src[i - w4 + 1] + src[i - w1 - 4] + src[i - w1 + 4] + src[i + w1 - 4] + src[i + w1 + 4] + src[i + w4 - 1] + src[i + w4 + 1])
          + blurmat[10] * (src[i - w4] + src[i - 4] + src[i + 4] + src[i + w4])
          + blurmat[9] * (src[i - w3 - 3] + src[i - w3 + 3] + src[i + w3 - 3] + src[i + w3 + 3])
          + blurmat[8] * (src[i - w3 - 2] + src[i - w3 + 2] + src[i - w2 - 3] + src[i - w2 + 3] + src[i + w2 - 3] + src[i + w2 + 3] + src[i + w3 - 2] + src[i + w3 + 2])
          + blurmat[7] * (src[i - w3 - 1] + src[i - w3 + 1] + src[i - w1 - 3] + src[i - w1 + 3] + src[i + w1 - 3] + src[i + w1 + 3] + src[i + w3 - 1] + src[i + w3 + 1])
          + blurmat[6] * (src[i - w3] + src[i - 3] + src[i + 3] + src[i + w3])
          + blurmat[5] * (src[i - w2 - 2] + src[i - w2 + 2] + src[i + w2 - 2] + src[i + w2 + 2])
          + blurmat[4] * (src[i - w2 - 1] + src[i - w2 + 1] + src[i - w1 - 2] + src[i - w1 + 2] + src[i + w1 - 2] + src[i + w1 + 2] + src[i + w2 - 1] + src[i + w2 + 1])
          + blurmat[3] * (src[i - w2] + src[i - 2] + src[i + 2] + src[i + w2])
          + blurmat[2] * (src[i - w1 - 1] + src[i - w1 + 1] + src[i + w1 - 1] + src[i + w1 + 1])
          + blurmat[1] * (src[i - w1] + src[i - 1] + src[i + 1] + src[i + w1])
          + blurmat[0] * src[i];
      blend[i] = fminf(1.0f, fmaxf(0.0f, v));
    }
  extend_border(blend, width, height, 4);
  free(tmp);
}

/* rgb: the sharp demosaicer's frame, blended in place with VNG4 of `raw` (the module's input mosaic); mask != 0 writes the
 * blend mask into all four channels instead */
int orc_dual_demosaic(float *rgb, const float *raw, int width, int height, int x0, int y0, uint32_t filters, const float wb[3], float dual_threshold, int mask)
{
  if(width < 16 || height < 16 || dual_threshold <= 0.0f) return 0;
  const size_t n = (size_t)width * height;
  float *blend = malloc(sizeof(float) * n), *vng = malloc(sizeof(float) * 4 * n);
  orc_vng_interpolate(vng, raw, width, height, x0, y0, filters, 0);
  orc_color_smoothing(vng, width, height, 2);
  orc_dual_mask(rgb, blend, width, height, wb, dual_threshold);
  for(size_t idx = 0; idx < n; idx++)
    for(int c = 0; c < 4; c++) rgb[4 * idx + c] = mask ? blend[idx] : blend[idx] * (rgb[4 * idx + c] - vng[4 * idx + c]) + vng[4 * idx + c];
  free(blend), free(vng);
  return 0;
}

/* X-Trans (filters == 9): three colours, 6x6 periods, no green averaging; lane 3 of `out` is not written for the bilinear
 * pixels and is uninitialised memory in the reference for the VNG pixels (its row buffer is malloc'ed): left as found here. */
int orc_vng_interpolate_xtrans(float *out, const float *in, int width, int height, int x0, int y0, const uint8_t xtrans[36], int only_linear)
{
  g_xtrans = (const uint8_t(*)[6])xtrans;
  if(only_linear)
  {
    orc_lin_interpolate(out, in, width, height, x0, y0, 9u);
    return 0;
  }
  float *lin = malloc(sizeof(float) * 4 * width * height);
  memcpy(lin, out, sizeof(float) * 4 * width * height); /* carries lane 3 */
  orc_lin_interpolate(lin, in, width, height, x0, y0, 9u);
  vng_from_linear(out, lin, width, height, x0, y0, 9u);
  free(lin);
  return 0;
}
