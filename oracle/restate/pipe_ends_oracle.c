/* CPU restatement of the pointwise modules either side of the demosaic .. colorout path: rawprepare, temperature,
 * highlights (clip mode and the bypass) in front; exposure, gamma and the export's float -> integer conversions behind.
 * TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src:
 *   iop/rawprepare.c    compute_proper_crop :206-210, BL :413-418, process() :466-633
 *   iop/temperature.c   process() :486-608  (FC / FCxtrans: develop/imageop_math.h:190-222)
 *   iop/highlights.c    _hl_count_thresholds :232-253, _hl_count_clipped :266-292, _hl_copy_input :296-302,
 *                       process() :679-789;  iop/highlights/clip.c process_clip :60-85;  iop/highlights/inpaint.c
 *                       process_inpaint_bayer :63-82 with iop/highlights/lch.c interpolate_color :206-303
 *   iop/exposure.c      process() :501-544
 *   iop/gamma.c         _copy_output :352-364 (process() :367-377 without mask/channel display)
 *   imageio/imageio_core.c  _clamp_float_to_uint8 :706-714, _swap_byteorder_float_to_uint8 :717-728,
 *                           _export_final_buffer_to_uint16 :731-738
 * Pinned bit-for-bit against those lines cut verbatim (oracle/_ref: ref_rawprepare.c, ref_temperature.c,
 * ref_highlights.c, ref_exposure.c, ref_pipe_end.c) by tests/test_cpu_pipe_ends.py.
 */
#include "oracle_common.h"
#include "b200iop.h"
#include <float.h>
#include <string.h>

/* ---- rawprepare ------------------------------------------------------------------------------------------------- */
static int proper_crop(double roi_scale, int value) { return (int)roundf((double)value * roi_scale); } /* the double product is
                                                                                    converted to float by roundf's prototype */

/* in: roi_in samples; out: roi_out floats (x channels).  gain: NULL or the four maps (host pointers inside d). */
int orc_rawprepare(const b200_piece_t *piece, const void *ivoid, void *ovoid)
{
  const b200_rawprepare_data_t *d = (const b200_rawprepare_data_t *)piece->data;
  const int width = piece->roi_out.width, height = piece->roi_out.height, input_width = piece->roi_in.width;
  const int roi_x = piece->roi_out.x, roi_y = piece->roi_out.y;
  const int cfa_x = roi_x + d->x, cfa_y = roi_y + d->y;
  const int csx = proper_crop(piece->roi_in.scale, d->x), csy = proper_crop(piece->roi_in.scale, d->y);
  float *const out = (float *)ovoid;
  const int mosaic = piece->filters && piece->channels == 1;

  if(mosaic && (piece->datatype == B200_TYPE_UINT16 || piece->datatype == B200_TYPE_FLOAT))
  {
    float inv_div[4];
    for(int k = 0; k < 4; k++) inv_div[k] = 1.0f / d->div[k];
    for(int j = 0; j < height; j++)
      for(int i = 0; i < width; i++)
      {
        const int id = (((j + cfa_y) & 1) << 1) + ((cfa_x + i) & 1);
        const size_t pin = (size_t)input_width * (j + csy) + csx + i;
        const float v = piece->datatype == B200_TYPE_UINT16 ? (float)((const uint16_t *)ivoid)[pin] : ((const float *)ivoid)[pin];
        out[(size_t)j * width + i] = (v - d->sub[id]) * inv_div[id];
      }
  }
  else
  {
    const float *const in = (const float *)ivoid;
    const float sub = d->sub[0], div = d->div[0];
    const int ch = (int)piece->channels;
    for(int j = 0; j < height; j++)
      for(int i = 0; i < width; i++)
        for(int c = 0; c < ch; c++)
        {
          /* the reference's index is `(size_t)ch * (int product)`: the inner product is int arithmetic */
          const size_t pin = (size_t)ch * (input_width * (j + csy) + csx + i) + c;
          const size_t pout = (size_t)ch * (j * width + i) + c;
          out[pout] = (in[pin] - sub) / div;
        }
  }

  if(mosaic && d->apply_gainmaps)
  {
    const b200_dng_gain_map_t *const g0 = d->gainmaps[0];
    const uint32_t map_w = g0->map_points_h, map_h = g0->map_points_v;
    const float im_to_rel_x = 1.0f / piece->buf_in_width, im_to_rel_y = 1.0f / piece->buf_in_height;
    const float rel_to_map_x = 1.0f / g0->map_spacing_h, rel_to_map_y = 1.0f / g0->map_spacing_v; /* double division, then float */
    const float map_origin_h = g0->map_origin_h, map_origin_v = g0->map_origin_v;
    for(int j = 0; j < height; j++)
    {
      /* CLAMP(float, int 0, uint32): the comparisons and the result are float */
      float y_map = ((roi_y + csy + j) * im_to_rel_y - map_origin_v) * rel_to_map_y;
      y_map = y_map < 0 ? (float)0 : (y_map > (float)map_h ? (float)map_h : y_map);
      const uint32_t y_i0 = (uint32_t)(y_map < (float)(map_h - 1) ? y_map : (float)(map_h - 1));
      const uint32_t y_i1 = (y_i0 + 1 < map_h - 1) ? y_i0 + 1 : map_h - 1;
      const float y_frac = y_map - y_i0;
      for(int i = 0; i < width; i++)
      {
        const int id = (((j + roi_y + d->y) & 1) << 1) + ((i + roi_x + d->x) & 1);
        float x_map = ((roi_x + csx + i) * im_to_rel_x - map_origin_h) * rel_to_map_x;
        x_map = x_map < 0 ? (float)0 : (x_map > (float)map_w ? (float)map_w : x_map);
        const uint32_t x_i0 = (uint32_t)(x_map < (float)(map_w - 1) ? x_map : (float)(map_w - 1));
        const uint32_t x_i1 = (x_i0 + 1 < map_w - 1) ? x_i0 + 1 : map_w - 1;
        const float x_frac = x_map - x_i0;
        const float *row0 = &d->gainmaps[id]->map_gain[y_i0 * map_w], *row1 = &d->gainmaps[id]->map_gain[y_i1 * map_w];
        const float gain_top = (1.0f - x_frac) * row0[x_i0] + x_frac * row0[x_i1];
        const float gain_bottom = (1.0f - x_frac) * row1[x_i0] + x_frac * row1[x_i1];
        out[j * width + i] *= (1.0f - y_frac) * gain_top + y_frac * gain_bottom;
      }
    }
  }
  return 0;
}

/* ---- temperature ------------------------------------------------------------------------------------------------ */
static int fc_xtrans(int row, int col, int roi_x, int roi_y, const uint8_t xtrans[6][6])
{
  return xtrans[(row + 600 + roi_y) % 6][(col + 600 + roi_x) % 6];
}
int orc_temperature(const b200_piece_t *piece, const float *in, float *out)
{
  const b200_temperature_data_t *d = (const b200_temperature_data_t *)piece->data;
  const int width = piece->roi_out.width, height = piece->roi_out.height;
  const uint32_t filters = piece->filters;
  if(filters == 9u)
  {
    for(int j = 0; j < height; j++)
      for(int i = 0; i < width; i++)
        out[(size_t)j * width + i] = in[(size_t)j * width + i] * d->coeffs[fc_xtrans(j, i % 12, piece->roi_out.x, piece->roi_out.y, piece->xtrans)];
  }
  else if(filters)
  {
    for(int j = 0; j < height; j++)
      for(int i = 0; i < width; i++)
        out[(size_t)j * width + i] = in[(size_t)j * width + i] * d->coeffs[orc_fc(j + piece->roi_out.y, i + piece->roi_out.x, filters)];
  }
  else
  {
    const size_t ch = piece->channels, npixels = (size_t)width * height;
    for(size_t k = 0; k < npixels; k++)
    {
      for(int c = 0; c < 3; c++) out[ch * k + c] = in[ch * k + c] * d->coeffs[c];
      if(ch == 4) out[4 * k + 3] = in[4 * k + 3]; /* other channel counts leave the rest of the pixel unwritten */
    }
    if((piece->mask_display & 1) && ch == 4)
      for(size_t k = 3; k < npixels * 4; k += 4) out[k] = in[k];
  }
  return 0;
}

/* ---- highlights --------------------------------------------------------------------------------------------------- */
/* dt_dev_get_roi_filters (develop/imageop.c:139-142): the CFA word seen from the ROI origin.  The shift itself is
 * rawspeed's ColorFilterArray::shiftDcrawFilter (third party, pinned submodule external/rawspeed; ColorFilterArray.cpp:
 * 143-170), restated: odd x swaps the colours of each horizontal pair, y rotates by one row (four bits) per step. */
uint32_t orc_roi_filters(uint32_t filters, int x, int y)
{
  if(!filters || filters == 9u) return filters;
  if((x < 0 ? -x : x) & 1)
    for(int n = 0; n < 8; n++)
    {
      const int i = n * 4, j = i + 2;
      const uint32_t t = ((filters >> i) ^ (filters >> j)) & 3u;
      filters ^= (t << i) | (t << j);
    }
  y *= 4;
  y = y >= 0 ? y % 32 : 32 - ((-y) % 32);
  if(y != 0 && y != 32) filters = (filters >> y) | (filters << (32 - y));
  return filters;
}

/* one line of the colour inpainting, lch.c:206-303: a running ratio between neighbouring sites of the line, decayed
 * exponentially over unclipped pairs, restores a clipped sample from its neighbour; four passes (row left-to-right,
 * right-to-left, column down, up) are averaged.  dim 0 = along a row (`other` = the row), 1 = along a column. */
static void interpolate_color(const float *ivoid, float *ovoid, int width, int height, int dim, int dir, int other, const float *clip, uint32_t filters,
                              int pass)
{
  float ratio = 1.0f;
  int i = 0, j = 0;
  if(dim == 0) j = other; else i = other;
  ptrdiff_t offs = dim ? width : 1;
  if(dir < 0) offs = -offs;
  const int n = dim ? height : width;
  const int beg = dir == 1 ? 0 : n - 1, end = dir == 1 ? n : -1;
  const float *in = ivoid + (dim ? i + (size_t)beg * width : beg + (size_t)j * width);
  float *out = ovoid + (in - ivoid);
  for(int k = beg; k != end; k += dir)
  {
    if(dim == 1) j = k; else i = k;
    const float clip0 = clip[orc_fc(j, i, filters)];
    const float clip1 = clip[orc_fc(dim ? (j + 1) : j, dim ? i : (i + 1), filters)];
    if(i == 0 || i == width - 1 || j == 0 || j == height - 1)
    {
      if(pass == 3) out[0] = in[0];
    }
    else
    {
      if(in[0] < clip0 && in[0] > 1e-5f)
        if(in[offs] < clip1 && in[offs] > 1e-5f)
        {
          if(k & 1)
            ratio = (3.0f * ratio + in[0] / in[offs]) / 4.0f;
          else
            ratio = (3.0f * ratio + in[offs] / in[0]) / 4.0f;
        }
      if(in[0] >= clip0 - 1e-5f)
      {
        float add = 0.0f;
        if(in[offs] >= clip1 - 1e-5f)
          add = fmaxf(clip0, clip1);
        else if(k & 1)
          add = in[offs] * ratio;
        else
          add = in[offs] / ratio;
        if(pass == 0)
          out[0] = add;
        else if(pass == 3)
          out[0] = (out[0] + add) / 4.0f;
        else
          out[0] += add;
      }
      else if(pass == 3)
        out[0] = in[0];
    }
    out += offs;
    in += offs;
  }
}

/* LCh reconstruction on a Bayer mosaic, lch.c:315-411: every 2x2 block with a clipped sample is rebuilt from its clipped and
 * unclipped lightness / chroma / hue.  SQRT3 and SQRT12 are long double literals in the reference (common.h:618-619): the
 * products, the quotients and the sums they enter are x87 operations, rounded to float on assignment -- restated with the
 * same types.  filters: the sensor word; the ROI origin enters through x0, y0. */
static void lch_bayer(const float *ivoid, float *ovoid, int width, int height, int x0, int y0, uint32_t filters, float clip)
{
  static const long double SQRT3 = 1.7320508075688772935274463415058723669L, SQRT12 = 3.4641016151377545870548926830117447339L;
  for(int j = 0; j < height; j++)
    for(int i = 0; i < width; i++)
    {
      float *const out = ovoid + (size_t)width * j + i;
      const float *const in = ivoid + (size_t)width * j + i;
      if(i == width - 1 || j == height - 1)
      {
        out[0] = clip < in[0] ? clip : in[0];
        continue;
      }
      int clipped = 0;
      float R = 0.0f, Gmin = FLT_MAX, Gmax = -FLT_MAX, B = 0.0f;
      for(int jj = 0; jj <= 1; jj++)
        for(int ii = 0; ii <= 1; ii++)
        {
          const float val = in[(size_t)jj * width + ii];
          clipped = (clipped || (val > clip));
          switch(orc_fc(j + jj + y0, i + ii + x0, filters))
          {
            case 0: R = val; break;
            case 1:
              Gmin = Gmin < val ? Gmin : val; /* MIN(Gmin, val) */
              Gmax = Gmax > val ? Gmax : val; /* MAX(Gmax, val) */
              break;
            case 2: B = val; break;
          }
        }
      if(!clipped)
      {
        out[0] = in[0];
        continue;
      }
      const float Ro = R < clip ? R : clip, Go = Gmin < clip ? Gmin : clip, Bo = B < clip ? B : clip; /* MIN(x, clip) */
      const float L = (R + Gmax + B) / 3.0f;
      float C = SQRT3 * (R - Gmax);
      float H = 2.0f * B - Gmax - R;
      const float Co = SQRT3 * (Ro - Go);
      const float Ho = 2.0f * Bo - Go - Ro;
      if(R != Gmax && Gmax != B)
      {
        const float ratio = sqrtf((Co * Co + Ho * Ho) / (C * C + H * H));
        C *= ratio;
        H *= ratio;
      }
      float RGB[3];
      RGB[0] = L - H / 6.0f + C / SQRT12;
      RGB[1] = L - H / 6.0f - C / SQRT12;
      RGB[2] = L + H / 3.0f;
      out[0] = RGB[orc_fc(j + y0, i + x0, filters)];
    }
}

/* ---- the X-Trans variants: lch.c interp_pix_xtrans :66-88, interpolate_color_xtrans :90-204, process_lch_xtrans :412-537,
 * inpaint.c process_inpaint_xtrans :84-104.  FCxtrans(row, col, roi, xtrans) = xtrans[(row + 600 + roi.y) % 6][(col + 600 + roi.x) % 6] */
typedef struct { const uint8_t (*xtrans)[6]; int x0, y0; } xt_t;
static int fcx(const xt_t *X, int row, int col) { return X->xtrans[(row + 600 + X->y0) % 6][(col + 600 + X->x0) % 6]; }
static float interp_pix_xtrans(int ratio_next, ptrdiff_t offset_next, float clip0, float clip_next, const float *in, const float *ratios)
{
  const float clip_val = fmaxf(clip0, clip_next);
  if(in[offset_next] >= clip_next - 1e-5f) return clip_val;
  if(ratio_next > 0) return fminf(in[offset_next] / ratios[ratio_next], clip_val);
  return fminf(in[offset_next] * ratios[-ratio_next], clip_val);
}
static void interpolate_color_xtrans(const float *ivoid, float *ovoid, int width, int height, int dim, int dir, int other, const float *clip, const xt_t *X,
                                     int pass)
{
  static const int roff[3][3] = { { 0, -1, -2 }, { 1, 0, -3 }, { 2, 3, 0 } };
  float ratios[4] = { 1.0f, 1.0f, 1.0f, 1.0f };
  int i = (dim == 0) ? 0 : other, j = (dim == 0) ? other : 0;
  const ptrdiff_t offs = (ptrdiff_t)(dim ? width : 1) * ((dir < 0) ? -1 : 1);
  const ptrdiff_t offl = offs - (dim ? 1 : width), offr = offs + (dim ? 1 : width);
  const int n = dim ? height : width, beg = dir == 1 ? 0 : n - 1, end = dir == 1 ? n : -1;
  const float *in = ivoid + (dim ? (size_t)i + (size_t)beg * width : (size_t)beg + (size_t)j * width);
  float *out = ovoid + (in - ivoid);
  for(int k = beg; k != end; k += dir)
  {
    if(dim == 1) j = k; else i = k;
    const int f0 = fcx(X, j, i), f1 = fcx(X, dim ? (j + dir) : j, dim ? i : (i + dir));
    const int fl = fcx(X, dim ? (j + dir) : (j - 1), dim ? (i - 1) : (i + dir)), fr = fcx(X, dim ? (j + dir) : (j + 1), dim ? (i + 1) : (i + dir));
    const float clip0 = clip[f0], clip1 = clip[f1], clipl = clip[fl], clipr = clip[fr];
    const float clip_max = fmaxf(fmaxf(clip[0], clip[1]), clip[2]);
    if(i == 0 || i == width - 1 || j == 0 || j == height - 1)
    {
      if(pass == 3) out[0] = fminf(clip_max, in[0]);
    }
    else
    {
      if((f0 != f1) && (in[0] < clip0 && in[0] > 1e-5f) && (in[offs] < clip1 && in[offs] > 1e-5f))
      {
        const int r = roff[f0][f1];
        if(r > 0)
          ratios[r] = (3.f * ratios[r] + (in[offs] / in[0])) / 4.f;
        else
          ratios[-r] = (3.f * ratios[-r] + (in[0] / in[offs])) / 4.f;
      }
      if(in[0] >= clip0 - 1e-5f)
      {
        float add;
        if(f0 != f1)
          add = interp_pix_xtrans(roff[f0][f1], offs, clip0, clip1, in, ratios);
        else
          add = (fl != f0) ? interp_pix_xtrans(roff[f0][fl], offl, clip0, clipl, in, ratios) : interp_pix_xtrans(roff[f0][fr], offr, clip0, clipr, in, ratios);
        if(pass == 0)
          out[0] = add;
        else if(pass == 3)
          out[0] = fminf(clip_max, (out[0] + add) / 4.0f);
        else
          out[0] += add;
      }
      else if(pass == 3)
        out[0] = in[0];
    }
    out += offs;
    in += offs;
  }
}
static void lch_xtrans(const float *ivoid, float *ovoid, int width, int height, const xt_t *X, float clip)
{
  static const long double SQRT3 = 1.7320508075688772935274463415058723669L, SQRT12 = 3.4641016151377545870548926830117447339L;
  for(int j = 0; j < height; j++)
  {
    int cl = 0; /* clipping of the vertical triplets of this and the two previous columns */
    for(int i = 0; i < width; i++)
    {
      const float *in = ivoid + (size_t)width * j + i;
      float *out = ovoid + (size_t)width * j + i;
      cl = (cl << 1) & 6;
      if(j >= 2 && j <= height - 3) cl |= (in[-width] > clip) | (in[0] > clip) | (in[width] > clip);
      if(i < 2 || i > width - 3 || j < 2 || j > height - 3)
      {
        out[0] = clip < in[0] ? clip : in[0];
        continue;
      }
      int clipped = (in[0] > clip);
      if(!clipped)
      {
        clipped = cl;
        if(clipped)
          for(int offset_j = -2; offset_j <= 0; offset_j++)
            for(int offset_i = -2; offset_i <= 0; offset_i++)
              if(clipped)
              {
                clipped = 0;
                for(int jj = offset_j; jj <= offset_j + 2; jj++)
                  for(int ii = offset_i; ii <= offset_i + 2; ii++) clipped = (clipped || (in[(ptrdiff_t)jj * width + ii] > clip));
              }
      }
      if(!clipped)
      {
        out[0] = in[0];
        continue;
      }
      float mean[3] = { 0.0f, 0.0f, 0.0f }, RGBmax[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
      int cnt[3] = { 0, 0, 0 };
      for(int jj = -1; jj <= 1; jj++)
        for(int ii = -1; ii <= 1; ii++)
        {
          const float val = in[(ptrdiff_t)jj * width + ii];
          const int c = fcx(X, j + jj, i + ii);
          mean[c] += val;
          cnt[c]++;
          RGBmax[c] = RGBmax[c] > val ? RGBmax[c] : val;
        }
      const float m0 = mean[0] / cnt[0], m1 = mean[1] / cnt[1], m2 = mean[2] / cnt[2];
      const float Ro = m0 < clip ? m0 : clip, Go = m1 < clip ? m1 : clip, Bo = m2 < clip ? m2 : clip;
      const float R = RGBmax[0], G = RGBmax[1], B = RGBmax[2];
      const float L = (R + G + B) / 3.0f;
      float C = SQRT3 * (R - G);
      float H = 2.0f * B - G - R;
      const float Co = SQRT3 * (Ro - Go);
      const float Ho = 2.0f * Bo - Go - Ro;
      if(R != G && G != B)
      {
        const float ratio = sqrtf((Co * Co + Ho * Ho) / (C * C + H * H));
        C *= ratio;
        H *= ratio;
      }
      float RGB[3];
      RGB[0] = L - H / 6.0f + C / SQRT12;
      RGB[1] = L - H / 6.0f - C / SQRT12;
      RGB[2] = L + H / 3.0f;
      out[0] = RGB[fcx(X, j, i)];
    }
  }
}

/* returns 0 and the number of samples counted as clipped in *n_clipped; -1 for a mode that is not restated */
int orc_highlights(const b200_piece_t *piece, const float *in, float *out, size_t *n_clipped)
{
  const b200_highlights_data_t *data = (const b200_highlights_data_t *)piece->data;
  const uint32_t filters = piece->filters;
  const size_t n_pixels = (size_t)piece->roi_out.width * piece->roi_out.height;
  float pmax[4];
  for(int c = 0; c < 4; c++) pmax[c] = (piece->processed_maximum[c] > 0.f) ? piece->processed_maximum[c] : 1.0f;
  const float clip = data->clip * fminf(pmax[0], fminf(pmax[1], pmax[2]));
  float thresholds[4];
  float factor = 0.f;
  if(data->mode == B200_HIGHLIGHTS_INPAINT) factor = 0.987f;
  if(data->mode == B200_HIGHLIGHTS_LAPLACIAN || data->mode == B200_HIGHLIGHTS_HARMONIC) factor = 0.995f;
  for(int c = 0; c < 3; c++) thresholds[c] = (factor > 0.f) ? factor * data->clip * pmax[c] : clip;
  thresholds[3] = clip;

  size_t clipped = 0;
  const size_t ch = filters ? 1 : piece->channels;
  if(filters)
  {
    const float raw_threshold = fminf(fminf(thresholds[0], thresholds[1]), thresholds[2]);
    for(size_t k = 0; k < n_pixels; k++) clipped += (in[k] > raw_threshold);
  }
  else
  {
    const size_t n_colours = ch < 3 ? ch : 3;
    for(size_t k = 0; k < n_pixels; k++)
    {
      int over = 0;
      for(size_t c = 0; c < n_colours; c++) over |= (in[k * ch + c] > thresholds[c]);
      clipped += (over != 0);
    }
  }
  if(n_clipped) *n_clipped = clipped;
  if(clipped < 25)
  {
    memcpy(out, in, sizeof(float) * n_pixels * ch);
    return 0;
  }
  /* past the bypass: the reconstruction modes are not restated (on non-mosaic input LCh and inpainting are process_clip) */
  if(data->mode == B200_HIGHLIGHTS_LAPLACIAN || data->mode == B200_HIGHLIGHTS_HARMONIC) return -1;
  if(filters && filters != 9u && data->mode == B200_HIGHLIGHTS_INPAINT)
  { /* process() :735-746 */
    const float clips[4] = { 0.987f * data->clip * pmax[0], 0.987f * data->clip * pmax[1], 0.987f * data->clip * pmax[2], clip };
    const int w = piece->roi_out.width, h = piece->roi_out.height;
    const uint32_t shifted = orc_roi_filters(filters, piece->roi_in.x, piece->roi_in.y); /* :691 */
    for(int j = 0; j < h; j++)
    {
      interpolate_color(in, out, w, h, 0, 1, j, clips, shifted, 0);
      interpolate_color(in, out, w, h, 0, -1, j, clips, shifted, 1);
    }
    for(int i = 0; i < w; i++)
    {
      interpolate_color(in, out, w, h, 1, 1, i, clips, shifted, 2);
      interpolate_color(in, out, w, h, 1, -1, i, clips, shifted, 3);
    }
    return 0;
  }
  if(filters == 9u && (data->mode == B200_HIGHLIGHTS_LCH || data->mode == B200_HIGHLIGHTS_INPAINT))
  {
    const xt_t X = { piece->xtrans, piece->roi_in.x, piece->roi_in.y };
    const int w = piece->roi_out.width, h = piece->roi_out.height;
    if(data->mode == B200_HIGHLIGHTS_LCH)
      lch_xtrans(in, out, w, h, &X, clip);
    else
    {
      const float clips[4] = { 0.987f * data->clip * pmax[0], 0.987f * data->clip * pmax[1], 0.987f * data->clip * pmax[2], clip };
      for(int j = 0; j < h; j++)
      {
        interpolate_color_xtrans(in, out, w, h, 0, 1, j, clips, &X, 0);
        interpolate_color_xtrans(in, out, w, h, 0, -1, j, clips, &X, 1);
      }
      for(int i = 0; i < w; i++)
      {
        interpolate_color_xtrans(in, out, w, h, 1, 1, i, clips, &X, 2);
        interpolate_color_xtrans(in, out, w, h, 1, -1, i, clips, &X, 3);
      }
    }
    return 0;
  }
  if(filters && filters != 9u && data->mode == B200_HIGHLIGHTS_LCH)
  { /* process() :748-757: process_lch_bayer reads piece->dsc_in.filters with the ROI origin added to its coordinates */
    lch_bayer(in, out, piece->roi_out.width, piece->roi_out.height, piece->roi_out.x, piece->roi_out.y, filters, clip);
    return 0;
  }
  if(filters && data->mode != B200_HIGHLIGHTS_CLIP) return -1;
  for(size_t k = 0; k < ch * n_pixels; k++) out[k] = clip < in[k] ? clip : in[k];
  /* dt_iop_alpha_copy assumes four floats per pixel whatever the buffer holds: on a mosaic it would run past both buffers */
  if((piece->mask_display & 1) && !filters && ch == 4)
    for(size_t k = 3; k < n_pixels * 4; k += 4) out[k] = in[k];
  return 0;
}

/* ---- exposure ----------------------------------------------------------------------------------------------------- */
int orc_exposure(const b200_piece_t *piece, const float *in, float *out)
{
  const b200_exposure_data_t *d = (const b200_exposure_data_t *)piece->data;
  const size_t n = (size_t)piece->channels * piece->roi_out.width * piece->roi_out.height;
  for(size_t k = 0; k < n; k++) out[k] = (in[k] - d->black) * d->scale;
  if(piece->mask_display & 1)
    for(size_t k = 3; k < (size_t)piece->roi_out.width * piece->roi_out.height * 4; k += 4) out[k] = in[k];
  return 0;
}

/* ---- the float -> integer ends -------------------------------------------------------------------------------------- */
/* C's (int)float on the reference's hardware: cvttss2si, INT_MIN for NaN and out-of-range */
static int cvtt(float f) { return (f >= -2147483648.0f && f < 2147483648.0f) ? (int)f : (int)0x80000000u; }

void orc_gamma_copy_output(const float *in, uint8_t *out, size_t npixels)
{
  for(size_t j = 0; j < 4 * npixels; j += 4)
    for(size_t c = 0; c < 3; c++) out[j + 2 - c] = (uint8_t)cvtt(fminf(roundf(255.0f * fmaxf(in[j + c], 0.0f)), 255.0f));
}
static float clampf(float a, float mn, float mx) { return a >= mn ? (a <= mx ? a : mx) : mn; } /* CLAMPF, math/math.h:91 */
void orc_export_convert(const float *in, void *out, size_t npixels, int format)
{
  if(format == B200_EXPORT_UINT8)
    for(size_t k = 0; k < 4 * npixels; k++) ((uint8_t *)out)[k] = (uint8_t)cvtt(clampf(roundf(in[k] * 255.f), 0.f, 255.f));
  else if(format == B200_EXPORT_UINT8_SWAP)
    for(size_t k = 0; k < npixels; k++)
    {
      uint8_t *o = (uint8_t *)out + 4 * k;
      o[0] = (uint8_t)cvtt(clampf(roundf(in[4 * k + 2] * 255.f), 0.f, 255.f));
      o[1] = (uint8_t)cvtt(clampf(roundf(in[4 * k + 1] * 255.f), 0.f, 255.f));
      o[2] = (uint8_t)cvtt(clampf(roundf(in[4 * k + 0] * 255.f), 0.f, 255.f));
      o[3] = (uint8_t)cvtt(clampf(roundf(in[4 * k + 3] * 255.f), 0.f, 255.f));
    }
  else
    for(size_t k = 0; k < 4 * npixels; k++)
    {
      /* glib's CLAMP: x > high ? high : (x < low ? low : x) -- NaN passes through to the conversion */
      const float x = roundf(in[k] * 65535.f);
      const float c = x > 65535.f ? 65535.f : (x < 0.f ? 0.f : x);
      ((uint16_t *)out)[k] = (uint16_t)cvtt(c);
    }
}
