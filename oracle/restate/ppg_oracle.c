/* CPU restatement of the PPG demosaicer (the reference's fallback Bayer method).  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src/iop/demosaic/ppg.c demosaic_ppg :21-211 (border average :28-56, optional pre-median :57-67,
 * green interpolation :70-131, red/blue interpolation in place :136-203) and iop/demosaic/basic.c pre_median_b :136-180.
 * Pinned bit-for-bit against those lines cut verbatim (oracle/_ref: ref_ppg.c).  The caller passes the ROI-shifted
 * filters word and an output buffer whose outer three pixels keep their alpha, as in the reference.
 */
#include "oracle_common.h"
#include <stdlib.h>
#include <string.h>

static void pre_median(float *out, const float *in, int width, int height, uint32_t filters, float threshold)
{
  memcpy(out, in, sizeof(float) * (size_t)width * height);
  static const int lim[5] = { 0, 1, 2, 1, 0 };
  for(int row = 3; row < height - 3; row++)
  {
    int col = 3;
    if(orc_fc(row, col, filters) != 1 && orc_fc(row, col, filters) != 3) col++;
    for(; col < width - 3; col += 2)
    {
      const float *pixi = in + (size_t)width * row + col;
      float med[9];
      int cnt = 0, k = 0;
      for(int i = 0; i < 5; i++)
        for(int j = -lim[i]; j <= lim[i]; j += 2)
        {
          const float v = pixi[width * (i - 2) + j];
          if(fabsf(v - pixi[0]) < threshold)
          {
            med[k++] = v;
            cnt++;
          }
          else
            med[k++] = 64.0f + v;
        }
      for(int i = 0; i < 8; i++)
        for(int ii = i + 1; ii < 9; ii++)
          if(med[i] > med[ii])
          {
            const float t = med[ii];
            med[ii] = med[i];
            med[i] = t;
          }
      out[(size_t)width * row + col] = (cnt == 1 ? med[4] - 64.0f : med[(cnt - 1) / 2]);
    }
  }
}

int orc_demosaic_ppg(float *out, const float *in, int width, int height, uint32_t filters, float thrs)
{
  /* the three-pixel border: average of the neighbours of each colour inside the frame */
  for(int j = 0; j < height; j++)
    for(int i = 0; i < width; i++)
    {
      if(i == 3 && j >= 3 && j < height - 3) i = width - 3;
      if(i == width) break;
      float sum[8] = { 0 };
      for(int y = j - 1; y != j + 2; y++)
        for(int x = i - 1; x != i + 2; x++)
          if(y >= 0 && x >= 0 && y < height && x < width)
          {
            const int f = orc_fc(y, x, filters);
            sum[f] += in[(size_t)y * width + x];
            sum[f + 4]++;
          }
      const int f = orc_fc(j, i, filters);
      for(int c = 0; c < 3; c++)
        out[4 * ((size_t)j * width + i) + c] = (c != f && sum[c + 4] > 0.0f) ? sum[c] / sum[c + 4] : in[(size_t)j * width + i];
    }
  const float *input = in;
  float *med = NULL;
  if(thrs > 0.0f)
  {
    med = malloc(sizeof(float) * (size_t)width * height);
    pre_median(med, in, width, height, filters, thrs);
    input = med;
  }
  /* green */
  for(int j = 3; j < height - 3; j++)
    for(int i = 3; i < width - 3; i++)
    {
      const float *b = input + (size_t)width * j + i;
      float *o = out + 4 * ((size_t)width * j + i);
      const int c = orc_fc(j, i, filters);
      const float pc = b[0];
      if(c == 0 || c == 2)
      {
        o[c] = pc;
        const float pym = b[-width], pym2 = b[-2 * width], pym3 = b[-3 * width], pyM = b[width], pyM2 = b[2 * width], pyM3 = b[3 * width];
        const float pxm = b[-1], pxm2 = b[-2], pxm3 = b[-3], pxM = b[1], pxM2 = b[2], pxM3 = b[3];
        const float guessx = (pxm + pc + pxM) * 2.0f - pxM2 - pxm2;
        const float diffx = (fabsf(pxm2 - pc) + fabsf(pxM2 - pc) + fabsf(pxm - pxM)) * 3.0f + (fabsf(pxM3 - pxM) + fabsf(pxm3 - pxm)) * 2.0f;
        const float guessy = (pym + pc + pyM) * 2.0f - pyM2 - pym2;
        const float diffy = (fabsf(pym2 - pc) + fabsf(pyM2 - pc) + fabsf(pym - pyM)) * 3.0f + (fabsf(pyM3 - pyM) + fabsf(pym3 - pym)) * 2.0f;
        if(diffx > diffy)
          o[1] = fmaxf(fminf(guessy * .25f, fmaxf(pym, pyM)), fminf(pym, pyM));
        else
          o[1] = fmaxf(fminf(guessx * .25f, fmaxf(pxm, pxM)), fminf(pxm, pxM));
      }
      else
        o[1] = pc;
      o[3] = 0.0f; /* the other colours of the pixel are written by the next pass */
    }
  /* red and blue, in place: every value read here is one this pass does not write */
  for(int j = 1; j < height - 1; j++)
    for(int i = 1; i < width - 1; i++)
    {
      float *buf = out + 4 * ((size_t)width * j + i);
      const int c = orc_fc(j, i, filters);
      const int w4 = 4 * width;
      float color[4] = { buf[0], buf[1], buf[2], buf[3] };
      if(c & 1)
      {
        const float *nt = buf - w4, *nb = buf + w4, *nl = buf - 4, *nr = buf + 4;
        if(orc_fc(j, i + 1, filters) == 0)
        {
          color[2] = (nt[2] + nb[2] + 2.0f * color[1] - nt[1] - nb[1]) * .5f;
          color[0] = (nl[0] + nr[0] + 2.0f * color[1] - nl[1] - nr[1]) * .5f;
        }
        else
        {
          color[0] = (nt[0] + nb[0] + 2.0f * color[1] - nt[1] - nb[1]) * .5f;
          color[2] = (nl[2] + nr[2] + 2.0f * color[1] - nl[1] - nr[1]) * .5f;
        }
      }
      else
      {
        const float *ntl = buf - 4 - w4, *ntr = buf + 4 - w4, *nbl = buf - 4 + w4, *nbr = buf + 4 + w4;
        const int o = c == 0 ? 2 : 0; /* a red site gets blue, a blue site red */
        const float diff1 = fabsf(ntl[o] - nbr[o]) + fabsf(ntl[1] - color[1]) + fabsf(nbr[1] - color[1]);
        const float guess1 = ntl[o] + nbr[o] + 2.0f * color[1] - ntl[1] - nbr[1];
        const float diff2 = fabsf(ntr[o] - nbl[o]) + fabsf(ntr[1] - color[1]) + fabsf(nbl[1] - color[1]);
        const float guess2 = ntr[o] + nbl[o] + 2.0f * color[1] - ntr[1] - nbl[1];
        if(diff1 > diff2)
          color[o] = guess2 * .5f;
        else if(diff1 < diff2)
          color[o] = guess1 * .5f;
        else
          color[o] = (guess1 + guess2) * .25f;
      }
      memcpy(buf, color, sizeof(color));
    }
  free(med);
  return 0;
}

/* the passthrough methods, iop/demosaic/passthrough.c:22-87 (dispatch demosaic.c:1111-1118): roi_out's origin is zeroed by
 * process(), so the Bayer colour ignores the ROI origin while the X-Trans colour (FCxtrans with roi_in) follows it.
 * Lane 3 of the output is not written. */
int orc_demosaic_passthrough(float *out, const float *in, int width, int height, int x, int y, uint32_t filters, const uint8_t xtrans[36], int colour)
{
  for(int row = 0; row < height; row++)
    for(int col = 0; col < width; col++)
    {
      const float val = in[(size_t)row * width + col];
      float *o = out + 4 * ((size_t)row * width + col);
      if(!colour)
        o[0] = o[1] = o[2] = val;
      else
      {
        const int ch = filters != 9u ? orc_fc(row, col, filters) : xtrans[((row + 600 + y) % 6) * 6 + (col + 600 + x) % 6];
        o[0] = o[1] = o[2] = 0.0f;
        o[ch] = val;
      }
    }
  return 0;
}

/* the half-size "downsample" method for a Bayer sensor, demosaic.c:480-532: every output pixel is the 2x2 block behind it, each
 * colour the mean of its samples in the block (clamped to the frame at odd edges); alpha 0 */
int orc_demosaic_downsample(float *out, const float *in, int width, int height, uint32_t filters)
{
  const int ow = (width + 1) / 2, oh = (height + 1) / 2;
  for(int y = 0; y < oh; y++)
    for(int x = 0; x < ow; x++)
    {
      float cam[4] = { 0.0f };
      int samples[4] = { 0 };
      const int px = 2 * x < width - 1 ? 2 * x : width - 1, py = 2 * y < height - 1 ? 2 * y : height - 1;
      for(int j = 0; j < 2; j++)
        for(int i = 0; i < 2; i++)
        {
          const int xx = px + i < width - 1 ? px + i : width - 1, yy = py + j < height - 1 ? py + j : height - 1;
          const int c = orc_fc(yy, xx, filters);
          cam[c] += in[(size_t)yy * width + xx];
          samples[c]++;
        }
      for(int c = 0; c < 4; c++)
        if(samples[c] > 0) cam[c] /= (float)samples[c];
      float *o = out + 4 * ((size_t)y * ow + x);
      o[0] = cam[0], o[1] = cam[1], o[2] = cam[2], o[3] = 0.0f;
    }
  return 0;
}

/* demosaic.c:543-615: a colour the 2x2 block does not sample.  The nearest same-colour photosite of each quadrant around the
 * block centre, searched over x in [px-3, px+4], y in [py-3, py+4] clamped into the frame (first found wins ties, raster order);
 * four quadrants -> bilinear inside the rectangle of their mean coordinates; fewer -> the mean of those found; none -> the
 * nearest sample overall (0 when there is none) */
static inline int xt_colour(int row, int col, int x0, int y0, const uint8_t *xtrans) { return xtrans[((row + y0 + 600) % 6) * 6 + (col + x0 + 600) % 6]; }
static float xt_missing(const float *in, int width, int height, int x0, int y0, int px, int py, const uint8_t *xtrans, int colour)
{
  const float cx = px + 0.5f, cy = py + 0.5f;
  const int xmin = px - 3 > 0 ? px - 3 : 0, xmax = px + 4 < width - 1 ? px + 4 : width - 1;
  const int ymin = py - 3 > 0 ? py - 3 : 0, ymax = py + 4 < height - 1 ? py + 4 : height - 1;
  float qv[4] = { 0.0f }, qd[4] = { INFINITY, INFINITY, INFINITY, INFINITY }, nearest = 0.0f, nearest_d = INFINITY;
  int qx[4] = { 0 }, qy[4] = { 0 }, found[4] = { 0 };
  for(int yy = ymin; yy <= ymax; yy++)
    for(int xx = xmin; xx <= xmax; xx++)
    {
      if(xt_colour(yy, xx, x0, y0, xtrans) != colour) continue;
      const float dx = xx - cx, dy = yy - cy;
      const float d2 = dx * dx + dy * dy;
      const float v = in[(size_t)yy * width + xx];
      if(d2 < nearest_d) nearest_d = d2, nearest = v;
      const int q = (yy > cy ? 2 : 0) + (xx > cx ? 1 : 0);
      if(d2 < qd[q]) qd[q] = d2, qv[q] = v, qx[q] = xx, qy[q] = yy, found[q] = 1;
    }
  if(found[0] && found[1] && found[2] && found[3])
  {
    const float xl = 0.5f * (qx[0] + qx[2]), xr = 0.5f * (qx[1] + qx[3]), yt = 0.5f * (qy[0] + qy[1]), yb = 0.5f * (qy[2] + qy[3]);
    const float wx = xr - xl > 1e-6f ? xr - xl : 1e-6f, wy = yb - yt > 1e-6f ? yb - yt : 1e-6f;
    float tx = (cx - xl) / wx, ty = (cy - yt) / wy;
    tx = tx < 0.0f ? 0.0f : (tx > 1.0f ? 1.0f : tx); /* CLAMP = MIN(MAX(x, lo), hi) on finite values */
    ty = ty < 0.0f ? 0.0f : (ty > 1.0f ? 1.0f : ty);
    const float top = qv[0] + tx * (qv[1] - qv[0]);
    const float bottom = qv[2] + tx * (qv[3] - qv[2]);
    return top + ty * (bottom - top);
  }
  float sum = 0.0f;
  int count = 0;
  for(int q = 0; q < 4; q++)
    if(found[q]) sum += qv[q], count++;
  return count > 0 ? sum / (float)count : nearest;
}

/* the half-size method on an X-Trans sensor, demosaic.c:624-666: the colours sampled in the 2x2 block are their means, the others
 * come from xt_missing(); alpha 0.  (x0, y0) = roi_in's origin on the sensor. */
int orc_demosaic_downsample_xtrans(float *out, const float *in, int width, int height, int x0, int y0, const uint8_t xtrans[36])
{
  const int ow = (width + 1) / 2, oh = (height + 1) / 2;
  for(int y = 0; y < oh; y++)
    for(int x = 0; x < ow; x++)
    {
      float rgb[3] = { 0.0f };
      int samples[3] = { 0 };
      const int px = 2 * x < width - 1 ? 2 * x : width - 1, py = 2 * y < height - 1 ? 2 * y : height - 1;
      for(int j = 0; j < 2; j++)
        for(int i = 0; i < 2; i++)
        {
          const int xx = px + i < width - 1 ? px + i : width - 1, yy = py + j < height - 1 ? py + j : height - 1;
          const int c = xt_colour(yy, xx, x0, y0, xtrans);
          rgb[c] += in[(size_t)yy * width + xx];
          samples[c]++;
        }
      float *o = out + 4 * ((size_t)y * ow + x);
      for(int c = 0; c < 3; c++) o[c] = samples[c] > 0 ? rgb[c] / (float)samples[c] : xt_missing(in, width, height, x0, y0, px, py, xtrans, c);
      o[3] = 0.0f;
    }
  return 0;
}

/* demosaic.c:514-521, four-colour Bayer sensors: each of R, G, B starts at 0.0f and takes CAM_to_RGB[c][k] * cam[k] for k = 0..3,
 * every product and sum in double, the running value rounded to float after each term */
int orc_demosaic_downsample4(float *out, const float *in, int width, int height, uint32_t filters, const double cam_to_rgb[12])
{
  const int ow = (width + 1) / 2, oh = (height + 1) / 2;
  for(int y = 0; y < oh; y++)
    for(int x = 0; x < ow; x++)
    {
      float cam[4] = { 0.0f };
      int samples[4] = { 0 };
      const int px = 2 * x < width - 1 ? 2 * x : width - 1, py = 2 * y < height - 1 ? 2 * y : height - 1;
      for(int j = 0; j < 2; j++)
        for(int i = 0; i < 2; i++)
        {
          const int xx = px + i < width - 1 ? px + i : width - 1, yy = py + j < height - 1 ? py + j : height - 1;
          const int c = orc_fc(yy, xx, filters);
          cam[c] += in[(size_t)yy * width + xx];
          samples[c]++;
        }
      for(int c = 0; c < 4; c++)
        if(samples[c] > 0) cam[c] /= (float)samples[c];
      float *o = out + 4 * ((size_t)y * ow + x);
      for(int c = 0; c < 3; c++)
      {
        float acc = 0.0f;
        for(int k = 0; k < 4; k++) acc = (float)((double)acc + cam_to_rgb[4 * c + k] * (double)cam[k]);
        o[c] = acc;
      }
      o[3] = 0.0f;
    }
  return 0;
}
