/* CPU restatement of the LMMSE demosaicer (L. Zhang, X. Wu; RawTherapee's implementation tiled for darktable).  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src/iop/demosaic/lmmse.c: limf :67-70, median3f :72-75, median9f :77-121, calc_gamma :123-135,
 * lmmse_demosaic :136-576; the two gamma tables: iop/demosaic.c:1208-1213.  Pinned bit-for-bit against those lines cut verbatim
 * (oracle/_ref: ref_lmmse.c, compiled without OpenMP).
 *
 * The reference works in tiles of 136x136 (128 of input + a margin of 4, of which 8 more on every side overlap the neighbours: 112 kept)
 * on six planes it zeroes ONCE per thread and carries from tile to tile (:166-175).  A tile does not rewrite everything it reads: the two
 * outermost rows / columns of the difference planes and whatever lies behind a short last tile keep what the previous tile of the same
 * thread left there, so the reference's result depends, near tile borders, on the order in which a thread met its tiles.  Two modes:
 *   carry = 1   the planes carried through the serial raster walk of the tiles: equal to the reference compiled without OpenMP;
 *   carry = 0   the planes zeroed in front of every tile: what the CUDA kernel computes (a tile is then a function of its input alone).
 * tests/test_cpu_lmmse.py measures the distance between the two.
 */
#include "oracle_common.h"
#include <stdlib.h>
#include <string.h>

#define GRP 136
#define BORDER 4
#define OVERLAP 8
#define TILESIZE (GRP - 2 * BORDER)
#define TILEVALID (TILESIZE - 2 * OVERLAP)
#define NP (GRP * GRP)

static float limf(float x, float mn, float mx) { return fmaxf(mn, fminf(x, mx)); }
static float median3f(float x0, float x1, float x2) { return fmaxf(fminf(x0, x1), fminf(x2, fmaxf(x0, x1))); }
/* :77-121, the network as written (it is not a full sorting network: the value it returns is what the reference returns) */
static float median9f(float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7, float a8)
{
  float t;
  t = fminf(a1, a2); a2 = fmaxf(a1, a2); a1 = t;
  t = fminf(a4, a5); a5 = fmaxf(a4, a5); a4 = t;
  t = fminf(a7, a8); a8 = fmaxf(a7, a8); a7 = t;
  t = fminf(a0, a1); a1 = fmaxf(a0, a1); a0 = t;
  t = fminf(a3, a4); a4 = fmaxf(a3, a4); a3 = t;
  t = fminf(a6, a7); a7 = fmaxf(a6, a7); a6 = t;
  t = fminf(a1, a2); a2 = fmaxf(a1, a2); a1 = t;
  t = fminf(a4, a5); a5 = fminf(a4, a5); a4 = t; /* sic, :101: both the minimum */
  t = fminf(a7, a8); a8 = fmaxf(a7, a8);
  a3 = fmaxf(a0, a3);
  a5 = fminf(a5, a8);
  a7 = fmaxf(a4, t);
  t = fminf(a4, t);
  a6 = fmaxf(a3, a6);
  a4 = fmaxf(a1, t);
  a2 = fminf(a2, a5);
  a4 = fminf(a4, a7);
  t = fminf(a4, a2);
  a2 = fmaxf(a4, a2);
  a4 = fmaxf(a6, t);
  return fminf(a4, a2);
}
static float calc_gamma(float val, const float *table)
{ /* :123-135 */
  const float index = val * 65535.0f;
  if(index < 0.0f) return 0.0f;
  if(index > 65534.99f) return 1.0f;
  const int idx = (int)index;
  const float diff = index - (float)idx;
  const float p1 = table[idx];
  const float p2 = table[idx + 1] - p1;
  return p1 + p2 * diff;
}
static float sqf(float x) { return x * x; }

/* iop/demosaic.c:1208-1213 */
void orc_lmmse_gamma_tables(float *gamma_in, float *gamma_out)
{
  for(int j = 0; j < 65536; j++)
  {
    const double x = (double)j / 65535.0;
    gamma_in[j] = (x <= 0.001867) ? x * 17.0 : 1.044445 * exp(log(x) / 2.4) - 0.044445;
    gamma_out[j] = (x <= 0.031746) ? x / 17.0 : exp(log((x + 0.044445) / 1.044445) * 2.4);
  }
}

typedef struct
{
  const float *in;
  float *out;
  int width, height, nv, nh, medians, refine;
  uint32_t filters;
  float scaler, revscaler, h0, h1, h2, h3, h4;
  const float *gin, *gout;
} lm_t;

/* variance-weighted estimate along one direction, :246-270: lp = the low-passed differences, df = the differences, s = the step */
static void lm_estimate(const float *lp, const float *df, int s, float *x, float *v)
{
  float p[9];
  for(int k = 0; k < 9; k++) p[k] = lp[(k - 4) * s];
  const float mu = (p[0] + p[1] + p[2] + p[3] + p[4] + p[5] + p[6] + p[7] + p[8]) / 9.0f;
  float vx = 1e-7f;
  for(int k = 0; k < 9; k++) vx += sqf(p[k] - mu);
  for(int k = 0; k < 9; k++) p[k] -= df[(k - 4) * s];
  float vn = 1e-7f;
  for(int k = 0; k < 9; k++) vn += sqf(p[k]);
  *x = (df[0] * vx + lp[0] * vn) / (vx + vn);
  *v = vx * vn / (vx + vn);
}

static void lm_tile(const lm_t *m, float *const qix[6], int tv, int th)
{
  const int width = m->width, height = m->height;
  const uint32_t f = m->filters;
  const int rowStart = tv * TILEVALID, rowEnd = (rowStart + TILESIZE < height) ? rowStart + TILESIZE : height;
  const int colStart = th * TILEVALID, colEnd = (colStart + TILESIZE < width) ? colStart + TILESIZE : width;
  const int tileRows = rowEnd - rowStart, tileCols = colEnd - colStart;
  const int last_rr = tileRows + 2 * BORDER, last_cc = tileCols + 2 * BORDER;
  const int w1 = GRP, w2 = 2 * GRP, w3 = 3 * GRP, w4 = 4 * GRP;

  /* :191-200 the gamma-encoded mosaic */
  for(int r = 0; r < tileRows; r++)
    for(int c = 0; c < tileCols; c++)
      qix[5][(r + BORDER) * GRP + c + BORDER] = calc_gamma(m->revscaler * m->in[(size_t)(rowStart + r) * width + colStart + c], m->gin);

  /* :202-236 G - R(B) along rows and columns */
  for(int rr = 2; rr < last_rr - 2; rr++)
  {
    for(int cc = 2 + (orc_fc(rr, 2, f) & 1); cc < last_cc - 2; cc += 2)
    { /* at red / blue sites */
      const float *cfa = qix[5] + rr * GRP + cc;
      const float v0 = 0.0625f * (cfa[-w1 - 1] + cfa[-w1 + 1] + cfa[w1 - 1] + cfa[w1 + 1]) + 0.25f * cfa[0];
      float h = -0.25f * (cfa[-2] + cfa[2]) + 0.5f * (cfa[-1] + cfa[0] + cfa[1]);
      const float Y0 = v0 + 0.5f * h;
      h = (cfa[0] > 1.75f * Y0) ? median3f(h, cfa[-1], cfa[1]) : limf(h, 0.0f, 1.0f);
      qix[0][rr * GRP + cc] = h - cfa[0];
      float v = -0.25f * (cfa[-w2] + cfa[w2]) + 0.5f * (cfa[-w1] + cfa[0] + cfa[w1]);
      const float Y1 = v0 + 0.5f * v;
      v = (cfa[0] > 1.75f * Y1) ? median3f(v, cfa[-w1], cfa[w1]) : limf(v, 0.0f, 1.0f);
      qix[1][rr * GRP + cc] = v - cfa[0];
    }
    for(int cc = 2 + (orc_fc(rr, 3, f) & 1); cc < last_cc - 2; cc += 2)
    { /* at green sites */
      const float *cfa = qix[5] + rr * GRP + cc;
      const float h = 0.25f * (cfa[-2] + cfa[2]) - 0.5f * (cfa[-1] + cfa[0] + cfa[1]);
      const float v = 0.25f * (cfa[-w2] + cfa[w2]) - 0.5f * (cfa[-w1] + cfa[0] + cfa[w1]);
      qix[0][rr * GRP + cc] = limf(h, -1.0f, 0.0f) + cfa[0];
      qix[1][rr * GRP + cc] = limf(v, -1.0f, 0.0f) + cfa[0];
    }
  }
  /* :238-250 low pass of the differences */
  for(int rr = 4; rr < last_rr - 4; rr++)
    for(int cc = 4; cc < last_cc - 4; cc++)
    {
      const float *hd = qix[0] + rr * GRP + cc, *vd = qix[1] + rr * GRP + cc;
      qix[2][rr * GRP + cc] = m->h0 * hd[0] + m->h1 * (hd[-1] + hd[1]) + m->h2 * (hd[-2] + hd[2]) + m->h3 * (hd[-3] + hd[3]) + m->h4 * (hd[-4] + hd[4]);
      qix[3][rr * GRP + cc] = m->h0 * vd[0] + m->h1 * (vd[-w1] + vd[w1]) + m->h2 * (vd[-w2] + vd[w2]) + m->h3 * (vd[-w3] + vd[w3]) + m->h4 * (vd[-w4] + vd[w4]);
    }
  /* :252-314 the interpolated G - R(B) at red / blue sites */
  for(int rr = 4; rr < last_rr - 4; rr++)
    for(int cc = 4 + (orc_fc(rr, 4, f) & 1); cc < last_cc - 4; cc += 2)
    {
      const int i = rr * GRP + cc;
      float xh, vh, xv, vv;
      lm_estimate(qix[2] + i, qix[0] + i, 1, &xh, &vh);
      lm_estimate(qix[3] + i, qix[1] + i, w1, &xv, &vv);
      qix[4][i] = (xh * vv + xv * vh) / (vh + vv);
    }
  /* :316-336 the colour planes: the mosaic in its own plane, green at red / blue sites, zero outside the frame */
  for(int rr = 0, row_in = rowStart - BORDER; rr < last_rr; rr++, row_in++)
    for(int cc = 0, col_in = colStart - BORDER; cc < last_cc; cc++, col_in++)
    {
      const int c = orc_fc(rr, cc, f), i = rr * GRP + cc;
      const int inside = row_in >= 0 && row_in < height && col_in >= 0 && col_in < width;
      qix[c][i] = inside ? qix[5][i] : 0.0f;
      if(c != 1) qix[1][i] = inside ? qix[c][i] + qix[4][i] : 0.0f;
    }
  /* :338-363 bilinear red / blue on the colour differences: at green sites, then at blue / red sites */
  for(int rr = 1; rr < last_rr - 1; rr++)
    for(int cc = 1 + (orc_fc(rr, 2, f) & 1), c = orc_fc(rr, cc + 1, f); cc < last_cc - 1; cc += 2)
    {
      const int i = rr * GRP + cc;
      const float *g = qix[1] + i;
      float *p = qix[c] + i;
      p[0] = g[0] + 0.5f * (p[-1] - g[-1] + p[1] - g[1]);
      p = qix[2 - c] + i;
      p[0] = g[0] + 0.5f * (p[-w1] - g[-w1] + p[w1] - g[w1]);
    }
  for(int rr = 1; rr < last_rr - 1; rr++)
    for(int cc = 1 + (orc_fc(rr, 1, f) & 1), c = 2 - orc_fc(rr, cc, f); cc < last_cc - 1; cc += 2)
    {
      const int i = rr * GRP + cc;
      const float *g = qix[1] + i;
      float *p = qix[c] + i;
      p[0] = g[0] + 0.25f * (p[-w1] - g[-w1] + p[-1] - g[-1] + p[1] - g[1] + p[w1] - g[w1]);
    }

  /* :365-370 */
  const int ccmin = th == 0 ? 6 : 0, ccmax = last_cc - (th == m->nh - 1 ? 6 : 0);
  const int rrmin = tv == 0 ? 6 : 0, rrmax = last_rr - (tv == m->nv - 1 ? 6 : 0);

  /* :372-478 median passes: 3x3 medians of R - G into plane 3 and of B - G into plane 4, then every site rebuilt from them */
  for(int pass = 0; pass < m->medians; pass++)
  {
    for(int rr = 1; rr < last_rr - 1; rr++)
      for(int c = 0; c < 3; c += 2)
      {
        const int d = c + 3 - (c == 0 ? 0 : 1);
        for(int cc = 1; cc < last_cc - 1; cc++)
        {
          const float *p = qix[c] + rr * GRP + cc, *g = qix[1] + rr * GRP + cc;
          qix[d][rr * GRP + cc] = median9f(p[-w1 - 1] - g[-w1 - 1], p[-w1] - g[-w1], p[-w1 + 1] - g[-w1 + 1], p[-1] - g[-1], p[0] - g[0], p[1] - g[1],
                                           p[w1 - 1] - g[w1 - 1], p[w1] - g[w1], p[w1 + 1] - g[w1 + 1]);
        }
      }
    for(int rr = rrmin; rr < rrmax - 1; rr++)
    { /* :394-477: the reference walks the row in pairs of sites from ccmin (even) and finishes a trailing single site with the pair's first
         operation: every site of [ccmin, ccmax) gets the operation of its parity.  Green sites: red and blue rebuilt from green and the medians;
         red / blue sites: the opposite colour from green and its median, then green from the two colours and the medians */
      const int c0 = orc_fc(rr, 0, f), c1 = orc_fc(rr, 1, f);
      for(int cc = ccmin; cc < ccmax; cc++)
      {
        const int i = rr * GRP + cc;
        const int first = ((cc - ccmin) & 1) == 0;
        if((c0 == 1) == first)
        {
          qix[0][i] = qix[1][i] + qix[3][i];
          qix[2][i] = qix[1][i] + qix[4][i];
        }
        else
        {
          const int c = c0 == 1 ? 2 - c1 : 2 - c0, d = c + 3 - (c == 0 ? 0 : 1);
          qix[c][i] = qix[1][i] + qix[d][i];
          qix[1][i] = 0.5f * (qix[0][i] - qix[3][i] + qix[2][i] - qix[4][i]);
        }
      }
    }
  }
  /* :480-489 the mosaic back into its own plane */
  for(int rr = 4; rr < last_rr - 4; rr++)
    for(int cc = 4; cc < last_cc - 4; cc++) qix[orc_fc(rr, cc, f)][rr * GRP + cc] = qix[5][rr * GRP + cc];

  /* :491-546 refinement steps */
  for(int step = 0; step < m->refine; step++)
  {
    for(int rr = rrmin + 2; rr < rrmax - 2; rr++)
      for(int cc = ccmin + 2 + (orc_fc(rr, 2, f) & 1), c = orc_fc(rr, cc, f); cc < ccmax - 2; cc += 2)
      { /* green at red / blue sites */
        float *g = qix[1] + rr * GRP + cc;
        const float *p = qix[c] + rr * GRP + cc;
        const float dL = 1.0f / (1.0f + fabsf(p[-2] - p[0]) + fabsf(g[1] - g[-1])), dR = 1.0f / (1.0f + fabsf(p[2] - p[0]) + fabsf(g[1] - g[-1]));
        const float dU = 1.0f / (1.0f + fabsf(p[-w2] - p[0]) + fabsf(g[w1] - g[-w1])), dD = 1.0f / (1.0f + fabsf(p[w2] - p[0]) + fabsf(g[w1] - g[-w1]));
        g[0] = (p[0] + ((g[-1] - p[-1]) * dL + (g[1] - p[1]) * dR + (g[-w1] - p[-w1]) * dU + (g[w1] - p[w1]) * dD) / (dL + dR + dU + dD));
      }
    for(int rr = rrmin + 2; rr < rrmax - 2; rr++)
      for(int cc = ccmin + 2 + (orc_fc(rr, 3, f) & 1), c = orc_fc(rr, cc + 1, f); cc < ccmax - 2; cc += 2)
        for(int i = 0; i < 2; c = 2 - c, i++)
        { /* red and blue at green sites */
          const float *g = qix[1] + rr * GRP + cc;
          float *p = qix[c] + rr * GRP + cc;
          const float dL = 1.0f / (1.0f + fabsf(g[-2] - g[0]) + fabsf(p[1] - p[-1])), dR = 1.0f / (1.0f + fabsf(g[2] - g[0]) + fabsf(p[1] - p[-1]));
          const float dU = 1.0f / (1.0f + fabsf(g[-w2] - g[0]) + fabsf(p[w1] - p[-w1])), dD = 1.0f / (1.0f + fabsf(g[w2] - g[0]) + fabsf(p[w1] - p[-w1]));
          p[0] = (g[0] - ((g[-1] - p[-1]) * dL + (g[1] - p[1]) * dR + (g[-w1] - p[-w1]) * dU + (g[w1] - p[w1]) * dD) / (dL + dR + dU + dD));
        }
    for(int rr = rrmin + 2; rr < rrmax - 2; rr++)
      for(int cc = ccmin + 2 + (orc_fc(rr, 2, f) & 1), c = 2 - orc_fc(rr, cc, f); cc < ccmax - 2; cc += 2)
      { /* the opposite colour at red / blue sites */
        const float *g = qix[1] + rr * GRP + cc, *q = qix[2 - c] + rr * GRP + cc;
        float *p = qix[c] + rr * GRP + cc;
        const float dL = 1.0f / (1.0f + fabsf(q[-2] - q[0]) + fabsf(g[1] - g[-1])), dR = 1.0f / (1.0f + fabsf(q[2] - q[0]) + fabsf(g[1] - g[-1]));
        const float dU = 1.0f / (1.0f + fabsf(q[-w2] - q[0]) + fabsf(g[w1] - g[-w1])), dD = 1.0f / (1.0f + fabsf(q[w2] - q[0]) + fabsf(g[w1] - g[-w1]));
        p[0] = (g[0] - ((g[-1] - p[-1]) * dL + (g[1] - p[1]) * dR + (g[-w1] - p[-w1]) * dU + (g[w1] - p[w1]) * dD) / (dL + dR + dU + dD));
      }
  }
  /* :548-570 the kept part of the tile, decoded */
  const int first_v = rowStart + (tv == 0 ? 0 : OVERLAP), last_v = rowEnd - (tv == m->nv - 1 ? 0 : OVERLAP);
  const int first_h = colStart + (th == 0 ? 0 : OVERLAP), last_h = colEnd - (th == m->nh - 1 ? 0 : OVERLAP);
  for(int row = first_v; row < last_v; row++)
    for(int col = first_h; col < last_h; col++)
    {
      const int i = (row - rowStart + BORDER) * GRP + col - colStart + BORDER;
      float *d = m->out + 4 * ((size_t)row * width + col);
      d[0] = m->scaler * calc_gamma(qix[0][i], m->gout);
      d[1] = m->scaler * calc_gamma(qix[1][i], m->gout);
      d[2] = m->scaler * calc_gamma(qix[2][i], m->gout);
      d[3] = 0.0f;
    }
}

/* lmmse_demosaic(): mode = dt_iop_demosaic_lmmse_t (0 basic, 1 median, 2 three medians, 3 / 4 one / two refinement steps on top);
 * carry: see the head of this file */
void orc_lmmse(float *out, const float *in, int width, int height, uint32_t filters, int mode, const float processed_maximum[3], int carry)
{
  if(width < 16 || height < 16) return; /* :140-144 */
  static float *gin = NULL, *gout = NULL;
  if(!gin)
  {
    gin = malloc(65536 * sizeof(float));
    gout = malloc(65536 * sizeof(float));
    orc_lmmse_gamma_tables(gin, gout);
  }
  lm_t m;
  m.in = in;
  m.out = out;
  m.width = width;
  m.height = height;
  m.filters = filters;
  m.gin = gin;
  m.gout = gout;
  float h0 = 1.0f, h1 = expf(-1.0f / 8.0f), h2 = expf(-4.0f / 8.0f), h3 = expf(-9.0f / 8.0f), h4 = expf(-16.0f / 8.0f);
  const float hs = h0 + 2.0f * (h1 + h2 + h3 + h4);
  m.h0 = h0 / hs;
  m.h1 = h1 / hs;
  m.h2 = h2 / hs;
  m.h3 = h3 / hs;
  m.h4 = h4 / hs;
  m.medians = mode < 2 ? mode : 3;
  m.refine = mode > 2 ? mode - 2 : 0;
  m.scaler = fmaxf(processed_maximum[0], fmaxf(processed_maximum[1], processed_maximum[2]));
  m.revscaler = 1.0f / m.scaler;
  m.nv = 1 + (height - 2 * OVERLAP - 1) / TILEVALID;
  m.nh = 1 + (width - 2 * OVERLAP - 1) / TILEVALID;
  orc_fp_fast_mode(); /* the pipe's threads run with FTZ|DAZ (darktable.c:877, common/dtpthread.c:54) */
  float *buffer = calloc((size_t)6 * NP, sizeof(float));
  float *qix[6];
  for(int i = 0; i < 6; i++) qix[i] = buffer + (size_t)i * NP;
  for(int tv = 0; tv < m.nv; tv++)
    for(int th = 0; th < m.nh; th++)
    {
      if(!carry) memset(buffer, 0, sizeof(float) * 6 * NP);
      lm_tile(&m, qix, tv, th);
    }
  free(buffer);
}
