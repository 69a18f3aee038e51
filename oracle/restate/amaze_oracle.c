/* CPU restatement of the AMaZE demosaicer as the reference runs it.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src/iop/demosaic/amaze.cc (RawTherapee's amaze_interpolate_RT as vendored there):
 * helpers clampnan :60-75, xmul2f/xdiv2f/xdivf :77-122, intp/LIM/ULIM :136-153; amaze_demosaic_RT :181-1419 --
 * constants :195-262, scratch layout :275-329, tile loop :334-336, mirrored 16-px border fill :357-455,
 * gradients :460-470, H/V colour differences :474-577, variance choice + saturation bounds :580-686, direction
 * weights :688-746, Nyquist test / area interpolation :748-884, green at R/B :888-912, Nyquist refinement :918-955,
 * diagonal gradients :957-981, diagonal R/B :986-1104, R+B mix :1106-1124, diagonal-corrected green :1127-1241,
 * chroma :1247-1289, output :1291-1407.
 *
 * Three things fix the result beyond the formulas and are restated exactly:
 *   - the 160-px tile grid (128 kept, origin -16): several planes are updated IN PLACE in raster order (hcd/vcd
 *     :596-684, hvwt :899, pmwt :1117), so a pixel depends on the tile it falls in and on the sweep order;
 *   - the scratch layout: planes alias by lifetime (:298-328);
 *   - the scratch is allocated once per thread and NOT cleared between tiles (:283), so a few results depend on
 *     what the previous tile of the same thread left behind -- thread-count dependent in the reference
 *     (measured: 5 pixels of a 1300x900 frame differ between 1 and 8 threads).  `scratch_mode` selects:
 *        0  one scratch carried from tile to tile in raster order  == the reference with OMP_NUM_THREADS=1 (the pin)
 *        1  scratch zeroed before every tile                        == what a GPU kernel with fresh scratch computes
 *        2  scratch filled with NaN before every tile               -> marks every result that reads stale scratch
 * Mixed precision of the source is kept: `0.5 - varwt`, `cfa * 2.0 / ...`, `2.0 * (...)` are double expressions.
 * Evaluated with flush-to-zero / denormals-are-zero like the reference's pipe threads; pinned bit-for-bit against
 * amaze.cc compiled in place and run in the same mode (oracle/_ref, ref_amaze.cc).
 */
#include "oracle_common.h"
#include <stdlib.h>
#include <string.h>

#define TS 160
#define TSH (TS / 2)

static inline float amz_min(float a, float b) { return (b < a) ? b : a; }      /* std::min */
static inline float amz_max(float a, float b) { return (a < b) ? b : a; }      /* std::max */
static inline float LIMF(float a, float b, float c) { return amz_max(b, amz_min(a, c)); }
static inline float ULIMF(float a, float b, float c) { return (b < c) ? LIMF(a, b, c) : LIMF(a, c, b); }
static inline float SQ(float x) { return x * x; }
static inline float mixf(float a, float b, float c) { return a * (b - c) + c; }
#define GMIN(a, b) (((a) < (b)) ? (a) : (b)) /* glib MIN */

static inline float clampnan(float x, float m, float M)
{ /* :60-75 */
  float r;
  if(!isfinite(x))
    r = (isless(x, m) ? m : (isgreater(x, M) ? M : x));
  else if(isnan(x))
    r = (m + M) / 2.0f;
  else
    r = x;
  return r;
}
static inline float exp_add(float d, int n)
{ /* xmul2f / xdiv2f / xdivf :77-122: add n to the exponent field unless the value is +-0 */
  uint32_t u;
  memcpy(&u, &d, 4);
  if(u & 0x7FFFFFFFu) u += (uint32_t)(n << 23);
  memcpy(&d, &u, 4);
  return d;
}
#define XMUL2(x) exp_add((x), 1)
#define XDIV2(x) exp_add((x), -1)
#define XDIV4(x) exp_add((x), -2)

typedef struct
{
  float h, v;
} hv_t;

static void fill_scratch(char *buffer, size_t bytes, int mode)
{
  if(mode == 2)
  {
    uint32_t *w = (uint32_t *)buffer;
    for(size_t k = 0; k < bytes / 4; k++) w[k] = 0x7fc00000u;
  }
  else
    memset(buffer, 0, bytes);
}

int orc_amaze_demosaic(float *out, const float *in, int width, int height, uint32_t filters, const float processed_maximum[3], int scratch_mode)
{
  orc_fp_fast_mode(); /* the pipe's threads run with FTZ|DAZ (darktable.c:877, common/dtpthread.c:54) */
  const int winx = 0, winy = 0;
  const float clip_pt = fminf(processed_maximum[0], fminf(processed_maximum[1], processed_maximum[2]));
  const float clip_pt8 = 0.8f * clip_pt;
  const int ts = TS, tsh = TSH;

  int ex, ey; /* offset of the R site in a Bayer quartet, :206-234 */
  if(orc_fc(0, 0, filters) == 1)
  {
    if(orc_fc(0, 1, filters) == 0)
    {
      ey = 0;
      ex = 1;
    }
    else
    {
      ey = 1;
      ex = 0;
    }
  }
  else if(orc_fc(0, 0, filters) == 0)
  {
    ey = 0;
    ex = 0;
  }
  else
  {
    ey = 1;
    ex = 1;
  }
  const int v1 = ts, v2 = 2 * ts, v3 = 3 * ts, p1 = -ts + 1, p2 = -2 * ts + 2, p3 = -3 * ts + 3, m1 = ts + 1, m2 = 2 * ts + 2, m3 = 3 * ts + 3;
  const float eps = 1e-5, epssq = 1e-10, arthresh = 0.75;
  static const float gaussodd[4] = { 0.14659727707323927f, 0.103592713382435f, 0.0732036125103057f, 0.0365543548389495f };
  const float nyqthresh = 0.5;
  const float gaussgrad[6] = { nyqthresh * 0.07384411893421103f, nyqthresh * 0.06207511968171489f, nyqthresh * 0.0521818194747806f,
                               nyqthresh * 0.03687419286733595f, nyqthresh * 0.03099732204057846f, nyqthresh * 0.018413194161458882f };
  static const float gausseven[2] = { 0.13719494435797422f, 0.05640252782101291f };
  static const float gquinc[4] = { 0.169917f, 0.108947f, 0.069855f, 0.0287182f };

  /* scratch layout, :275-329 (cldf = 2: 128 bytes between planes) */
  const size_t pad = 2 * 64;
  const size_t bytes = sizeof(float) * 14 * ts * ts + sizeof(char) * ts * tsh + 18 * pad + 63;
  char *buffer = (char *)calloc(bytes, 1);
  if(!buffer) return 1;
  char *data = (char *)(((uintptr_t)buffer + (uintptr_t)63) / 64 * 64);
  const size_t full = sizeof(float) * ts * ts, half = sizeof(float) * ts * tsh;
  float *rgbgreen = (float *)data;
  float *delhvsqsum = (float *)((char *)rgbgreen + full + pad);
  float *dirwts0 = (float *)((char *)delhvsqsum + full + pad);
  float *dirwts1 = (float *)((char *)dirwts0 + full + pad);
  float *vcd = (float *)((char *)dirwts1 + full + pad);
  float *hcd = (float *)((char *)vcd + full + pad);
  float *vcdalt = (float *)((char *)hcd + full + pad);
  float *hcdalt = (float *)((char *)vcdalt + full + pad);
  float *cddiffsq = (float *)((char *)hcdalt + full + pad);
  float *hvwt = (float *)((char *)cddiffsq + full + 2 * pad);
  float *Dgrb0 = vcdalt, *Dgrb1 = vcdalt + ts * tsh; /* float (*Dgrb)[ts*tsh] over vcdalt */
  float *delp = cddiffsq;
  float *delm = (float *)((char *)delp + half + pad);
  float *rbint = delm;
  hv_t *Dgrb2 = (hv_t *)((char *)hvwt + half + pad);
  float *dgintv = (float *)Dgrb2;
  float *dginth = (float *)((char *)dgintv + full + pad);
  float *Dgrbsq1m = (float *)((char *)dginth + full + pad);
  float *Dgrbsq1p = (float *)((char *)Dgrbsq1m + half + pad);
  float *cfa = (float *)((char *)Dgrbsq1p + half + pad);
  float *pmwt = delhvsqsum;
  float *rbm = vcd;
  float *rbp = (float *)((char *)rbm + half + pad);
  unsigned char *nyquist = (unsigned char *)((char *)cfa + full + pad);
  unsigned char *nyquist2 = (unsigned char *)cddiffsq;
  float *nyqutest = (float *)((char *)nyquist + sizeof(unsigned char) * ts * tsh + pad);

  for(int top = winy - 16; top < winy + height; top += ts - 32)
    for(int left = winx - 16; left < winx + width; left += ts - 32)
    {
      if(scratch_mode) fill_scratch(buffer, bytes, scratch_mode);
      memset(&nyquist[3 * tsh], 0, sizeof(unsigned char) * (ts - 6) * tsh);
      const int bottom = GMIN(top + ts, winy + height + 16);
      const int right = GMIN(left + ts, winx + width + 16);
      const int rr1 = bottom - top, cc1 = right - left;
      const int rrmin = top < winy ? 16 : 0, ccmin = left < winx ? 16 : 0;
      const int rrmax = bottom > (winy + height) ? winy + height - top : rr1;
      const int ccmax = right > (winx + width) ? winx + width - left : cc1;

      /* ---- tile load with a mirrored 16-px border at the frame edges, :357-455 ---- */
#define PUT(rr_, cc_, v_)                     \
  do                                          \
  {                                           \
    const float t_ = (v_);                    \
    cfa[(rr_) * ts + (cc_)] = t_;             \
    rgbgreen[(rr_) * ts + (cc_)] = t_;        \
  } while(0)
      if(rrmin > 0)
        for(int rr = 0; rr < 16; rr++)
          for(int cc = ccmin, row = 32 - rr + top; cc < ccmax; cc++) PUT(rr, cc, in[row * width + (cc + left)]);
      for(int rr = rrmin; rr < rrmax; rr++)
        for(int cc = ccmin; cc < ccmax; cc++) PUT(rr, cc, in[(rr + top) * width + (cc + left)]);
      if(rrmax < rr1)
        for(int rr = 0; rr < 16; rr++)
          for(int cc = ccmin; cc < ccmax; cc++) PUT(rrmax + rr, cc, in[(winy + height - rr - 2) * width + (left + cc)]);
      if(ccmin > 0)
        for(int rr = rrmin; rr < rrmax; rr++)
          for(int cc = 0, row = rr + top; cc < 16; cc++) PUT(rr, cc, in[row * width + (32 - cc + left)]);
      if(ccmax < cc1)
        for(int rr = rrmin; rr < rrmax; rr++)
          for(int cc = 0; cc < 16; cc++) PUT(rr, ccmax + cc, in[(top + rr) * width + ((winx + width - cc - 2))]);
      if(rrmin > 0 && ccmin > 0)
        for(int rr = 0; rr < 16; rr++)
          for(int cc = 0; cc < 16; cc++) PUT(rr, cc, in[(winy + 32 - rr) * width + (winx + 32 - cc)]);
      if(rrmax < rr1 && ccmax < cc1)
        for(int rr = 0; rr < 16; rr++)
          for(int cc = 0; cc < 16; cc++) PUT(rrmax + rr, ccmax + cc, in[(winy + height - rr - 2) * width + ((winx + width - cc - 2))]);
      if(rrmin > 0 && ccmax < cc1)
        for(int rr = 0; rr < 16; rr++)
          for(int cc = 0; cc < 16; cc++) PUT(rr, ccmax + cc, in[(winy + 32 - rr) * width + ((winx + width - cc - 2))]);
      if(rrmax < rr1 && ccmin > 0)
        for(int rr = 0; rr < 16; rr++)
          for(int cc = 0; cc < 16; cc++) PUT(rrmax + rr, cc, in[(winy + height - rr - 2) * width + ((winx + 32 - cc))]);
#undef PUT

      /* ---- gradients and direction weights, :460-470 ---- */
      for(int rr = 2; rr < rr1 - 2; rr++)
        for(int cc = 2, i = rr * ts + cc; cc < cc1 - 2; cc++, i++)
        {
          const float delh = fabsf(cfa[i + 1] - cfa[i - 1]);
          const float delv = fabsf(cfa[i + v1] - cfa[i - v1]);
          dirwts0[i] = eps + fabsf(cfa[i + v2] - cfa[i]) + fabsf(cfa[i] - cfa[i - v2]) + delv;
          dirwts1[i] = eps + fabsf(cfa[i + 2] - cfa[i]) + fabsf(cfa[i] - cfa[i - 2]) + delh;
          delhvsqsum[i] = SQ(delh) + SQ(delv);
        }

      /* ---- vertical / horizontal colour differences, :474-577 ---- */
      for(int rr = 4; rr < rr1 - 4; rr++)
      {
        int fcswitch = orc_fc(rr, 4, filters) & 1;
        for(int cc = 4, i = rr * ts + cc; cc < cc1 - 4; cc++, i++)
        {
          const float c0 = cfa[i];
          const float cru = cfa[i - v1] * (dirwts0[i - v2] + dirwts0[i]) / (dirwts0[i - v2] * (eps + c0) + dirwts0[i] * (eps + cfa[i - v2]));
          const float crd = cfa[i + v1] * (dirwts0[i + v2] + dirwts0[i]) / (dirwts0[i + v2] * (eps + c0) + dirwts0[i] * (eps + cfa[i + v2]));
          const float crl = cfa[i - 1] * (dirwts1[i - 2] + dirwts1[i]) / (dirwts1[i - 2] * (eps + c0) + dirwts1[i] * (eps + cfa[i - 2]));
          const float crr = cfa[i + 1] * (dirwts1[i + 2] + dirwts1[i]) / (dirwts1[i + 2] * (eps + c0) + dirwts1[i] * (eps + cfa[i + 2]));
          const float guha = cfa[i - v1] + XDIV2(c0 - cfa[i - v2]);
          const float gdha = cfa[i + v1] + XDIV2(c0 - cfa[i + v2]);
          const float glha = cfa[i - 1] + XDIV2(c0 - cfa[i - 2]);
          const float grha = cfa[i + 1] + XDIV2(c0 - cfa[i + 2]);
          float guar = (fabsf(1.f - cru) < arthresh) ? c0 * cru : guha;
          float gdar = (fabsf(1.f - crd) < arthresh) ? c0 * crd : gdha;
          float glar = (fabsf(1.f - crl) < arthresh) ? c0 * crl : glha;
          float grar = (fabsf(1.f - crr) < arthresh) ? c0 * crr : grha;
          const float hwt = dirwts1[i - 1] / (dirwts1[i - 1] + dirwts1[i + 1]);
          const float vwt = dirwts0[i - v1] / (dirwts0[i + v1] + dirwts0[i - v1]);
          const float Gintvha = vwt * gdha + (1.f - vwt) * guha;
          const float Ginthha = hwt * grha + (1.f - hwt) * glha;
          if(fcswitch)
          {
            vcd[i] = c0 - (vwt * gdar + (1.f - vwt) * guar);
            hcd[i] = c0 - (hwt * grar + (1.f - hwt) * glar);
            vcdalt[i] = c0 - Gintvha;
            hcdalt[i] = c0 - Ginthha;
          }
          else
          {
            vcd[i] = (vwt * gdar + (1.f - vwt) * guar) - c0;
            hcd[i] = (hwt * grar + (1.f - hwt) * glar) - c0;
            vcdalt[i] = Gintvha - c0;
            hcdalt[i] = Ginthha - c0;
          }
          fcswitch = !fcswitch;
          if(c0 > clip_pt8 || Gintvha > clip_pt8 || Ginthha > clip_pt8)
          {
            guar = guha;
            gdar = gdha;
            glar = glha;
            grar = grha;
            vcd[i] = vcdalt[i];
            hcd[i] = hcdalt[i];
          }
          dgintv[i] = GMIN(SQ(guha - gdha), SQ(guar - gdar));
          dginth[i] = GMIN(SQ(glha - grha), SQ(glar - grar));
        }
      }

      /* ---- smaller-variance choice and saturation bounds, IN PLACE in raster order, :580-686 ---- */
      for(int rr = 4; rr < rr1 - 4; rr++)
        for(int cc = 4, i = rr * ts + cc, c = orc_fc(rr, cc, filters) & 1; cc < cc1 - 4; cc++, i++)
        {
          const float hcdvar = 3.f * (SQ(hcd[i - 2]) + SQ(hcd[i]) + SQ(hcd[i + 2])) - SQ(hcd[i - 2] + hcd[i] + hcd[i + 2]);
          const float hcdaltvar = 3.f * (SQ(hcdalt[i - 2]) + SQ(hcdalt[i]) + SQ(hcdalt[i + 2])) - SQ(hcdalt[i - 2] + hcdalt[i] + hcdalt[i + 2]);
          const float vcdvar = 3.f * (SQ(vcd[i - v2]) + SQ(vcd[i]) + SQ(vcd[i + v2])) - SQ(vcd[i - v2] + vcd[i] + vcd[i + v2]);
          const float vcdaltvar = 3.f * (SQ(vcdalt[i - v2]) + SQ(vcdalt[i]) + SQ(vcdalt[i + v2])) - SQ(vcdalt[i - v2] + vcdalt[i] + vcdalt[i + v2]);
          if(hcdaltvar < hcdvar) hcd[i] = hcdalt[i];
          if(vcdaltvar < vcdvar) vcd[i] = vcdalt[i];
          float Gintv, Ginth;
          if(c)
          { /* G site */
            Ginth = -hcd[i] + cfa[i];
            Gintv = -vcd[i] + cfa[i];
            if(hcd[i] > 0)
            {
              if(3.f * hcd[i] > (Ginth + cfa[i]))
                hcd[i] = -ULIMF(Ginth, cfa[i - 1], cfa[i + 1]) + cfa[i];
              else
              {
                const float hwt = 1.f - 3.f * hcd[i] / (eps + Ginth + cfa[i]);
                hcd[i] = hwt * hcd[i] + (1.f - hwt) * (-ULIMF(Ginth, cfa[i - 1], cfa[i + 1]) + cfa[i]);
              }
            }
            if(vcd[i] > 0)
            {
              if(3.f * vcd[i] > (Gintv + cfa[i]))
                vcd[i] = -ULIMF(Gintv, cfa[i - v1], cfa[i + v1]) + cfa[i];
              else
              {
                const float vwt = 1.f - 3.f * vcd[i] / (eps + Gintv + cfa[i]);
                vcd[i] = vwt * vcd[i] + (1.f - vwt) * (-ULIMF(Gintv, cfa[i - v1], cfa[i + v1]) + cfa[i]);
              }
            }
            if(Ginth > clip_pt) hcd[i] = -ULIMF(Ginth, cfa[i - 1], cfa[i + 1]) + cfa[i];
            if(Gintv > clip_pt) vcd[i] = -ULIMF(Gintv, cfa[i - v1], cfa[i + v1]) + cfa[i];
          }
          else
          { /* R or B site */
            Ginth = hcd[i] + cfa[i];
            Gintv = vcd[i] + cfa[i];
            if(hcd[i] < 0)
            {
              if(3.f * hcd[i] < -(Ginth + cfa[i]))
                hcd[i] = ULIMF(Ginth, cfa[i - 1], cfa[i + 1]) - cfa[i];
              else
              {
                const float hwt = 1.f + 3.f * hcd[i] / (eps + Ginth + cfa[i]);
                hcd[i] = hwt * hcd[i] + (1.f - hwt) * (ULIMF(Ginth, cfa[i - 1], cfa[i + 1]) - cfa[i]);
              }
            }
            if(vcd[i] < 0)
            {
              if(3.f * vcd[i] < -(Gintv + cfa[i]))
                vcd[i] = ULIMF(Gintv, cfa[i - v1], cfa[i + v1]) - cfa[i];
              else
              {
                const float vwt = 1.f + 3.f * vcd[i] / (eps + Gintv + cfa[i]);
                vcd[i] = vwt * vcd[i] + (1.f - vwt) * (ULIMF(Gintv, cfa[i - v1], cfa[i + v1]) - cfa[i]);
              }
            }
            if(Ginth > clip_pt) hcd[i] = ULIMF(Ginth, cfa[i - 1], cfa[i + 1]) - cfa[i];
            if(Gintv > clip_pt) vcd[i] = ULIMF(Gintv, cfa[i - v1], cfa[i + v1]) - cfa[i];
            cddiffsq[i] = SQ(vcd[i] - hcd[i]);
          }
          c = !c;
        }

      /* ---- adaptive H/V weight at R/B sites, :688-746 ---- */
      for(int rr = 6; rr < rr1 - 6; rr++)
        for(int cc = 6 + (orc_fc(rr, 2, filters) & 1), i = rr * ts + cc; cc < cc1 - 6; cc += 2, i += 2)
        {
          const float uave = vcd[i] + vcd[i - v1] + vcd[i - v2] + vcd[i - v3];
          const float dave = vcd[i] + vcd[i + v1] + vcd[i + v2] + vcd[i + v3];
          const float lave = hcd[i] + hcd[i - 1] + hcd[i - 2] + hcd[i - 3];
          const float rave = hcd[i] + hcd[i + 1] + hcd[i + 2] + hcd[i + 3];
          float Dgrbvvaru = SQ(vcd[i] - uave) + SQ(vcd[i - v1] - uave) + SQ(vcd[i - v2] - uave) + SQ(vcd[i - v3] - uave);
          float Dgrbvvard = SQ(vcd[i] - dave) + SQ(vcd[i + v1] - dave) + SQ(vcd[i + v2] - dave) + SQ(vcd[i + v3] - dave);
          float Dgrbhvarl = SQ(hcd[i] - lave) + SQ(hcd[i - 1] - lave) + SQ(hcd[i - 2] - lave) + SQ(hcd[i - 3] - lave);
          float Dgrbhvarr = SQ(hcd[i] - rave) + SQ(hcd[i + 1] - rave) + SQ(hcd[i + 2] - rave) + SQ(hcd[i + 3] - rave);
          const float hwt = dirwts1[i - 1] / (dirwts1[i - 1] + dirwts1[i + 1]);
          const float vwt = dirwts0[i - v1] / (dirwts0[i + v1] + dirwts0[i - v1]);
          const float vcdvar = epssq + vwt * Dgrbvvard + (1.f - vwt) * Dgrbvvaru;
          const float hcdvar = epssq + hwt * Dgrbhvarr + (1.f - hwt) * Dgrbhvarl;
          Dgrbvvaru = (dgintv[i]) + (dgintv[i - v1]) + (dgintv[i - v2]);
          Dgrbvvard = (dgintv[i]) + (dgintv[i + v1]) + (dgintv[i + v2]);
          Dgrbhvarl = (dginth[i]) + (dginth[i - 1]) + (dginth[i - 2]);
          Dgrbhvarr = (dginth[i]) + (dginth[i + 1]) + (dginth[i + 2]);
          const float vcdvar1 = epssq + vwt * Dgrbvvard + (1.f - vwt) * Dgrbvvaru;
          const float hcdvar1 = epssq + hwt * Dgrbhvarr + (1.f - hwt) * Dgrbhvarl;
          const float varwt = hcdvar / (vcdvar + hcdvar);
          const float diffwt = hcdvar1 / (vcdvar1 + hcdvar1);
          if((0.5 - varwt) * (0.5 - diffwt) > 0 && fabsf(0.5f - diffwt) < fabsf(0.5f - varwt))
            hvwt[i >> 1] = varwt;
          else
            hvwt[i >> 1] = diffwt;
        }

      /* ---- Nyquist texture test, :748-815 ---- */
      for(int rr = 6; rr < rr1 - 6; rr++)
        for(int cc = 6 + (orc_fc(rr, 2, filters) & 1), i = rr * ts + cc; cc < cc1 - 6; cc += 2, i += 2)
          nyqutest[i >> 1]
              = (gaussodd[0] * cddiffsq[i] + gaussodd[1] * (cddiffsq[(i - m1)] + cddiffsq[(i + p1)] + cddiffsq[(i - p1)] + cddiffsq[(i + m1)])
                 + gaussodd[2] * (cddiffsq[(i - v2)] + cddiffsq[(i - 2)] + cddiffsq[(i + 2)] + cddiffsq[(i + v2)])
                 + gaussodd[3] * (cddiffsq[(i - m2)] + cddiffsq[(i + p2)] + cddiffsq[(i - p2)] + cddiffsq[(i + m2)]))
                - (gaussgrad[0] * delhvsqsum[i] + gaussgrad[1] * (delhvsqsum[i - v1] + delhvsqsum[i + 1] + delhvsqsum[i - 1] + delhvsqsum[i + v1])
                   + gaussgrad[2] * (delhvsqsum[i - m1] + delhvsqsum[i + p1] + delhvsqsum[i - p1] + delhvsqsum[i + m1])
                   + gaussgrad[3] * (delhvsqsum[i - v2] + delhvsqsum[i - 2] + delhvsqsum[i + 2] + delhvsqsum[i + v2])
                   + gaussgrad[4]
                         * (delhvsqsum[i - v2 - 1] + delhvsqsum[i - v2 + 1] + delhvsqsum[i - ts - 2] + delhvsqsum[i - ts + 2] + delhvsqsum[i + ts - 2]
                            + delhvsqsum[i + ts + 2] + delhvsqsum[i + v2 - 1] + delhvsqsum[i + v2 + 1])
                   + gaussgrad[5] * (delhvsqsum[i - m2] + delhvsqsum[i + p2] + delhvsqsum[i - p2] + delhvsqsum[i + m2]));
      int nystartrow = 0, nyendrow = 0, nystartcol = ts + 1, nyendcol = 0;
      for(int rr = 6; rr < rr1 - 6; rr++)
        for(int cc = 6 + (orc_fc(rr, 2, filters) & 1), i = rr * ts + cc; cc < cc1 - 6; cc += 2, i += 2)
          if(nyqutest[i >> 1] > 0.f)
          {
            nyquist[i >> 1] = 1;
            nystartrow = nystartrow ? nystartrow : rr;
            nyendrow = rr;
            nystartcol = nystartcol > cc ? cc : nystartcol;
            nyendcol = nyendcol < cc ? cc : nyendcol;
          }
      const int doNyquist = nystartrow != nyendrow && nystartcol != nyendcol;
      if(doNyquist)
      { /* :819-884 */
        nyendrow++;
        nyendcol++;
        nystartcol -= (nystartcol & 1);
        nystartrow = nystartrow > 8 ? nystartrow : 8;
        nyendrow = (rr1 - 8) < nyendrow ? (rr1 - 8) : nyendrow;
        nystartcol = nystartcol > 8 ? nystartcol : 8;
        nyendcol = (cc1 - 8) < nyendcol ? (cc1 - 8) : nyendcol;
        memset(&nyquist2[4 * tsh], 0, sizeof(char) * (ts - 8) * tsh);
        for(int rr = nystartrow; rr < nyendrow; rr++)
          for(int i = rr * ts + nystartcol + (orc_fc(rr, 2, filters) & 1); i < rr * ts + nyendcol; i += 2)
          {
            const unsigned int t = (nyquist[(i - v2) >> 1] + nyquist[(i - m1) >> 1] + nyquist[(i + p1) >> 1] + nyquist[(i - 2) >> 1]
                                    + nyquist[(i + 2) >> 1] + nyquist[(i - p1) >> 1] + nyquist[(i + m1) >> 1] + nyquist[(i + v2) >> 1]);
            nyquist2[i >> 1] = t > 4 ? 1 : (t < 4 ? 0 : nyquist[i >> 1]);
          }
        for(int rr = nystartrow; rr < nyendrow; rr++)
          for(int i = rr * ts + nystartcol + (orc_fc(rr, 2, filters) & 1); i < rr * ts + nyendcol; i += 2)
            if(nyquist2[i >> 1])
            { /* area interpolation */
              float sumcfa = 0.f, sumh = 0.f, sumv = 0.f, sumsqh = 0.f, sumsqv = 0.f, areawt = 0.f;
              for(int a = -6; a < 7; a += 2)
              {
                int i1 = i + (a * ts) - 6;
                for(int b = -6; b < 7; b += 2, i1 += 2)
                  if(nyquist2[i1 >> 1])
                  {
                    const float ct = cfa[i1];
                    sumcfa += ct;
                    sumh += (cfa[i1 - 1] + cfa[i1 + 1]);
                    sumv += (cfa[i1 - v1] + cfa[i1 + v1]);
                    sumsqh += SQ(ct - cfa[i1 - 1]) + SQ(ct - cfa[i1 + 1]);
                    sumsqv += SQ(ct - cfa[i1 - v1]) + SQ(ct - cfa[i1 + v1]);
                    areawt += 1;
                  }
              }
              sumh = sumcfa - XDIV2(sumh);
              sumv = sumcfa - XDIV2(sumv);
              areawt = XDIV2(areawt);
              const float hcdvar = epssq + fabsf(areawt * sumsqh - sumh * sumh);
              const float vcdvar = epssq + fabsf(areawt * sumsqv - sumv * sumv);
              hvwt[i >> 1] = hcdvar / (vcdvar + hcdvar);
            }
      }

      /* ---- green at R/B sites; hvwt is refined IN PLACE in raster order, :888-912 ---- */
      for(int rr = 8; rr < rr1 - 8; rr++)
        for(int i = rr * ts + 8 + (orc_fc(rr, 2, filters) & 1); i < rr * ts + cc1 - 8; i += 2)
        {
          const float hvwtalt = XDIV4(hvwt[(i - m1) >> 1] + hvwt[(i + p1) >> 1] + hvwt[(i - p1) >> 1] + hvwt[(i + m1) >> 1]);
          hvwt[i >> 1] = fabsf(0.5f - hvwt[i >> 1]) < fabsf(0.5f - hvwtalt) ? hvwtalt : hvwt[i >> 1];
          Dgrb0[i >> 1] = mixf(hvwt[i >> 1], vcd[i], hcd[i]);
          rgbgreen[i] = cfa[i] + Dgrb0[i >> 1];
          Dgrb2[i >> 1].h = nyquist2[i >> 1] ? SQ(rgbgreen[i] - XDIV2(rgbgreen[i - 1] + rgbgreen[i + 1])) : 0.f;
          Dgrb2[i >> 1].v = nyquist2[i >> 1] ? SQ(rgbgreen[i] - XDIV2(rgbgreen[i - v1] + rgbgreen[i + v1])) : 0.f;
        }

      /* ---- Nyquist refinement with green curvatures, :918-955 ---- */
      if(doNyquist)
        for(int rr = nystartrow; rr < nyendrow; rr++)
          for(int i = rr * ts + nystartcol + (orc_fc(rr, 2, filters) & 1); i < rr * ts + nyendcol; i += 2)
            if(nyquist2[i >> 1])
            {
              const float gvarh
                  = epssq
                    + (gquinc[0] * Dgrb2[i >> 1].h
                       + gquinc[1] * (Dgrb2[(i - m1) >> 1].h + Dgrb2[(i + p1) >> 1].h + Dgrb2[(i - p1) >> 1].h + Dgrb2[(i + m1) >> 1].h)
                       + gquinc[2] * (Dgrb2[(i - v2) >> 1].h + Dgrb2[(i - 2) >> 1].h + Dgrb2[(i + 2) >> 1].h + Dgrb2[(i + v2) >> 1].h)
                       + gquinc[3] * (Dgrb2[(i - m2) >> 1].h + Dgrb2[(i + p2) >> 1].h + Dgrb2[(i - p2) >> 1].h + Dgrb2[(i + m2) >> 1].h));
              const float gvarv
                  = epssq
                    + (gquinc[0] * Dgrb2[i >> 1].v
                       + gquinc[1] * (Dgrb2[(i - m1) >> 1].v + Dgrb2[(i + p1) >> 1].v + Dgrb2[(i - p1) >> 1].v + Dgrb2[(i + m1) >> 1].v)
                       + gquinc[2] * (Dgrb2[(i - v2) >> 1].v + Dgrb2[(i - 2) >> 1].v + Dgrb2[(i + 2) >> 1].v + Dgrb2[(i + v2) >> 1].v)
                       + gquinc[3] * (Dgrb2[(i - m2) >> 1].v + Dgrb2[(i + p2) >> 1].v + Dgrb2[(i - p2) >> 1].v + Dgrb2[(i + m2) >> 1].v));
              Dgrb0[i >> 1] = (hcd[i] * gvarv + vcd[i] * gvarh) / (gvarv + gvarh);
              rgbgreen[i] = cfa[i] + Dgrb0[i >> 1];
            }

      /* ---- diagonal gradients, :957-981 ---- */
      for(int rr = 6; rr < rr1 - 6; rr++)
      {
        if((orc_fc(rr, 2, filters) & 1) == 0)
          for(int cc = 6, i = rr * ts + cc; cc < cc1 - 6; cc += 2, i += 2)
          {
            delp[i >> 1] = fabsf(cfa[i + p1] - cfa[i - p1]);
            delm[i >> 1] = fabsf(cfa[i + m1] - cfa[i - m1]);
            Dgrbsq1p[i >> 1] = (SQ(cfa[i + 1] - cfa[i + 1 - p1]) + SQ(cfa[i + 1] - cfa[i + 1 + p1]));
            Dgrbsq1m[i >> 1] = (SQ(cfa[i + 1] - cfa[i + 1 - m1]) + SQ(cfa[i + 1] - cfa[i + 1 + m1]));
          }
        else
          for(int cc = 6, i = rr * ts + cc; cc < cc1 - 6; cc += 2, i += 2)
          {
            Dgrbsq1p[i >> 1] = (SQ(cfa[i] - cfa[i - p1]) + SQ(cfa[i] - cfa[i + p1]));
            Dgrbsq1m[i >> 1] = (SQ(cfa[i] - cfa[i - m1]) + SQ(cfa[i] - cfa[i + m1]));
            delp[i >> 1] = fabsf(cfa[i + 1 + p1] - cfa[i + 1 - p1]);
            delm[i >> 1] = fabsf(cfa[i + 1 + m1] - cfa[i + 1 - m1]);
          }
      }

      /* ---- diagonal interpolation of the opposite colour, :986-1104 ---- */
      for(int rr = 8; rr < rr1 - 8; rr++)
        for(int cc = 8 + (orc_fc(rr, 2, filters) & 1), i = rr * ts + cc, j = i >> 1; cc < cc1 - 8; cc += 2, i += 2, j++)
        {
          const float crse = XMUL2(cfa[i + m1]) / (eps + cfa[i] + (cfa[i + m2]));
          const float crnw = XMUL2(cfa[i - m1]) / (eps + cfa[i] + (cfa[i - m2]));
          const float crne = XMUL2(cfa[i + p1]) / (eps + cfa[i] + (cfa[i + p2]));
          const float crsw = XMUL2(cfa[i - p1]) / (eps + cfa[i] + (cfa[i - p2]));
          const float rbse = (fabsf(1.f - crse) < arthresh) ? cfa[i] * crse : (cfa[i + m1]) + XDIV2(cfa[i] - cfa[i + m2]);
          const float rbnw = (fabsf(1.f - crnw) < arthresh) ? cfa[i] * crnw : (cfa[i - m1]) + XDIV2(cfa[i] - cfa[i - m2]);
          const float rbne = (fabsf(1.f - crne) < arthresh) ? cfa[i] * crne : (cfa[i + p1]) + XDIV2(cfa[i] - cfa[i + p2]);
          const float rbsw = (fabsf(1.f - crsw) < arthresh) ? cfa[i] * crsw : (cfa[i - p1]) + XDIV2(cfa[i] - cfa[i - p2]);
          const float wtse = eps + delm[j] + delm[(i + m1) >> 1] + delm[(i + m2) >> 1];
          const float wtnw = eps + delm[j] + delm[(i - m1) >> 1] + delm[(i - m2) >> 1];
          const float wtne = eps + delp[j] + delp[(i + p1) >> 1] + delp[(i + p2) >> 1];
          const float wtsw = eps + delp[j] + delp[(i - p1) >> 1] + delp[(i - p2) >> 1];
          rbm[j] = (wtse * rbnw + wtnw * rbse) / (wtse + wtnw);
          rbp[j] = (wtne * rbsw + wtsw * rbne) / (wtne + wtsw);
          const float rbvarm
              = epssq
                + (gausseven[0] * (Dgrbsq1m[(i - v1) >> 1] + Dgrbsq1m[(i - 1) >> 1] + Dgrbsq1m[(i + 1) >> 1] + Dgrbsq1m[(i + v1) >> 1])
                   + gausseven[1]
                         * (Dgrbsq1m[(i - v2 - 1) >> 1] + Dgrbsq1m[(i - v2 + 1) >> 1] + Dgrbsq1m[(i - 2 - v1) >> 1] + Dgrbsq1m[(i + 2 - v1) >> 1]
                            + Dgrbsq1m[(i - 2 + v1) >> 1] + Dgrbsq1m[(i + 2 + v1) >> 1] + Dgrbsq1m[(i + v2 - 1) >> 1] + Dgrbsq1m[(i + v2 + 1) >> 1]));
          pmwt[j] = rbvarm
                    / ((epssq
                        + (gausseven[0] * (Dgrbsq1p[(i - v1) >> 1] + Dgrbsq1p[(i - 1) >> 1] + Dgrbsq1p[(i + 1) >> 1] + Dgrbsq1p[(i + v1) >> 1])
                           + gausseven[1]
                                 * (Dgrbsq1p[(i - v2 - 1) >> 1] + Dgrbsq1p[(i - v2 + 1) >> 1] + Dgrbsq1p[(i - 2 - v1) >> 1]
                                    + Dgrbsq1p[(i + 2 - v1) >> 1] + Dgrbsq1p[(i - 2 + v1) >> 1] + Dgrbsq1p[(i + 2 + v1) >> 1]
                                    + Dgrbsq1p[(i + v2 - 1) >> 1] + Dgrbsq1p[(i + v2 + 1) >> 1])))
                       + rbvarm);
          if(rbp[j] < cfa[i])
          {
            if(XMUL2(rbp[j]) < cfa[i])
              rbp[j] = ULIMF(rbp[j], cfa[i - p1], cfa[i + p1]);
            else
            {
              const float pwt = XMUL2(cfa[i] - rbp[j]) / (eps + rbp[j] + cfa[i]);
              rbp[j] = pwt * rbp[j] + (1.f - pwt) * ULIMF(rbp[j], cfa[i - p1], cfa[i + p1]);
            }
          }
          if(rbm[j] < cfa[i])
          {
            if(XMUL2(rbm[j]) < cfa[i])
              rbm[j] = ULIMF(rbm[j], cfa[i - m1], cfa[i + m1]);
            else
            {
              const float mwt = XMUL2(cfa[i] - rbm[j]) / (eps + rbm[j] + cfa[i]);
              rbm[j] = mwt * rbm[j] + (1.f - mwt) * ULIMF(rbm[j], cfa[i - m1], cfa[i + m1]);
            }
          }
          if(rbp[j] > clip_pt) rbp[j] = ULIMF(rbp[j], cfa[i - p1], cfa[i + p1]);
          if(rbm[j] > clip_pt) rbm[j] = ULIMF(rbm[j], cfa[i - m1], cfa[i + m1]);
        }

      /* ---- R+B from the two diagonals; pmwt is refined IN PLACE in raster order, :1106-1124 ---- */
      for(int rr = 10; rr < rr1 - 10; rr++)
        for(int cc = 10 + (orc_fc(rr, 2, filters) & 1), i = rr * ts + cc, j = i >> 1; cc < cc1 - 10; cc += 2, i += 2, j++)
        {
          const float pmwtalt = XDIV4(pmwt[(i - m1) >> 1] + pmwt[(i + p1) >> 1] + pmwt[(i - p1) >> 1] + pmwt[(i + m1) >> 1]);
          if(fabsf(0.5f - pmwt[j]) < fabsf(0.5f - pmwtalt)) pmwt[j] = pmwtalt;
          rbint[j] = XDIV2(cfa[i] + rbm[j] * (1.f - pmwt[j]) + rbp[j] * pmwt[j]);
        }

      /* ---- green re-interpolated where the diagonal direction discriminates better, :1127-1241 ---- */
      for(int rr = 12; rr < rr1 - 12; rr++)
        for(int cc = 12 + (orc_fc(rr, 2, filters) & 1), i = rr * ts + cc, j = i >> 1; cc < cc1 - 12; cc += 2, i += 2, j++)
        {
          if(fabsf(0.5f - pmwt[i >> 1]) < fabsf(0.5f - hvwt[i >> 1])) continue;
          const float cru = cfa[i - v1] * 2.0 / (eps + rbint[j] + rbint[(j - v1)]);
          const float crd = cfa[i + v1] * 2.0 / (eps + rbint[j] + rbint[(j + v1)]);
          const float crl = cfa[i - 1] * 2.0 / (eps + rbint[j] + rbint[(j - 1)]);
          const float crr = cfa[i + 1] * 2.0 / (eps + rbint[j] + rbint[(j + 1)]);
          const float gu = (fabsf(1.f - cru) < arthresh) ? rbint[j] * cru : cfa[i - v1] + XDIV2(rbint[j] - rbint[(j - v1)]);
          const float gd = (fabsf(1.f - crd) < arthresh) ? rbint[j] * crd : cfa[i + v1] + XDIV2(rbint[j] - rbint[(j + v1)]);
          const float gl = (fabsf(1.f - crl) < arthresh) ? rbint[j] * crl : cfa[i - 1] + XDIV2(rbint[j] - rbint[(j - 1)]);
          const float gr = (fabsf(1.f - crr) < arthresh) ? rbint[j] * crr : cfa[i + 1] + XDIV2(rbint[j] - rbint[(j + 1)]);
          float Gintv = (dirwts0[i - v1] * gd + dirwts0[i + v1] * gu) / (dirwts0[i + v1] + dirwts0[i - v1]);
          float Ginth = (dirwts1[i - 1] * gr + dirwts1[i + 1] * gl) / (dirwts1[i - 1] + dirwts1[i + 1]);
          if(Gintv < rbint[j])
          {
            if(2 * Gintv < rbint[j])
              Gintv = ULIMF(Gintv, cfa[i - v1], cfa[i + v1]);
            else
            {
              const float vwt = 2.0 * (rbint[j] - Gintv) / (eps + Gintv + rbint[j]);
              Gintv = vwt * Gintv + (1.f - vwt) * ULIMF(Gintv, cfa[i - v1], cfa[i + v1]);
            }
          }
          if(Ginth < rbint[j])
          {
            if(2 * Ginth < rbint[j])
              Ginth = ULIMF(Ginth, cfa[i - 1], cfa[i + 1]);
            else
            {
              const float hwt = 2.0 * (rbint[j] - Ginth) / (eps + Ginth + rbint[j]);
              Ginth = hwt * Ginth + (1.f - hwt) * ULIMF(Ginth, cfa[i - 1], cfa[i + 1]);
            }
          }
          if(Ginth > clip_pt) Ginth = ULIMF(Ginth, cfa[i - 1], cfa[i + 1]);
          if(Gintv > clip_pt) Gintv = ULIMF(Gintv, cfa[i - v1], cfa[i + v1]);
          rgbgreen[i] = Ginth * (1.f - hvwt[j]) + Gintv * hvwt[j];
          Dgrb0[i >> 1] = rgbgreen[i] - cfa[i];
        }

      /* ---- chroma: split G-B out of G-R, then interpolate each at the other colour's sites, :1247-1289 ---- */
      for(int rr = 13 - ey; rr < rr1 - 12; rr += 2)
        for(int j = (rr * ts + 13 - ex) >> 1; j < (rr * ts + cc1 - 12) >> 1; j++)
        {
          Dgrb1[j] = Dgrb0[j];
          Dgrb0[j] = 0;
        }
      for(int rr = 14; rr < rr1 - 14; rr++)
        for(int cc = 14 + (orc_fc(rr, 2, filters) & 1), i = rr * ts + cc, c = 1 - orc_fc(rr, cc, filters) / 2; cc < cc1 - 14; cc += 2, i += 2)
        {
          float *D = c ? Dgrb1 : Dgrb0;
          const float wtnw = 1.f / (eps + fabsf(D[(i - m1) >> 1] - D[(i + m1) >> 1]) + fabsf(D[(i - m1) >> 1] - D[(i - m3) >> 1])
                                    + fabsf(D[(i + m1) >> 1] - D[(i - m3) >> 1]));
          const float wtne = 1.f / (eps + fabsf(D[(i + p1) >> 1] - D[(i - p1) >> 1]) + fabsf(D[(i + p1) >> 1] - D[(i + p3) >> 1])
                                    + fabsf(D[(i - p1) >> 1] - D[(i + p3) >> 1]));
          const float wtsw = 1.f / (eps + fabsf(D[(i - p1) >> 1] - D[(i + p1) >> 1]) + fabsf(D[(i - p1) >> 1] - D[(i + m3) >> 1])
                                    + fabsf(D[(i + p1) >> 1] - D[(i - p3) >> 1]));
          const float wtse = 1.f / (eps + fabsf(D[(i + m1) >> 1] - D[(i - m1) >> 1]) + fabsf(D[(i + m1) >> 1] - D[(i - p3) >> 1])
                                    + fabsf(D[(i - m1) >> 1] - D[(i + m3) >> 1]));
          D[i >> 1] = (wtnw * (1.325f * D[(i - m1) >> 1] - 0.175f * D[(i - m3) >> 1] - 0.075f * D[(i - m1 - 2) >> 1] - 0.075f * D[(i - m1 - v2) >> 1])
                       + wtne * (1.325f * D[(i + p1) >> 1] - 0.175f * D[(i + p3) >> 1] - 0.075f * D[(i + p1 + 2) >> 1] - 0.075f * D[(i + p1 + v2) >> 1])
                       + wtsw * (1.325f * D[(i - p1) >> 1] - 0.175f * D[(i - p3) >> 1] - 0.075f * D[(i - p1 - 2) >> 1] - 0.075f * D[(i - p1 - v2) >> 1])
                       + wtse * (1.325f * D[(i + m1) >> 1] - 0.175f * D[(i + m3) >> 1] - 0.075f * D[(i + m1 + 2) >> 1] - 0.075f * D[(i + m1 + v2) >> 1]))
                      / (wtnw + wtne + wtsw + wtse);
        }

      /* ---- output: R and B at green sites from the four cardinal neighbours, at R/B sites directly; then G, :1291-1407 ---- */
#define GSITE(i_, D_)                                                                                                             \
  clampnan(rgbgreen[i_]                                                                                                            \
               - ((hvwt[((i_) - v1) >> 1]) * D_[((i_) - v1) >> 1] + (1.f - hvwt[((i_) + 1) >> 1]) * D_[((i_) + 1) >> 1]                   \
                  + (1.f - hvwt[((i_) - 1) >> 1]) * D_[((i_) - 1) >> 1] + (hvwt[((i_) + v1) >> 1]) * D_[((i_) + v1) >> 1])                \
                     * temp,                                                                                                       \
           0.0f, 1.0f)
      for(int rr = 16; rr < rr1 - 16; rr++)
      {
        const int row = rr + top;
        int col = left + 16;
        int i = rr * ts + 16;
        const int green_first = ((orc_fc(rr, 2, filters) & 1) == 1);
        /* the sites of a row alternate; `green_first` says whether the first one (tile column 16) is a green site */
        int at_green = green_first;
        const int end = rr * ts + cc1 - 16 - (cc1 & 1);
        for(; i < end; i++, col++)
        {
          for(int k = 0; k < 2; k++)
          {
            if(k)
            {
              i++;
              col++;
            }
            if(col < width && row < height)
            {
              float *o = out + ((size_t)row * width + col) * 4;
              if(at_green)
              {
                const float temp = 1.f / (hvwt[(i - v1) >> 1] + 2.f - hvwt[(i + 1) >> 1] - hvwt[(i - 1) >> 1] + hvwt[(i + v1) >> 1]);
                o[0] = GSITE(i, Dgrb0);
                o[2] = GSITE(i, Dgrb1);
              }
              else
              {
                o[0] = clampnan(rgbgreen[i] - Dgrb0[i >> 1], 0.0f, 1.0f);
                o[2] = clampnan(rgbgreen[i] - Dgrb1[i >> 1], 0.0f, 1.0f);
              }
            }
            at_green = !at_green;
          }
        }
        if(cc1 & 1)
        { /* odd tile width: one more site of the first kind */
          if(col < width && row < height)
          {
            float *o = out + ((size_t)row * width + col) * 4;
            if(green_first)
            {
              const float temp = 1.f / (hvwt[(i - v1) >> 1] + 2.f - hvwt[(i + 1) >> 1] - hvwt[(i - 1) >> 1] + hvwt[(i + v1) >> 1]);
              o[0] = GSITE(i, Dgrb0);
              o[2] = GSITE(i, Dgrb1);
            }
            else
            {
              o[0] = clampnan(rgbgreen[i] - Dgrb0[i >> 1], 0.0f, 1.0f);
              o[2] = clampnan(rgbgreen[i] - Dgrb1[i >> 1], 0.0f, 1.0f);
            }
          }
        }
      }
#undef GSITE
      for(int rr = 16; rr < rr1 - 16; rr++)
      {
        const int row = rr + top;
        for(int cc = 16; cc < cc1 - 16; cc++)
        {
          const int col = cc + left;
          if(col < width && row < height) out[((size_t)row * width + col) * 4 + 1] = clampnan(rgbgreen[rr * ts + cc], 0.0f, 1.0f);
        }
      }
    }
  free(buffer);
  return 0;
}
