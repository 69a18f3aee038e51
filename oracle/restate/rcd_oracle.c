/* CPU restatement of Ratio-Corrected Demosaicing as the reference runs it.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src/iop/demosaic/rcd.c:
 *   frame-edge ring   rcd_ppg_border()  rcd.c:91-272
 *   tile walk         rcd_demosaic()    rcd.c:274-564   (112x112 tiles, 94x94 kept, border 9 / margin 6)
 *
 * What is restated exactly and why it matters for parity:
 *  - The reference output depends on its 112/94 tile grid (values at tile-local row/col 4 see a
 *    zero VH_Dir ring, rcd.c:302-304,334-338), so the same grid is walked here.
 *  - Half-width planes (lpf/PQ_Dir, P/Q_CDiff_Hpf) are addressed with `flat_index / 2` exactly
 *    as rcd.c:396,444,453,464 do; lpf and PQ_Dir share storage (rcd.c:314).
 *  - `fabs()` on float operands is the double function: the gradient sums of steps 3.1, 4.2, 4.3
 *    are accumulated in double and rounded once to float (rcd.c:413-416,473-476,503-522).
 *  - Scratch the reference never writes inside a tile (it inherits malloc contents or the
 *    previous tile's data, rcd.c:302-311,339) is set to `scratch_fill` at the start of every
 *    tile.  scratch_fill = 0 is the defined behaviour; scratch_fill = NaN marks every output
 *    pixel that is a function of such memory (see orc_rcd_undefined_mask).
 */
#include "oracle_common.h"
#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__) || defined(__i386__)
#include <xmmintrin.h>
#endif

#define T 112 /* RCD_TILESIZE rcd.c:53-55 */
#define KEEP (T - 2 * 9) /* RCD_TILEVALID rcd.c:75 */
#define RING 9 /* RCD_BORDER rcd.c:73 */
#define EDGE 6 /* RCD_MARGIN rcd.c:74 */

static const float k_eps = 1e-5f;    /* rcd.c:81 */
static const float k_epssq = 1e-10f; /* rcd.c:82 */

void orc_fp_fast_mode(void)
{
#if defined(__x86_64__) || defined(__i386__)
  _mm_setcsr(_mm_getcsr() | 0x8040u); /* FTZ | DAZ */
#endif
}

/* the same on every thread of the OpenMP pool (shared with the reference build loaded into this process): the
 * reference's pipe threads all run with FTZ|DAZ (darktable.c:877, common/dtpthread.c:54) */
void orc_fp_fast_mode_all(void)
{
  orc_fp_fast_mode();
#pragma omp parallel
  orc_fp_fast_mode();
}

static inline float pos(float v) { return fmaxf(0.0f, v); }
/* fmaxf()/compare that let a NaN through: used when the never-written scratch is filled with NaN
 * so that every value the reference derives from uninitialised memory stays marked */
static inline float vmax(float a, float b, int taint)
{
  if(taint && (a != a || b != b)) return NAN;
  return fmaxf(a, b);
}
static inline float sq(float v) { return v * v; }
/* iop/demosaic.c:250-257 */
static inline float mix(float a, float b, float c) { return a * (b - c) + c; }

/* 7-tap colour-difference high-pass along a stride, squared: rcd.c:360,376,380,446,447 */
static inline float hpf2(const float *p, ptrdiff_t s)
{
  return sq((p[-3 * s] - p[-s] - p[s] + p[3 * s]) - 3.0f * (p[-2 * s] + p[2 * s]) + 6.0f * p[0]);
}

/* pick the neighbourhood value when it is the more decisive one: rcd.c:433,470,501 */
static inline float refine(float centre, float nb, int taint)
{
  if(taint && (centre != centre || nb != nb)) return NAN;
  return (fabsf(0.5f - centre) < fabsf(0.5f - nb)) ? nb : centre;
}

typedef struct
{
  float cfa[T * T];
  float vh[T * T];
  float rgb[3][T * T];
  float pq[T * T / 2]; /* low-pass first, then PQ_Dir: same storage, rcd.c:314 */
  float pd[T * T / 2]; /* P_CDiff_Hpf */
  float qd[T * T / 2]; /* Q_CDiff_Hpf */
} rcd_tile_t;

static void fill(float *p, size_t n, float v)
{
  for(size_t k = 0; k < n; k++) p[k] = v;
}

static void rcd_one_tile(rcd_tile_t *t, float *out, const float *in, int width, int height, uint32_t filters,
                         int tv, int th, int nv, int nh, float scaler, float revscaler, float scratch_fill)
{
  const int row0 = tv * KEEP, col0 = th * KEEP;
  const int row1 = (row0 + T < height) ? row0 + T : height;
  const int col1 = (col0 + T < width) ? col0 + T : width;
  const int tr = row1 - row0, tc = col1 - col0;
  const int taint = scratch_fill != scratch_fill;

  /* state the reference leaves to chance; VH_Dir's ring is genuinely zero (rcd.c:304,338) */
  memset(t->vh, 0, sizeof(t->vh));
  fill(t->cfa, T * T, scratch_fill);
  fill(t->rgb[0], 3 * T * T, scratch_fill);
  fill(t->pq, T * T / 2, scratch_fill);
  fill(t->pd, T * T / 2, scratch_fill);
  fill(t->qd, T * T / 2, scratch_fill);
  if(row0 + T > height || col0 + T > width) memset(t->rgb, 0, sizeof(t->rgb)); /* rcd.c:334-341 */

  float *const cfa = t->cfa, *const vh = t->vh, *const lp = t->pq, *const pqd = t->pq;
  float *const G = t->rgb[1];

  /* step 0: rcd.c:343-351 */
  for(int r = 0; r < tr; r++)
  {
    const int ca = orc_fc(row0 + r, col0, filters), cb = orc_fc(row0 + r, col0 + 1, filters);
    for(int c = 0; c < tc; c++)
    {
      const float v = pos(in[(size_t)(row0 + r) * width + col0 + c]) * revscaler;
      const int i = r * T + c;
      cfa[i] = v;
      t->rgb[ca][i] = v;
      t->rgb[cb][i] = v;
    }
  }

  /* step 1: V/H direction strength, rcd.c:353-390.  The reference rolls three line buffers;
   * the same numbers are V(row) = hpf2 vertical at `row`, H(col) = hpf2 horizontal at `col`. */
  for(int r = 4; r < tr - 4; r++)
    for(int c = 4; c < tc - 4; c++)
    {
      const int i = r * T + c;
      const float vs = vmax(k_epssq, hpf2(cfa + i - T, T) + hpf2(cfa + i, T) + hpf2(cfa + i + T, T), taint);
      const float hs = vmax(k_epssq, hpf2(cfa + i - 1, 1) + hpf2(cfa + i, 1) + hpf2(cfa + i + 1, 1), taint);
      vh[i] = vs / (vs + hs);
    }

  /* step 2.1: low-pass at the non-green sites, rcd.c:394-402 */
  for(int r = 2; r < tr - 2; r++)
    for(int c = 2 + (orc_fc(r, 0, filters) & 1); c < tc - 2; c += 2)
    {
      const int i = r * T + c;
      lp[i / 2] = cfa[i] + 0.5f * (cfa[i - T] + cfa[i + T] + cfa[i - 1] + cfa[i + 1])
                  + 0.25f * (cfa[i - T - 1] + cfa[i - T + 1] + cfa[i + T - 1] + cfa[i + T + 1]);
    }

  /* step 3.1: green at red/blue sites, rcd.c:406-437 */
  for(int r = 4; r < tr - 4; r++)
    for(int c = 4 + (orc_fc(r, 0, filters) & 1); c < tc - 4; c += 2)
    {
      const int i = r * T + c, h = i / 2;
      const float x = cfa[i];
      /* double accumulation: fabs() is the double function */
      const float gn = (float)((double)k_eps + fabs((double)(cfa[i - T] - cfa[i + T])) + fabs((double)(x - cfa[i - 2 * T]))
                               + fabs((double)(cfa[i - T] - cfa[i - 3 * T])) + fabs((double)(cfa[i - 2 * T] - cfa[i - 4 * T])));
      const float gs = (float)((double)k_eps + fabs((double)(cfa[i - T] - cfa[i + T])) + fabs((double)(x - cfa[i + 2 * T]))
                               + fabs((double)(cfa[i + T] - cfa[i + 3 * T])) + fabs((double)(cfa[i + 2 * T] - cfa[i + 4 * T])));
      const float gw = (float)((double)k_eps + fabs((double)(cfa[i - 1] - cfa[i + 1])) + fabs((double)(x - cfa[i - 2]))
                               + fabs((double)(cfa[i - 1] - cfa[i - 3])) + fabs((double)(cfa[i - 2] - cfa[i - 4])));
      const float ge = (float)((double)k_eps + fabs((double)(cfa[i - 1] - cfa[i + 1])) + fabs((double)(x - cfa[i + 2]))
                               + fabs((double)(cfa[i + 1] - cfa[i + 3])) + fabs((double)(cfa[i + 2] - cfa[i + 4])));

      const float l = lp[h];
      const float en = cfa[i - T] * (l + l) / (k_eps + l + lp[h - T]);
      const float es = cfa[i + T] * (l + l) / (k_eps + l + lp[h + T]);
      const float ew = cfa[i - 1] * (l + l) / (k_eps + l + lp[h - 1]);
      const float ee = cfa[i + 1] * (l + l) / (k_eps + l + lp[h + 1]);

      const float ev = (gs * en + gn * es) / (gn + gs);
      const float eh = (gw * ee + ge * ew) / (ge + gw);

      const float nb = 0.25f * (vh[i - T - 1] + vh[i - T + 1] + vh[i + T - 1] + vh[i + T + 1]);
      G[i] = mix(refine(vh[i], nb, taint), eh, ev);
    }

  /* step 4.0: diagonal high-pass, every second column from 3 whatever the site colour, rcd.c:442-449 */
  for(int r = 3; r < tr - 3; r++)
    for(int c = 3; c < tc - 3; c += 2)
    {
      const int i = r * T + c;
      t->pd[i / 2] = hpf2(cfa + i, T + 1);
      t->qd[i / 2] = hpf2(cfa + i, T - 1);
    }

  /* step 4.1: P/Q direction strength; overwrites the low-pass where it lands, rcd.c:451-459 */
  for(int r = 4; r < tr - 4; r++)
    for(int c = 4 + (orc_fc(r, 0, filters) & 1); c < tc - 4; c += 2)
    {
      const int i = r * T + c, h = i / 2, hu = (i - T - 1) / 2, hd = (i + T - 1) / 2;
      const float ps = vmax(k_epssq, t->pd[hu] + t->pd[h] + t->pd[hd + 1], taint);
      const float qs = vmax(k_epssq, t->qd[hu + 1] + t->qd[h] + t->qd[hd], taint);
      pqd[h] = ps / (ps + qs);
    }

  /* step 4.2: the opposite colour at red/blue sites, rcd.c:462-491 */
  for(int r = 4; r < tr - 4; r++)
    for(int c = 4 + (orc_fc(r, 0, filters) & 1); c < tc - 4; c += 2)
    {
      const int i = r * T + c, h = i / 2, hu = (i - T - 1) / 2, hd = (i + T - 1) / 2;
      float *const C = t->rgb[2 - orc_fc(r, c, filters)];
      const float nb = 0.25f * (pqd[hu] + pqd[hu + 1] + pqd[hd] + pqd[hd + 1]);
      const float disc = refine(pqd[h], nb, taint);

      const int nw = i - T - 1, ne = i - T + 1, sw = i + T - 1, se = i + T + 1;
      const float gnw = (float)((double)k_eps + fabs((double)(C[nw] - C[se])) + fabs((double)(C[nw] - C[i - 3 * T - 3]))
                                + fabs((double)(G[i] - G[i - 2 * T - 2])));
      const float gne = (float)((double)k_eps + fabs((double)(C[ne] - C[sw])) + fabs((double)(C[ne] - C[i - 3 * T + 3]))
                                + fabs((double)(G[i] - G[i - 2 * T + 2])));
      const float gsw = (float)((double)k_eps + fabs((double)(C[ne] - C[sw])) + fabs((double)(C[sw] - C[i + 3 * T - 3]))
                                + fabs((double)(G[i] - G[i + 2 * T - 2])));
      const float gse = (float)((double)k_eps + fabs((double)(C[nw] - C[se])) + fabs((double)(C[se] - C[i + 3 * T + 3]))
                                + fabs((double)(G[i] - G[i + 2 * T + 2])));

      const float dnw = C[nw] - G[nw], dne = C[ne] - G[ne], dsw = C[sw] - G[sw], dse = C[se] - G[se];
      const float ep = (gnw * dse + gse * dnw) / (gnw + gse);
      const float eq = (gne * dsw + gsw * dne) / (gne + gsw);
      C[i] = G[i] + mix(disc, eq, ep);
    }

  /* step 4.3: red and blue at green sites, rcd.c:494-538 */
  for(int r = 4; r < tr - 4; r++)
    for(int c = 4 + (orc_fc(r, 1, filters) & 1); c < tc - 4; c += 2)
    {
      const int i = r * T + c;
      const float nb = 0.25f * (vh[i - T - 1] + vh[i - T + 1] + vh[i + T - 1] + vh[i + T + 1]);
      const float disc = refine(vh[i], nb, taint);
      const float g = G[i];
      /* each of these is rounded to float on its own (const float N1 = eps + fabs(..)) */
      const float n1 = (float)((double)k_eps + fabs((double)(g - G[i - 2 * T])));
      const float s1 = (float)((double)k_eps + fabs((double)(g - G[i + 2 * T])));
      const float w1 = (float)((double)k_eps + fabs((double)(g - G[i - 2])));
      const float e1 = (float)((double)k_eps + fabs((double)(g - G[i + 2])));
      const float gu = G[i - T], gd = G[i + T], gl = G[i - 1], gr = G[i + 1];

      for(int k = 0; k <= 2; k += 2)
      {
        float *const C = t->rgb[k];
        const float sn = fabsf(C[i - T] - C[i + T]);
        const float ew = fabsf(C[i - 1] - C[i + 1]);
        /* (float + float) first, then the double fabs() term */
        const float gn = (float)((double)(n1 + sn) + fabs((double)(C[i - T] - C[i - 3 * T])));
        const float gs = (float)((double)(s1 + sn) + fabs((double)(C[i + T] - C[i + 3 * T])));
        const float gw = (float)((double)(w1 + ew) + fabs((double)(C[i - 1] - C[i - 3])));
        const float ge = (float)((double)(e1 + ew) + fabs((double)(C[i + 1] - C[i + 3])));

        const float dn = C[i - T] - gu, ds = C[i + T] - gd, dw = C[i - 1] - gl, de = C[i + 1] - gr;
        const float ev = (gn * ds + gs * dn) / (gn + gs);
        const float eh = (ge * dw + gw * de) / (ge + gw);
        C[i] = g + mix(disc, eh, ev);
      }
    }

  /* keep the tile interior: rcd.c:541-554 */
  const int ra = row0 + (tv == 0 ? EDGE : RING), rb = row1 - (tv == nv - 1 ? EDGE : RING);
  const int ca = col0 + (th == 0 ? EDGE : RING), cb = col1 - (th == nh - 1 ? EDGE : RING);
  for(int r = ra; r < rb; r++)
    for(int c = ca; c < cb; c++)
    {
      const int i = (r - row0) * T + (c - col0);
      float *o = out + 4 * ((size_t)r * width + c);
      o[0] = scaler * vmax(0.0f, t->rgb[0][i], taint);
      o[1] = scaler * vmax(0.0f, t->rgb[1][i], taint);
      o[2] = scaler * vmax(0.0f, t->rgb[2][i], taint);
      o[3] = 0.0f;
    }
}

/* ---- frame-edge ring: rcd_ppg_border(), rcd.c:91-272, restated per pixel -------------------
 * The reference makes three in-place sweeps over `out`.  Read back per pixel at frame distance
 * d = min(row, col, height-1-row, width-1-col) they amount to:
 *   d < 3          : every non-native channel = mean of that colour over the 3x3 neighbourhood
 *                    clipped to the frame (rcd.c:98-125); alpha is NOT written (left as found)
 *   3 <= d < 9     : green at red/blue sites from the PPG gradient test (rcd.c:130-191), alpha 0
 *   1 <= d < 6     : red/blue from neighbours' native value and green (rcd.c:196-270)
 * Only d < 6 survives in the final image; the tiles overwrite the rest (rcd.c:541-554).
 */
typedef struct
{
  const float *in;
  int w, h;
  uint32_t f;
} ring_t;

static inline int ring_dist(const ring_t *q, int r, int c)
{
  int d = r < c ? r : c;
  if(q->h - 1 - r < d) d = q->h - 1 - r;
  if(q->w - 1 - c < d) d = q->w - 1 - c;
  return d;
}
static inline float ring_raw(const ring_t *q, int r, int c) { return pos(q->in[(size_t)r * q->w + c]); }

/* first sweep value of channel k at (r,c): rcd.c:98-125 */
static float ring_mean(const ring_t *q, int r, int c, int k)
{
  const int f = orc_fc(r, c, q->f);
  if(k == f) return ring_raw(q, r, c);
  float sum = 0.0f, cnt = 0.0f;
  for(int y = r - 1; y != r + 2; y++)
    for(int x = c - 1; x != c + 2; x++)
      if(y >= 0 && x >= 0 && y < q->h && x < q->w && orc_fc(y, x, q->f) == k)
      {
        sum += ring_raw(q, y, x);
        cnt++;
      }
  return cnt > 0.0f ? sum / cnt : ring_raw(q, r, c);
}

/* green after the first two sweeps */
static float ring_green(const ring_t *q, int r, int c)
{
  if(ring_dist(q, r, c) < 3) return ring_mean(q, r, c, 1);
  const int f = orc_fc(r, c, q->f);
  const float pc = ring_raw(q, r, c);
  if(!(f == 0 || f == 2)) return pc;
  const float ym = ring_raw(q, r - 1, c), ym2 = ring_raw(q, r - 2, c), ym3 = ring_raw(q, r - 3, c);
  const float yp = ring_raw(q, r + 1, c), yp2 = ring_raw(q, r + 2, c), yp3 = ring_raw(q, r + 3, c);
  const float xm = ring_raw(q, r, c - 1), xm2 = ring_raw(q, r, c - 2), xm3 = ring_raw(q, r, c - 3);
  const float xp = ring_raw(q, r, c + 1), xp2 = ring_raw(q, r, c + 2), xp3 = ring_raw(q, r, c + 3);
  const float guessx = (xm + pc + xp) * 2.0f - xp2 - xm2;
  const float diffx = (fabsf(xm2 - pc) + fabsf(xp2 - pc) + fabsf(xm - xp)) * 3.0f + (fabsf(xp3 - xp) + fabsf(xm3 - xm)) * 2.0f;
  const float guessy = (ym + pc + yp) * 2.0f - yp2 - ym2;
  const float diffy = (fabsf(ym2 - pc) + fabsf(yp2 - pc) + fabsf(ym - yp)) * 3.0f + (fabsf(yp3 - yp) + fabsf(ym3 - ym)) * 2.0f;
  if(diffx > diffy) return fmaxf(fminf(guessy * .25f, fmaxf(ym, yp)), fminf(ym, yp));
  return fmaxf(fminf(guessx * .25f, fmaxf(xm, xp)), fminf(xm, xp));
}

/* channel k of neighbour (r,c) as the third sweep finds it in `out` (k is that site's own colour
 * for every standard Bayer phase; for d < 3 the first sweep's value is what is stored) */
static float ring_chan(const ring_t *q, int r, int c, int k)
{
  if(ring_dist(q, r, c) < 3) return ring_mean(q, r, c, k);
  return ring_raw(q, r, c);
}

static void ring_pixel(const ring_t *q, int r, int c, float px[3])
{
  const int d = ring_dist(q, r, c);
  const int f = orc_fc(r, c, q->f);
  if(d < 3)
    for(int k = 0; k < 3; k++) px[k] = ring_mean(q, r, c, k);
  else
  {
    px[0] = px[2] = 0.0f;
    if(f == 0 || f == 2) px[f] = ring_raw(q, r, c);
  }
  px[1] = ring_green(q, r, c);
  if(d < 1) return;

  const float g = px[1];
  if(f & 1)
  {
    const float gt = ring_green(q, r - 1, c), gb = ring_green(q, r + 1, c);
    const float gl = ring_green(q, r, c - 1), gr = ring_green(q, r, c + 1);
    if(orc_fc(r, c + 1, q->f) == 0)
    { /* red beside, blue above/below: rcd.c:219-223 */
      px[2] = (ring_chan(q, r - 1, c, 2) + ring_chan(q, r + 1, c, 2) + 2.0f * g - gt - gb) * .5f;
      px[0] = (ring_chan(q, r, c - 1, 0) + ring_chan(q, r, c + 1, 0) + 2.0f * g - gl - gr) * .5f;
    }
    else
    { /* rcd.c:224-229 */
      px[0] = (ring_chan(q, r - 1, c, 0) + ring_chan(q, r + 1, c, 0) + 2.0f * g - gt - gb) * .5f;
      px[2] = (ring_chan(q, r, c - 1, 2) + ring_chan(q, r, c + 1, 2) + 2.0f * g - gl - gr) * .5f;
    }
  }
  else
  { /* rcd.c:231-266 */
    const int k = (f == 0) ? 2 : 0;
    const float tl = ring_chan(q, r - 1, c - 1, k), tr = ring_chan(q, r - 1, c + 1, k);
    const float bl = ring_chan(q, r + 1, c - 1, k), br = ring_chan(q, r + 1, c + 1, k);
    const float gtl = ring_green(q, r - 1, c - 1), gtr = ring_green(q, r - 1, c + 1);
    const float gbl = ring_green(q, r + 1, c - 1), gbr = ring_green(q, r + 1, c + 1);
    const float diff1 = fabsf(tl - br) + fabsf(gtl - g) + fabsf(gbr - g);
    const float guess1 = tl + br + 2.0f * g - gtl - gbr;
    const float diff2 = fabsf(tr - bl) + fabsf(gtr - g) + fabsf(gbl - g);
    const float guess2 = tr + bl + 2.0f * g - gtr - gbl;
    if(diff1 > diff2)
      px[k] = guess2 * .5f;
    else if(diff1 < diff2)
      px[k] = guess1 * .5f;
    else
      px[k] = (guess1 + guess2) * .25f;
  }
}

static void rcd_ring(float *out, const float *in, int width, int height, uint32_t filters)
{
  const ring_t q = { in, width, height, filters };
#pragma omp parallel for schedule(static)
  for(int r = 0; r < height; r++)
    for(int c = 0; c < width; c++)
    {
      if(c == EDGE && r >= EDGE && r < height - EDGE) c = width - EDGE;
      float px[3];
      ring_pixel(&q, r, c, px);
      float *o = out + 4 * ((size_t)r * width + c);
      o[0] = px[0];
      o[1] = px[1];
      o[2] = px[2];
      o[3] = 0.0f; /* the reference leaves alpha of the outer 3 px as it found it */
    }
}

/* Full-frame RCD.  `filters` already carries the ROI phase (iop/demosaic.c:1071).
 * Returns 0, or 1 for the reference's "too small area" early-out (rcd.c:280-284; output untouched). */
int orc_rcd_demosaic(float *out, const float *in, int width, int height, uint32_t filters,
                     const float processed_maximum[3], float scratch_fill)
{
  if(width < 16 || height < 16) return 1;
  rcd_ring(out, in, width, height, filters);

  const float scaler = fmaxf(processed_maximum[0], fmaxf(processed_maximum[1], processed_maximum[2]));
  const float revscaler = 1.0f / scaler;
  const int nv = 1 + (height - 2 * RING - 1) / KEEP, nh = 1 + (width - 2 * RING - 1) / KEEP;

#pragma omp parallel
  {
    orc_fp_fast_mode(); /* rcd.c:300 */
    rcd_tile_t *t = aligned_alloc(64, ((sizeof(rcd_tile_t) + 63) / 64) * 64);
#pragma omp for schedule(dynamic, 4) collapse(2)
    for(int tv = 0; tv < nv; tv++)
      for(int th = 0; th < nh; th++)
        rcd_one_tile(t, out, in, width, height, filters, tv, th, nv, nh, scaler, revscaler, scratch_fill);
    free(t);
  }
  return 0;
}

/* mask[r*width+c]: bit 0 = some colour channel of the reference's result is a function of memory
 * the reference never initialised (found by filling that memory with NaN and letting it
 * propagate: a conservative superset of the pixels that actually move); bit 1 = the pixel lies
 * in the outer 3-px ring whose alpha lane the reference never writes (rcd.c:117-123). */
int orc_rcd_undefined_mask(uint8_t *mask, const float *in, int width, int height, uint32_t filters,
                           const float processed_maximum[3])
{
  const size_t n = (size_t)width * height;
  float *a = malloc(n * 16);
  if(!a) return 2;
  const int rc = orc_rcd_demosaic(a, in, width, height, filters, processed_maximum, NAN);
  if(!rc)
    for(size_t k = 0; k < n; k++)
    {
      const int r = (int)(k / width), c = (int)(k % width);
      int d = r < c ? r : c;
      if(height - 1 - r < d) d = height - 1 - r;
      if(width - 1 - c < d) d = width - 1 - c;
      const float *p = a + 4 * k;
      mask[k] = (uint8_t)((p[0] != p[0] || p[1] != p[1] || p[2] != p[2]) | ((d < 3) << 1));
    }
  free(a);
  return rc;
}
