/* CPU restatement of Frank Markesteijn's demosaicer for X-Trans sensors (1 and 3 passes).  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src/iop/demosaic/markesteijn.c xtrans_markesteijn_interpolate :47-521 (hexmap :30-40; FCxtrans:
 * develop/imageop_math.h:197-216).  Pinned bit-for-bit against those lines cut verbatim (oracle/_ref: ref_markesteijn.c).
 *
 * The reference walks the frame in tiles of 122x122 with a border of 12 (17 with three passes) on every side; the tile grid
 * is kept (a tile is the unit of work of the CUDA kernel too).  Within a tile every stage below is a function of the planes
 * the stages before it left, pixel by pixel -- except the first one, the bounds of green at the red/blue pairs (:199-246):
 * that loop hops between the two rows of a vertical pair by changing its own row counter, revisits pixels and lets the last
 * visit win.  It is restated as what it is, a walk: mk_walk() replays the loop's control flow (which depends on the tile's
 * position and size only) and records for every red/blue pixel which pixel started the run of the last visit; the values
 * follow from that record (mk_bounds).  The product builds the same record on the host and ships it to the kernel.
 */
#include "oracle_common.h"
#include <float.h>
#include <stdlib.h>
#include <string.h>

#define TS 122
#define NPX (TS * TS)

typedef struct
{
  int width, height, rx, ry, passes, ndir, pad;
  const uint8_t *xt; /* 6x6 */
  short hex[3][3][8];
  int sgrow, sgcol;
} mk_t;

static int mk_fc(const mk_t *m, int row, int col) { return m->xt[((row + 600 + m->ry) % 6) * 6 + (col + 600 + m->rx) % 6]; }
static const short *mk_hex(const mk_t *m, int row, int col) { return m->hex[(row + 600) % 3][(col + 600) % 3]; }
static int mk_mirror(int n, int size) { return n >= size ? 2 * size - n - 2 : abs(n); } /* TRANSLATE, :158 */

/* :52-103: the green hexagon around every non-green pixel (and the other way round), the position of the solitary greens */
static void mk_hexagons(mk_t *m)
{
  static const short orth[12] = { 1, 0, 0, 1, -1, 0, 0, -1, 1, 0, 0, 1 };
  static const short patt[2][16] = { { 0, 1, 0, -1, 2, 0, -1, 0, 1, 1, 1, -1, 0, 0, 0, 0 }, { 0, 1, 0, -2, 1, 0, -2, 0, 1, 1, -2, -2, 1, -1, -1, 1 } };
  m->sgrow = m->sgcol = 0;
  memset(m->hex, 0, sizeof(m->hex));
  for(int row = 0; row < 3; row++)
    for(int col = 0; col < 3; col++)
    {
      const int g = mk_fc(m, row, col) == 1;
      int ng = 0;
      for(int d = 0; d < 10; d += 2)
      {
        ng = (mk_fc(m, row + orth[d], col + orth[d + 2]) == 1) ? 0 : ng + 1;
        if(ng == 4)
        {
          m->sgrow = row;
          m->sgcol = col;
        }
        if(ng == g + 1)
          for(int c = 0; c < 8; c++)
          {
            const int v = orth[d] * patt[g][c * 2] + orth[d + 1] * patt[g][c * 2 + 1];
            const int h = orth[d + 2] * patt[g][c * 2] + orth[d + 3] * patt[g][c * 2 + 1];
            m->hex[row][col][c ^ (g * 2 & d)] = (short)(h + v * TS);
          }
      }
    }
}

/* The walk of :199-246 over one tile.  start[p] = tile index of the pixel whose hexagon opened the run that last wrote pixel p
 * (-1: never written).  A run is at most two pixels long: the second pixel of a pair keeps the bounds of the first -- unless the
 * first one's maximum came out as 0.0f, the loop's marker for "new pair", in which case it goes on with its own hexagon
 * (data dependent, resolved in mk_bounds).  Returns 0 if a run ever got longer than two pixels (it does not). */
int orc_markesteijn_walk(short *start, int top, int left, int mrow, int mcol, int sgrow, const uint8_t *xt, int rx, int ry)
{
  mk_t m = { 0 };
  m.xt = xt;
  m.rx = rx;
  m.ry = ry;
  for(int k = 0; k < NPX; k++) start[k] = -1;
  int ok = 1;
  for(int row = top + 3; row < mrow - 3; row++)
  {
    int open = -1, len = 0; /* the run: where it started, how many pixels it has written */
    for(int col = left + 3; col < mcol - 3; col++)
    {
      if(mk_fc(&m, row, col) == 1)
      {
        open = -1;
        len = 0;
        continue;
      }
      const int p = (row - top) * TS + (col - left);
      if(open < 0)
      {
        open = p;
        len = 0;
      }
      if(++len > 2) ok = 0;
      start[p] = (short)open;
      switch((row - sgrow) % 3)
      {
        case 1:
          if(row < mrow - 4) row++, col--;
          break;
        case 2:
          open = -1;
          len = 0;
          if((col += 2) < mcol - 4 && row > top + 3) row--;
      }
    }
  }
  return ok;
}

static void mk_minmax6(const float (*pix)[3], const short *hex, float *mn, float *mx)
{
  for(int c = 0; c < 6; c++)
  {
    const float v = pix[hex[c]][1];
    if(*mn > v) *mn = v;
    if(*mx < v) *mx = v;
  }
}
/* :203-231 for one pixel, from the record of the walk */
static void mk_bounds(const mk_t *m, const float (*rgb0)[3], const short *start, int top, int left, int p, float *gmin, float *gmax)
{
  float mn = FLT_MAX, mx = 0.0f;
  const int s = start[p];
  mk_minmax6(rgb0 + s, mk_hex(m, top + s / TS, left + s % TS), &mn, &mx);
  if(s != p && mx == 0.0f) mk_minmax6(rgb0 + p, mk_hex(m, top + p / TS, left + p % TS), &mn, &mx);
  *gmin = mn;
  *gmax = mx;
}

static float mk_clamps(float a, float l, float h) { return a > l ? (a < h ? a : h) : l; } /* CLAMPS, math/math.h:78 */
static float mk_sqr(float x) { return x * x; }

static void mk_tile(const mk_t *m, float *out, const float *in, int top, int left, float (*rgb)[NPX][3], float *gmin, float *gmax, float (*drv)[NPX],
                    uint8_t (*homo)[NPX], short *start)
{
  const int width = m->width, height = m->height, ndir = m->ndir, passes = m->passes;
  const int mrow = (top + TS < height + m->pad) ? top + TS : height + m->pad, mcol = (left + TS < width + m->pad) ? left + TS : width + m->pad;
  /* :139-186 the tile, mirrored beyond the frame; the same values in the first four planes */
  for(int row = top; row < mrow; row++)
    for(int col = left; col < mcol; col++)
    {
      float *pix = rgb[0][(row - top) * TS + (col - left)];
      const int f = mk_fc(m, row, col);
      pix[0] = pix[1] = pix[2] = 0.0f;
      if(col >= 0 && row >= 0 && col < width && row < height)
        pix[f] = in[(size_t)width * row + col];
      else
      {
        const int cy = mk_mirror(row, height), cx = mk_mirror(col, width);
        if(f == mk_fc(m, cy, cx))
          pix[f] = in[(size_t)width * cy + cx];
        else
        {
          float sum = 0.0f;
          uint8_t count = 0;
          for(int y = row - 1; y <= row + 1; y++)
            for(int x = col - 1; x <= col + 1; x++)
            {
              const int yy = mk_mirror(y, height), xx = mk_mirror(x, width);
              if(mk_fc(m, yy, xx) == f)
              {
                sum += in[(size_t)width * yy + xx];
                count++;
              }
            }
          pix[f] = sum / count;
        }
      }
    }
  for(int c = 1; c < 4; c++) memcpy(rgb[c], rgb[0], sizeof(rgb[0]));

  /* :199-246 bounds of green at the red/blue pairs */
  orc_markesteijn_walk(start, top, left, mrow, mcol, m->sgrow, m->xt, m->rx, m->ry);
  for(int p = 0; p < NPX; p++)
    if(start[p] >= 0) mk_bounds(m, rgb[0], start, top, left, p, gmin + p, gmax + p);

  /* :251-271 green along the four directions */
  for(int row = top + 3; row < mrow - 3; row++)
    for(int col = left + 3; col < mcol - 3; col++)
    {
      const int f = mk_fc(m, row, col);
      if(f == 1) continue;
      const int p = (row - top) * TS + (col - left);
      const float(*pix)[3] = rgb[0] + p;
      const short *hex = mk_hex(m, row, col);
      float color[4];
      color[0] = 0.6796875f * (pix[hex[1]][1] + pix[hex[0]][1]) - 0.1796875f * (pix[2 * hex[1]][1] + pix[2 * hex[0]][1]);
      color[1] = 0.87109375f * pix[hex[3]][1] + pix[hex[2]][1] * 0.13f + 0.359375f * (pix[0][f] - pix[-hex[2]][f]);
      for(int c = 0; c < 2; c++)
        color[2 + c] = 0.640625f * pix[hex[4 + c]][1] + 0.359375f * pix[-2 * hex[4 + c]][1]
                       + 0.12890625f * (2 * pix[0][f] - pix[3 * hex[4 + c]][f] - pix[-3 * hex[4 + c]][f]);
      const int flip = !((row - m->sgrow) % 3);
      for(int c = 0; c < 4; c++) rgb[c ^ flip][p][1] = mk_clamps(color[c], gmin[p], gmax[p]);
    }

  for(int pass = 0; pass < passes; pass++)
  {
    if(pass == 1)
    { /* :275-281 the second set of planes */
      memcpy(rgb + 4, rgb, sizeof(rgb[0]) * 4);
      rgb += 4;
    }
    if(pass)
    { /* :284-302 green again from the closer interpolated values, in place and in raster order */
      for(int row = top + 6; row < mrow - 6; row++)
        for(int col = left + 6; col < mcol - 6; col++)
        {
          const int f = mk_fc(m, row, col);
          if(f == 1) continue;
          const int p = (row - top) * TS + (col - left);
          const short *hex = mk_hex(m, row, col);
          for(int d = 3; d < 6; d++)
          {
            float(*rfx)[3] = rgb[(d - 2) ^ !((row - m->sgrow) % 3)] + p;
            const float val = rfx[-2 * hex[d]][1] + 2 * rfx[hex[d]][1] - rfx[-2 * hex[d]][f] - 2 * rfx[hex[d]][f] + 3 * rfx[0][f];
            rfx[0][1] = mk_clamps(val / 3.0f, gmin[p], gmax[p]);
          }
        }
    }
    /* :304-354 red and blue at the solitary greens */
    const int pad_sg = (passes == 1) ? 6 : 5;
    for(int row = (top - m->sgrow + pad_sg + 2) / 3 * 3 + m->sgrow; row < mrow - pad_sg; row += 3)
      for(int col = (left - m->sgcol + pad_sg + 2) / 3 * 3 + m->sgcol; col < mcol - pad_sg; col += 3)
      {
        float(*rfx)[3] = rgb[0] + (row - top) * TS + (col - left);
        int h = mk_fc(m, row, col + 1);
        float diff[6] = { 0.0f };
        float color[2][6];
        for(int i = 1, d = 0; d < 6; d++, i ^= TS ^ 1, h ^= 2)
        {
          for(int c = 0; c < 2; c++, h ^= 2)
          {
            const int o = i << c;
            const float g = 2 * rfx[0][1] - rfx[o][1] - rfx[-o][1];
            color[h != 0][d] = g + rfx[o][h] + rfx[-o][h];
            if(d > 1) diff[d] += mk_sqr(rfx[o][1] - rfx[-o][1] - rfx[o][h] + rfx[-o][h]) + mk_sqr(g);
          }
          if(d < 2 || (d & 1))
          {
            const int d_out = d - ((d > 1) && (diff[d - 1] < diff[d]));
            rfx[0][0] = color[0][d_out] / 2.f;
            rfx[0][2] = color[1][d_out] / 2.f;
            rfx += NPX;
          }
        }
      }
    /* :356-373 red at the blue pixels and blue at the red ones */
    const int pad_rb = (passes == 1) ? 6 : 5;
    for(int row = top + pad_rb; row < mrow - pad_rb; row++)
      for(int col = left + pad_rb; col < mcol - pad_rb; col++)
      {
        const int f = 2 - mk_fc(m, row, col);
        if(f == 1) continue;
        float(*rfx)[3] = rgb[0] + (row - top) * TS + (col - left);
        const int c = (row - m->sgrow) % 3 ? TS : 1;
        const int h = 3 * (c ^ TS ^ 1);
        for(int d = 0; d < 4; d++, rfx += NPX)
        {
          const int i = d > 1 || ((d ^ c) & 1)
                                || ((fabsf(rfx[0][1] - rfx[c][1]) + fabsf(rfx[0][1] - rfx[-c][1]))
                                    < 2.f * (fabsf(rfx[0][1] - rfx[h][1]) + fabsf(rfx[0][1] - rfx[-h][1])))
                            ? c
                            : h;
          rfx[0][f] = (rfx[i][f] + rfx[-i][f] + 2.f * rfx[0][1] - rfx[i][1] - rfx[-i][1]) / 2.f;
        }
      }
    /* :375-399 red and blue in the 2x2 blocks of green */
    const int pad_g22 = (passes == 1) ? 8 : 4;
    for(int row = top + pad_g22; row < mrow - pad_g22; row++)
    {
      if(!((row - m->sgrow) % 3)) continue;
      for(int col = left + pad_g22; col < mcol - pad_g22; col++)
      {
        if(!((col - m->sgcol) % 3)) continue;
        float(*rfx)[3] = rgb[0] + (row - top) * TS + (col - left);
        const short *hex = mk_hex(m, row, col);
        for(int d = 0; d < ndir; d += 2, rfx += NPX)
          if(hex[d] + hex[d + 1])
          {
            const float g = 3.f * rfx[0][1] - 2.f * rfx[hex[d]][1] - rfx[hex[d + 1]][1];
            for(int c = 0; c < 4; c += 2) rfx[0][c] = (g + 2.f * rfx[hex[d]][c] + rfx[hex[d + 1]][c]) / 3.f;
          }
          else
          {
            const float g = 2.f * rfx[0][1] - rfx[hex[d]][1] - rfx[hex[d + 1]][1];
            for(int c = 0; c < 4; c += 2) rfx[0][c] = (g + rfx[hex[d]][c] + rfx[hex[d + 1]][c]) / 2.f;
          }
      }
    }
  }
  if(passes > 1) rgb -= 4;

  /* :408-448 tile-local from here; luma/chroma differences along each direction */
  const int nrow = mrow - top, ncol = mcol - left;
  const int pad_yuv = (passes == 1) ? 8 : 13, pad_drv = pad_yuv + 1, pad_homo = pad_yuv + 2;
  static const int dir[4] = { 1, TS, TS + 1, TS - 1 };
  float(*yuv)[3] = malloc(sizeof(float) * 3 * NPX);
  for(int d = 0; d < ndir; d++)
  {
    for(int row = pad_yuv; row < nrow - pad_yuv; row++)
      for(int col = pad_yuv; col < ncol - pad_yuv; col++)
      {
        const float *rx = rgb[d][row * TS + col];
        const float y = 0.2627f * rx[0] + 0.6780f * rx[1] + 0.0593f * rx[2];
        float *t = yuv[row * TS + col];
        t[0] = y;
        t[1] = (rx[2] - y) * 0.56433f;
        t[2] = (rx[0] - y) * 0.67815f;
      }
    const int f = dir[d & 3];
    for(int row = pad_drv; row < nrow - pad_drv; row++)
      for(int col = pad_drv; col < ncol - pad_drv; col++)
      {
        const float(*t)[3] = (const float(*)[3])yuv + row * TS + col;
        drv[d][row * TS + col] = mk_sqr(2 * t[0][0] - t[f][0] - t[-f][0]) + mk_sqr(2 * t[0][1] - t[f][1] - t[-f][1]) + mk_sqr(2 * t[0][2] - t[f][2] - t[-f][2]);
      }
  }
  free(yuv);

  /* :450-464 homogeneity: how many of the 3x3 neighbours are no steeper than 8 times the flattest direction of the centre */
  memset(homo, 0, sizeof(uint8_t) * ndir * NPX);
  for(int row = pad_homo; row < nrow - pad_homo; row++)
    for(int col = pad_homo; col < ncol - pad_homo; col++)
    {
      float tr = FLT_MAX;
      for(int d = 0; d < ndir; d++)
        if(tr > drv[d][row * TS + col]) tr = drv[d][row * TS + col];
      tr *= 8;
      for(int d = 0; d < ndir; d++)
        for(int v = -1; v <= 1; v++)
          for(int h = -1; h <= 1; h++) homo[d][row * TS + col] += (drv[d][(row + v) * TS + col + h] <= tr) ? 1 : 0;
    }

  /* :466-515 5x5 sums of the maps (the reference rolls them along the row in uint8 arithmetic: the same sums), the average of the
   * most homogeneous directions */
  for(int row = m->pad; row < nrow - m->pad; row++)
    for(int col = m->pad; col < ncol - m->pad; col++)
    {
      uint8_t hm[8] = { 0 };
      uint8_t maxval = 0;
      for(int d = 0; d < ndir; d++)
      {
        unsigned s = 0;
        for(int v = -2; v <= 2; v++)
          for(int h = -2; h <= 2; h++) s += homo[d][(row + v) * TS + col + h];
        hm[d] = (uint8_t)s;
        maxval = maxval < hm[d] ? hm[d] : maxval;
      }
      maxval -= maxval >> 3;
      for(int d = 0; d < ndir - 4; d++)
      {
        if(hm[d] < hm[d + 4])
          hm[d] = 0;
        else if(hm[d] > hm[d + 4])
          hm[d + 4] = 0;
      }
      float avg[4] = { 0.0f };
      for(int d = 0; d < ndir; d++)
        if(hm[d] >= maxval)
        {
          for(int c = 0; c < 3; c++) avg[c] += rgb[d][row * TS + col][c];
          avg[3]++;
        }
      for(int c = 0; c < 3; c++) out[4 * ((size_t)width * (row + top) + col + left) + c] = avg[c] / avg[3];
    }
}

/* xtrans_markesteijn_interpolate(): in = the mosaic of the region, (x, y) = its origin on the sensor; lane 3 of out is left alone */
void orc_markesteijn(float *out, const float *in, int width, int height, int x, int y, const uint8_t xtrans[36], int passes)
{
  mk_t m = { 0 };
  m.width = width;
  m.height = height;
  m.rx = x;
  m.ry = y;
  m.xt = xtrans;
  m.passes = passes;
  m.ndir = 4 << (passes > 1);
  m.pad = (passes == 1) ? 12 : 17;
  mk_hexagons(&m);
  orc_fp_fast_mode(); /* the pipe's threads run with FTZ|DAZ (darktable.c:877, common/dtpthread.c:54) */
  float(*rgb)[NPX][3] = calloc(8, sizeof(rgb[0]));
  float *gmin = calloc(NPX, sizeof(float)), *gmax = calloc(NPX, sizeof(float));
  float(*drv)[NPX] = calloc(8, sizeof(drv[0]));
  uint8_t(*homo)[NPX] = calloc(8, sizeof(homo[0]));
  short *start = malloc(sizeof(short) * NPX);
  for(int top = -m.pad; top < height - m.pad; top += TS - 2 * m.pad)
    for(int left = -m.pad; left < width - m.pad; left += TS - 2 * m.pad) mk_tile(&m, out, in, top, left, rgb, gmin, gmax, drv, homo, start);
  free(rgb);
  free(gmin);
  free(gmax);
  free(drv);
  free(homo);
  free(start);
}
void orc_markesteijn_hexagons(short hex[72], int *sgrow, int *sgcol, int x, int y, const uint8_t xtrans[36])
{
  mk_t m = { 0 };
  m.rx = x;
  m.ry = y;
  m.xt = xtrans;
  mk_hexagons(&m);
  memcpy(hex, m.hex, sizeof(m.hex));
  *sgrow = m.sgrow;
  *sgcol = m.sgcol;
}
