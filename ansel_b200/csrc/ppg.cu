// PPG demosaic (the reference's fallback Bayer method) and its optional pre-median.
//
// Reference: iop/demosaic/ppg.c demosaic_ppg :21-211; iop/demosaic/basic.c pre_median_b :136-180; dispatch
// iop/demosaic.c:1218-1226.
//
// The reference makes three passes over the frame: a border average into the outer three pixels, the green plane of the
// interior, then red and blue in place from the greens around each pixel.  The in-place pass only ever reads values it
// does not write (a site's own colour and the greens), so every output pixel is a pure function of the mosaic within
// +-4 pixels.  The kernel computes it that way: one thread per pixel re-derives the (at most five) greens it needs from
// the mosaic, which L1/L2 serve, and stores the pixel once -- 20 bytes per pixel of HBM traffic (4 in, 16 out) instead
// of the 52 the three passes would move.  Arithmetic and its order are the reference's, so results are bit-identical.
#ifndef B200_KERNELS_ON_CPU // tests/emul compiles the kernels of this file with g++ to check them against the oracle without a GPU
#include "runtime.h"
#endif
#include <math.h>

namespace
{
constexpr int PNT = 256;

struct ppg_frame_t
{
  const float *in;    // the mosaic as the module received it (border averages read this)
  const float *input; // the pre-median'ed mosaic when the threshold is > 0, else == in (interior reads this)
  int width, height;
  unsigned filters; // ROI-shifted
};

__device__ __forceinline__ int ppg_fc(int row, int col, unsigned filters) { return (filters >> ((((row << 1) & 14) + (col & 1)) << 1)) & 3; }
__device__ __forceinline__ bool in_ring(const ppg_frame_t &F, int y, int x) { return y < 3 || y >= F.height - 3 || x < 3 || x >= F.width - 3; }

// channel c of a pixel of the three-pixel border as ppg.c:31-56 leaves it: the average of the 3x3 neighbours of that
// colour that lie inside the frame (raster order), or the site's own sample for its own colour / when there is none
__device__ float border_value(const ppg_frame_t &F, int j, int i, int c)
{
  const float own = F.in[(size_t)j * F.width + i];
  if(c == ppg_fc(j, i, F.filters)) return own;
  float sum = 0.0f, count = 0.0f;
  for(int y = j - 1; y != j + 2; y++)
    for(int x = i - 1; x != i + 2; x++)
      if(y >= 0 && x >= 0 && y < F.height && x < F.width && ppg_fc(y, x, F.filters) == c)
      {
        sum += F.in[(size_t)y * F.width + x];
        count += 1.0f;
      }
  return count > 0.0f ? sum / count : own;
}

// the green of pixel (y, x) after the first two passes: :70-131 in the interior, the border average elsewhere
__device__ float green_at(const ppg_frame_t &F, int y, int x)
{
  if(in_ring(F, y, x)) return border_value(F, y, x, 1);
  const int w = F.width;
  const float *b = F.input + (size_t)y * w + x;
  const float pc = b[0];
  const int c = ppg_fc(y, x, F.filters);
  if(c != 0 && c != 2) return pc;
  const float pym = b[-w], pym2 = b[-2 * w], pym3 = b[-3 * w], pyM = b[w], pyM2 = b[2 * w], pyM3 = b[3 * w];
  const float pxm = b[-1], pxm2 = b[-2], pxm3 = b[-3], pxM = b[1], pxM2 = b[2], pxM3 = b[3];
  const float guessx = (pxm + pc + pxM) * 2.0f - pxM2 - pxm2;
  const float diffx = (fabsf(pxm2 - pc) + fabsf(pxM2 - pc) + fabsf(pxm - pxM)) * 3.0f + (fabsf(pxM3 - pxM) + fabsf(pxm3 - pxm)) * 2.0f;
  const float guessy = (pym + pc + pyM) * 2.0f - pyM2 - pym2;
  const float diffy = (fabsf(pym2 - pc) + fabsf(pyM2 - pc) + fabsf(pym - pyM)) * 3.0f + (fabsf(pyM3 - pyM) + fabsf(pym3 - pym)) * 2.0f;
  if(diffx > diffy) return fmaxf(fminf(guessy * .25f, fmaxf(pym, pyM)), fminf(pym, pyM));
  return fmaxf(fminf(guessx * .25f, fmaxf(pxm, pxM)), fminf(pxm, pxM));
}
// a site's own sample as the output buffer holds it after the first two passes
__device__ __forceinline__ float own_at(const ppg_frame_t &F, int y, int x)
{
  return (in_ring(F, y, x) ? F.in : F.input)[(size_t)y * F.width + x];
}

__global__ void __launch_bounds__(PNT) ppg_kernel(ppg_frame_t F, float4 *__restrict__ out)
{
  const int i = blockIdx.x * PNT + threadIdx.x, j = blockIdx.y;
  if(i >= F.width) return;
  const size_t p = (size_t)j * F.width + i;
  const int c = ppg_fc(j, i, F.filters);
  const bool ring = in_ring(F, j, i);
  float4 o;
  o.w = ring ? out[p].w : 0.0f; // the border loop does not write alpha, the green pass writes 0
  if(j == 0 || i == 0 || j == F.height - 1 || i == F.width - 1)
  { // outermost pixels: not touched by the red/blue pass
    o.x = border_value(F, j, i, 0);
    o.y = border_value(F, j, i, 1);
    o.z = border_value(F, j, i, 2);
    out[p] = o;
    return;
  }
  const float g = green_at(F, j, i);
  float r, b;
  if(c & 1)
  { // green site :151-169: red and blue from the two pairs of direct neighbours
    const float gt = green_at(F, j - 1, i), gb = green_at(F, j + 1, i), gl = green_at(F, j, i - 1), gr = green_at(F, j, i + 1);
    const float vert = (own_at(F, j - 1, i) + own_at(F, j + 1, i) + 2.0f * g - gt - gb) * .5f;
    const float horz = (own_at(F, j, i - 1) + own_at(F, j, i + 1) + 2.0f * g - gl - gr) * .5f;
    if(ppg_fc(j, i + 1, F.filters) == 0)
    { // red neighbours in the row, blue ones in the column
      b = vert;
      r = horz;
    }
    else
    {
      r = vert;
      b = horz;
    }
  }
  else
  { // red or blue site :170-203: the other of the two from the diagonal neighbours, along the flatter diagonal
    const float ntl = own_at(F, j - 1, i - 1), ntr = own_at(F, j - 1, i + 1), nbl = own_at(F, j + 1, i - 1), nbr = own_at(F, j + 1, i + 1);
    const float gtl = green_at(F, j - 1, i - 1), gtr = green_at(F, j - 1, i + 1), gbl = green_at(F, j + 1, i - 1), gbr = green_at(F, j + 1, i + 1);
    const float diff1 = fabsf(ntl - nbr) + fabsf(gtl - g) + fabsf(gbr - g);
    const float guess1 = ntl + nbr + 2.0f * g - gtl - gbr;
    const float diff2 = fabsf(ntr - nbl) + fabsf(gtr - g) + fabsf(gbl - g);
    const float guess2 = ntr + nbl + 2.0f * g - gtr - gbl;
    float other;
    if(diff1 > diff2)
      other = guess2 * .5f;
    else if(diff1 < diff2)
      other = guess1 * .5f;
    else
      other = (guess1 + guess2) * .25f;
    const float own = own_at(F, j, i);
    r = c == 0 ? own : other;
    b = c == 0 ? other : own;
  }
  o.x = r;
  o.y = g;
  o.z = b;
  out[p] = o;
}

// pre_median_b with one pass :136-180: green sites of the interior become a thresholded median of their nine green
// neighbours (the reference's exchange sort, compare for compare, so that NaNs land where they land there)
__global__ void __launch_bounds__(PNT) pre_median_kernel(const float *__restrict__ in, float *__restrict__ out, int width, int height, unsigned filters,
                                                         float threshold)
{
  const int col = blockIdx.x * PNT + threadIdx.x, row = blockIdx.y;
  if(col >= width) return;
  const size_t p = (size_t)row * width + col;
  const float centre = in[p];
  const int c = ppg_fc(row, col, filters);
  // the row loop starts at column 3 or 4, whichever is green, and steps by two: the green sites of that row
  if(row < 3 || row >= height - 3 || col < 3 || col >= width - 3 || (c != 1 && c != 3))
  {
    out[p] = centre;
    return;
  }
  float med[9];
  int cnt = 0;
  {
    const int dy[9] = { -2, -1, -1, 0, 0, 0, 1, 1, 2 }, dx[9] = { 0, -1, 1, -2, 0, 2, -1, 1, 0 };
#pragma unroll
    for(int k = 0; k < 9; k++)
    {
      const float v = in[p + (ptrdiff_t)width * dy[k] + dx[k]];
      if(fabsf(v - centre) < threshold)
      {
        med[k] = v;
        cnt++;
      }
      else
        med[k] = 64.0f + v;
    }
  }
#pragma unroll
  for(int a = 0; a < 8; a++)
#pragma unroll
    for(int b = a + 1; b < 9; b++)
      if(med[a] > med[b])
      {
        const float t = med[b];
        med[b] = med[a];
        med[a] = t;
      }
  float result = med[4] - 64.0f;
  if(cnt != 1)
  { // med[(cnt - 1) / 2] without indexing the register array dynamically; cnt == 0 reads med[0] like (0 - 1) / 2 == 0 in C
    const int k = (cnt - 1) / 2;
    result = med[0];
#pragma unroll
    for(int q = 1; q < 5; q++)
      if(k == q) result = med[q];
  }
  out[p] = result;
}
// the two passthrough methods, iop/demosaic/passthrough.c:22-88: every channel = the sample (monochrome sensors), or the
// sample in its CFA colour and zeros elsewhere (debug view).  Lane 3 is not written by the reference: kept.  The Bayer
// colour is taken from the sensor's filters word at the frame's own coordinates (demosaic.c:1117 passes roi_out with its
// origin zeroed), the X-Trans colour through roi_in.
__global__ void __launch_bounds__(PNT) passthrough_kernel(const float *__restrict__ in, float4 *__restrict__ out, int width, int height, int colour,
                                                          unsigned filters, int x0, int y0, const unsigned char *__restrict__ xtrans36)
{
  const int col = blockIdx.x * PNT + threadIdx.x, row = blockIdx.y;
  if(col >= width) return;
  const size_t p = (size_t)row * width + col;
  const float v = in[p];
  float4 o = make_float4(v, v, v, out[p].w);
  if(colour)
  {
    const int ch = filters == 9u ? xtrans36[((row + 600 + y0) % 6) * 6 + (col + 600 + x0) % 6] : ppg_fc(row, col, filters);
    o.x = ch == 0 ? v : 0.0f;
    o.y = ch == 1 ? v : 0.0f;
    o.z = ch == 2 ? v : 0.0f;
  }
  out[p] = o;
}
// the half-size "downsample" method, demosaic.c:480-532 (Bayer, three colours): the mean of each colour's samples in the 2x2
// block behind the output pixel, the block clamped into the frame at odd edges; alpha 0
__global__ void __launch_bounds__(PNT) downsample_kernel(const float *__restrict__ in, float4 *__restrict__ out, int width, int height, int out_width, unsigned filters)
{
  const int x = blockIdx.x * PNT + threadIdx.x, y = blockIdx.y;
  if(x >= out_width) return;
  const int px = min(2 * x, width - 1), py = min(2 * y, height - 1);
  float cam[3] = { 0.f, 0.f, 0.f };
  int samples[3] = { 0, 0, 0 };
#pragma unroll
  for(int j = 0; j < 2; j++)
#pragma unroll
    for(int i = 0; i < 2; i++)
    {
      const int xx = min(px + i, width - 1), yy = min(py + j, height - 1);
      const int c = ppg_fc(yy, xx, filters);
      const float v = in[(size_t)yy * width + xx];
#pragma unroll
      for(int q = 0; q < 3; q++)
        if(q == c)
        {
          cam[q] += v;
          samples[q]++;
        }
    }
#pragma unroll
  for(int q = 0; q < 3; q++)
    if(samples[q] > 0) cam[q] /= (float)samples[q];
  out[(size_t)y * out_width + x] = make_float4(cam[0], cam[1], cam[2], 0.0f);
}
// the same for a four-colour Bayer sensor (demosaic.c:514-521): four camera primaries, then R, G, B = CAM_to_RGB rows times them,
// products and sums in double, the running value rounded to float after every term as the reference's float accumulator is
struct cam_to_rgb_t
{
  double m[12];
};
__global__ void __launch_bounds__(PNT) downsample4_kernel(const float *__restrict__ in, float4 *__restrict__ out, int width, int height, int out_width, unsigned filters,
                                                          cam_to_rgb_t M)
{
  const int x = blockIdx.x * PNT + threadIdx.x, y = blockIdx.y;
  if(x >= out_width) return;
  const int px = min(2 * x, width - 1), py = min(2 * y, height - 1);
  float cam[4] = { 0.f, 0.f, 0.f, 0.f };
  int samples[4] = { 0, 0, 0, 0 };
#pragma unroll
  for(int j = 0; j < 2; j++)
#pragma unroll
    for(int i = 0; i < 2; i++)
    {
      const int xx = min(px + i, width - 1), yy = min(py + j, height - 1);
      const int c = ppg_fc(yy, xx, filters);
      const float v = in[(size_t)yy * width + xx];
#pragma unroll
      for(int q = 0; q < 4; q++)
        if(q == c)
        {
          cam[q] += v;
          samples[q]++;
        }
    }
#pragma unroll
  for(int q = 0; q < 4; q++)
    if(samples[q] > 0) cam[q] /= (float)samples[q];
  float rgb[3];
#pragma unroll
  for(int c = 0; c < 3; c++)
  {
    float acc = 0.0f;
#pragma unroll
    for(int k = 0; k < 4; k++) acc = (float)((double)acc + M.m[4 * c + k] * (double)cam[k]);
    rgb[c] = acc;
  }
  out[(size_t)y * out_width + x] = make_float4(rgb[0], rgb[1], rgb[2], 0.0f);
}
// The same method on an X-Trans sensor, demosaic.c:543-666.  A 2x2 block of the 6x6 pattern misses one or two colours; a missing
// colour is rebuilt from the nearest same-colour photosite of each quadrant around the block centre (8x8 window, first in raster
// order wins a tie), bilinear inside the rectangle those four span, their plain mean when the frame edge hides a quadrant.
// The reference walks the window once and sorts sites into quadrants; a site's quadrant is a function of its position alone, so
// four 4x4 walks in the same order find the same sites with nothing indexed at run time.  ("nearest overall", the reference's last
// fallback, is only reached when no quadrant found anything, i.e. when it is still 0.)
//
// The sensor's table arrives as two 36-bit words (rows 0-2, rows 3-5, two bits a site) already rotated by roi_in's origin.
struct xtrans_words_t
{
  unsigned long long lo, hi;
};
__device__ __forceinline__ int xw_colour(const xtrans_words_t &T, int r6, int c6)
{
  const unsigned long long w = r6 < 3 ? T.lo : T.hi;
  return (int)(w >> (((r6 < 3 ? r6 : r6 - 3) * 6 + c6) * 2)) & 3;
}
struct quadrant_t
{
  float value, dist;
  int x, y;
  bool found;
};
template <int Q>
__device__ __forceinline__ quadrant_t nearest_in_quadrant(const float *__restrict__ in, int width, int height, int px, int py, const xtrans_words_t &T, int colour)
{
  quadrant_t q = { 0.0f, INFINITY, 0, 0, false };
  const float cx = px + 0.5f, cy = py + 0.5f;
  const int xs = (Q & 1) ? px + 1 : px - 3, ys = (Q & 2) ? py + 1 : py - 3;
  int r6 = (ys + 6) % 6;
#pragma unroll
  for(int j = 0; j < 4; j++)
  {
    const int yy = ys + j;
    int c6 = (xs + 6) % 6;
#pragma unroll
    for(int i = 0; i < 4; i++)
    {
      const int xx = xs + i;
      if(yy >= 0 && yy < height && xx >= 0 && xx < width && xw_colour(T, r6, c6) == colour)
      {
        const float dx = xx - cx, dy = yy - cy;
        const float d2 = dx * dx + dy * dy;
        if(d2 < q.dist)
        {
          q.dist = d2;
          q.value = in[(size_t)yy * width + xx];
          q.x = xx;
          q.y = yy;
          q.found = true;
        }
      }
      c6 = c6 == 5 ? 0 : c6 + 1;
    }
    r6 = r6 == 5 ? 0 : r6 + 1;
  }
  return q;
}
__device__ float xtrans_missing_colour(const float *__restrict__ in, int width, int height, int px, int py, const xtrans_words_t &T, int colour)
{
  const quadrant_t q0 = nearest_in_quadrant<0>(in, width, height, px, py, T, colour), q1 = nearest_in_quadrant<1>(in, width, height, px, py, T, colour),
                   q2 = nearest_in_quadrant<2>(in, width, height, px, py, T, colour), q3 = nearest_in_quadrant<3>(in, width, height, px, py, T, colour);
  if(q0.found && q1.found && q2.found && q3.found)
  {
    const float cx = px + 0.5f, cy = py + 0.5f;
    const float x_left = 0.5f * (q0.x + q2.x), x_right = 0.5f * (q1.x + q3.x), y_top = 0.5f * (q0.y + q1.y), y_bottom = 0.5f * (q2.y + q3.y);
    const float tx = fminf(fmaxf((cx - x_left) / fmaxf(x_right - x_left, 1e-6f), 0.0f), 1.0f);
    const float ty = fminf(fmaxf((cy - y_top) / fmaxf(y_bottom - y_top, 1e-6f), 0.0f), 1.0f);
    const float top = q0.value + tx * (q1.value - q0.value);
    const float bottom = q2.value + tx * (q3.value - q2.value);
    return top + ty * (bottom - top);
  }
  float sum = 0.0f;
  int count = 0;
  if(q0.found) sum += q0.value, count++;
  if(q1.found) sum += q1.value, count++;
  if(q2.found) sum += q2.value, count++;
  if(q3.found) sum += q3.value, count++;
  return count > 0 ? sum / (float)count : 0.0f;
}
__global__ void __launch_bounds__(PNT) downsample_xtrans_kernel(const float *__restrict__ in, float4 *__restrict__ out, int width, int height, int out_width,
                                                                xtrans_words_t T)
{
  const int x = blockIdx.x * PNT + threadIdx.x, y = blockIdx.y;
  if(x >= out_width) return;
  const int px = min(2 * x, width - 1), py = min(2 * y, height - 1);
  float rgb[3] = { 0.f, 0.f, 0.f };
  int samples[3] = { 0, 0, 0 };
#pragma unroll
  for(int j = 0; j < 2; j++)
#pragma unroll
    for(int i = 0; i < 2; i++)
    {
      const int xx = min(px + i, width - 1), yy = min(py + j, height - 1);
      const int c = xw_colour(T, yy % 6, xx % 6);
      const float v = in[(size_t)yy * width + xx];
#pragma unroll
      for(int q = 0; q < 3; q++)
        if(q == c)
        {
          rgb[q] += v;
          samples[q]++;
        }
    }
#pragma unroll
  for(int q = 0; q < 3; q++) rgb[q] = samples[q] > 0 ? rgb[q] / (float)samples[q] : xtrans_missing_colour(in, width, height, px, py, T, q);
  out[(size_t)y * out_width + x] = make_float4(rgb[0], rgb[1], rgb[2], 0.0f);
}
// the table FCxtrans() sees through roi_in (develop/imageop_math.h:216-219), packed for xw_colour()
inline xtrans_words_t pack_xtrans(const unsigned char *xtrans36, int x0, int y0)
{
  xtrans_words_t T = { 0ull, 0ull };
  for(int r = 0; r < 6; r++)
    for(int c = 0; c < 6; c++)
    {
      const unsigned long long v = xtrans36[((r + y0 + 600) % 6) * 6 + (c + x0 + 600) % 6] & 3u;
      (r < 3 ? T.lo : T.hi) |= v << (((r % 3) * 6 + c) * 2);
    }
  return T;
}
} // namespace

#ifndef B200_KERNELS_ON_CPU
namespace b200
{
// demosaic.c:1101-1108 for a three-colour Bayer sensor; out is (width + 1) / 2 x (height + 1) / 2
int downsample_demosaic_dev(const float *d_in, float *d_out, int width, int height, uint32_t filters, cudaStream_t s)
{
  const int ow = (width + 1) / 2, oh = (height + 1) / 2;
  if(oh > 65535) return fail(B200_ERR_ARG, "demosaic: frame height %d", height);
  downsample_kernel<<<dim3((unsigned)((ow + PNT - 1) / PNT), (unsigned)oh), PNT, 0, s>>>(d_in, (float4 *)d_out, width, height, ow, filters);
  B200_CUDA_TRY(cudaGetLastError());
  return B200_OK;
}
// demosaic.c:1106-1107 for a four-colour Bayer sensor; cam_to_rgb = data->CAM_to_RGB
int downsample4_demosaic_dev(const float *d_in, float *d_out, int width, int height, uint32_t filters, const double cam_to_rgb[3][4], cudaStream_t s)
{
  const int ow = (width + 1) / 2, oh = (height + 1) / 2;
  if(oh > 65535) return fail(B200_ERR_ARG, "demosaic: frame height %d", height);
  cam_to_rgb_t M;
  for(int k = 0; k < 12; k++) M.m[k] = cam_to_rgb[k / 4][k % 4];
  downsample4_kernel<<<dim3((unsigned)((ow + PNT - 1) / PNT), (unsigned)oh), PNT, 0, s>>>(d_in, (float4 *)d_out, width, height, ow, filters, M);
  B200_CUDA_TRY(cudaGetLastError());
  return B200_OK;
}
// demosaic.c:1103-1104 for an X-Trans sensor: (x0, y0) = roi_in's origin, xtrans = piece->dsc_in.xtrans
int downsample_xtrans_demosaic_dev(const float *d_in, float *d_out, int width, int height, int x0, int y0, const uint8_t xtrans[6][6], cudaStream_t s)
{
  const int ow = (width + 1) / 2, oh = (height + 1) / 2;
  if(oh > 65535) return fail(B200_ERR_ARG, "demosaic: frame height %d", height);
  downsample_xtrans_kernel<<<dim3((unsigned)((ow + PNT - 1) / PNT), (unsigned)oh), PNT, 0, s>>>(d_in, (float4 *)d_out, width, height, ow,
                                                                                              pack_xtrans(&xtrans[0][0], x0, y0));
  B200_CUDA_TRY(cudaGetLastError());
  return B200_OK;
}
// demosaic.c:1111-1118.  filters: piece->dsc_in.filters (not ROI-shifted); xtrans: piece->dsc_in.xtrans (used when filters == 9)
int passthrough_demosaic_dev(const float *d_in, float *d_out, int width, int height, int colour, uint32_t filters, int x0, int y0, const uint8_t xtrans[6][6],
                             cudaStream_t s)
{
  if(height > 65535) return fail(B200_ERR_ARG, "demosaic: frame height %d", height);
  void *dx = nullptr;
  if(colour && filters == 9u)
  {
    int rc = scratch(SLOT_SMALL + 3, 64, &dx);
    if(rc) return rc;
    B200_CUDA_TRY(cudaMemcpyAsync(dx, xtrans, 36, cudaMemcpyHostToDevice, s)); // pageable source: staged before the call returns
  }
  passthrough_kernel<<<dim3((unsigned)((width + PNT - 1) / PNT), (unsigned)height), PNT, 0, s>>>(d_in, (float4 *)d_out, width, height, colour, filters, x0, y0,
                                                                                                 (const unsigned char *)dx);
  B200_CUDA_TRY(cudaGetLastError());
  return B200_OK;
}
// demosaic.c:1218-1226: d_in = the (green-equilibrated) mosaic, filters = ROI-shifted word, out keeps the alpha of its
// outer three pixels
int ppg_demosaic_dev(const float *d_in, float *d_out, int width, int height, uint32_t filters, float median_thrs, cudaStream_t s)
{
  if(width < 8 || height < 8) return fail(B200_ERR_UNSUPPORTED, "demosaic: PPG on a %dx%d frame (the reference's border loop does not terminate under 6 px)", width, height);
  if(height > 65535) return fail(B200_ERR_ARG, "demosaic: PPG frame height %d", height);
  ppg_frame_t F = { d_in, d_in, width, height, filters };
  const dim3 grid((unsigned)((width + PNT - 1) / PNT), (unsigned)height);
  if(median_thrs > 0.0f)
  {
    void *med = nullptr;
    int rc = scratch(SLOT_TMP2, (size_t)width * height * sizeof(float), &med);
    if(rc) return rc;
    pre_median_kernel<<<grid, PNT, 0, s>>>(d_in, (float *)med, width, height, filters, median_thrs);
    B200_CUDA_TRY(cudaGetLastError());
    F.input = (const float *)med;
  }
  ppg_kernel<<<grid, PNT, 0, s>>>(F, (float4 *)d_out);
  B200_CUDA_TRY(cudaGetLastError());
  return B200_OK;
}
} // namespace b200
#endif
