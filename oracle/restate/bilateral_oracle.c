/* CPU restatement of the bilateral grid behind local contrast's "bilateral grid" mode.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src/pixel/bilateral.c: dt_bilateral_grid_size :51-78, image_to_grid :131-144, dt_bilateral_splat
 * :182-265, blur_line :302-338, blur_line_z :267-300, dt_bilateral_blur :341-353, dt_bilateral_slice :355-393; called from
 * iop/bilat.c process() :346-353 with sigma_s = data->sigma_s / module scale.  Pinned bit-for-bit against that file compiled
 * in place with ONE splat slice (oracle/_ref: ref_bilateral.c): the reference splats one horizontal slice per OpenMP thread and
 * adds the partial grids afterwards, so its rounding depends on the thread count; one slice is raster order.
 */
#include "oracle_common.h"
#include <stdlib.h>
#include <string.h>

typedef struct
{
  size_t size_x, size_y, size_z;
  int width, height;
  float sigma_s, sigma_r;
  float *buf;
} grid_t;

static float clamps(float a, float lo, float hi) { return a > lo ? (a < hi ? a : hi) : lo; } /* CLAMPS, math/math.h */

static void grid_size(grid_t *b, int width, int height, float L_range, float sigma_s, float sigma_r)
{
  if(sigma_s < 0.5) sigma_s = 0.5;
  /* CLAMPS((int)roundf(..), 4, MAX) on ints, stored in floats */
  int ix = (int)roundf(width / sigma_s), iy = (int)roundf(height / sigma_s), iz = (int)roundf(L_range / sigma_r);
  const float _x = ix > 4 ? (ix < 3000 ? ix : 3000) : 4, _y = iy > 4 ? (iy < 3000 ? iy : 3000) : 4, _z = iz > 4 ? (iz < 50 ? iz : 50) : 4;
  const float sy = height / _y, sx = width / _x;
  b->sigma_s = sy > sx ? sy : sx; /* MAX(height / _y, width / _x) */
  b->sigma_r = L_range / _z;
  b->size_x = (int)ceilf(width / b->sigma_s) + 1;
  b->size_y = (int)ceilf(height / b->sigma_s) + 1;
  b->size_z = (int)ceilf(L_range / b->sigma_r) + 1;
  b->width = width;
  b->height = height;
}
/* cell and weights of pixel (i, j) with lightness L */
static size_t to_grid(const grid_t *b, int i, int j, float L, float *xf, float *yf, float *zf)
{
  const float x = clamps(i / b->sigma_s, 0, b->size_x - 1), y = clamps(j / b->sigma_s, 0, b->size_y - 1), z = clamps(L / b->sigma_r, 0, b->size_z - 1);
  const int xi = (int)x < (int)b->size_x - 2 ? (int)x : (int)b->size_x - 2;
  const int yi = (int)y < (int)b->size_y - 2 ? (int)y : (int)b->size_y - 2;
  const int zi = (int)z < (int)b->size_z - 2 ? (int)z : (int)b->size_z - 2;
  *xf = x - xi;
  *yf = y - yi;
  *zf = z - zi;
  return ((xi + yi * b->size_x) * b->size_z) + zi;
}
static void splat(const grid_t *b, const float *in)
{
  const size_t ox = b->size_z, oy = b->size_x * b->size_z;
  const float sigma_s = b->sigma_s * b->sigma_s;
  for(int j = 0; j < b->height; j++)
    for(int i = 0; i < b->width; i++)
    {
      float xf, yf, zf;
      const float L = in[4 * ((size_t)j * b->width + i)];
      const size_t gi = to_grid(b, i, j, L, &xf, &yf, &zf);
      const float contrib[4] = { (1.0f - xf) * (1.0f - yf) * 100.0f / sigma_s, xf * (1.0f - yf) * 100.0f / sigma_s, (1.0f - xf) * yf * 100.0f / sigma_s,
                                 xf * yf * 100.0f / sigma_s };
      const size_t off[4] = { 0, ox, oy, ox + oy };
      for(int k = 0; k < 4; k++)
      {
        b->buf[gi + off[k]] += contrib[k] * (1.0f - zf);
        b->buf[gi + off[k] + 1] += contrib[k] * zf;
      }
    }
}
/* in-place 1-4-6-4-1 along the axis with stride o3 (size3 cells), for every line of the two other axes */
static void blur_line(float *buf, size_t o1, size_t o2, size_t o3, size_t s1, size_t s2, size_t s3)
{
  const float w0 = 6.f / 16.f, w1 = 4.f / 16.f, w2 = 1.f / 16.f;
  for(size_t k = 0; k < s1; k++)
    for(size_t j = 0; j < s2; j++)
    {
      float *p = buf + k * o1 + j * o2;
      float tmp1 = p[0];
      p[0] = p[0] * w0 + w1 * p[o3] + w2 * p[2 * o3];
      p += o3;
      float tmp2 = p[0];
      p[0] = p[0] * w0 + w1 * (p[o3] + tmp1) + w2 * p[2 * o3];
      p += o3;
      for(size_t i = 2; i + 2 < s3; i++)
      {
        const float tmp3 = p[0];
        p[0] = p[0] * w0 + w1 * (p[o3] + tmp2) + w2 * (p[2 * o3] + tmp1);
        p += o3;
        tmp1 = tmp2;
        tmp2 = tmp3;
      }
      const float tmp3 = p[0];
      p[0] = p[0] * w0 + w1 * (p[o3] + tmp2) + w2 * tmp1;
      p += o3;
      p[0] = p[0] * w0 + w1 * tmp3 + w2 * tmp2;
    }
}
/* the derivative-of-Gaussian along z */
static void blur_line_z(float *buf, size_t o1, size_t o2, size_t o3, size_t s1, size_t s2, size_t s3)
{
  const float w1 = 4.f / 16.f, w2 = 2.f / 16.f;
  for(size_t k = 0; k < s1; k++)
    for(size_t j = 0; j < s2; j++)
    {
      float *p = buf + k * o1 + j * o2;
      float tmp1 = p[0];
      p[0] = w1 * p[o3] + w2 * p[2 * o3];
      p += o3;
      float tmp2 = p[0];
      p[0] = w1 * (p[o3] - tmp1) + w2 * p[2 * o3];
      p += o3;
      for(size_t i = 2; i + 2 < s3; i++)
      {
        const float tmp3 = p[0];
        p[0] = +w1 * (p[o3] - tmp2) + w2 * (p[2 * o3] - tmp1);
        p += o3;
        tmp1 = tmp2;
        tmp2 = tmp3;
      }
      const float tmp3 = p[0];
      p[0] = w1 * (p[o3] - tmp2) - w2 * tmp1;
      p += o3;
      p[0] = -w1 * tmp3 - w2 * tmp2;
    }
}

/* grid != NULL: also return the grid after the splat (blur == 0) or after the blur; dims = size_x, size_y, size_z */
int orc_bilateral(const float *in, float *out, int width, int height, float sigma_s, float sigma_r, float detail, float *grid, int max_floats, int dims[3],
                  int blur)
{
  grid_t b;
  grid_size(&b, width, height, 100.0f, sigma_s, sigma_r);
  const size_t n = b.size_x * b.size_y * b.size_z;
  b.buf = calloc(n, sizeof(float));
  splat(&b, in);
  const size_t ox = b.size_z, oy = b.size_x * b.size_z, oz = 1;
  if(grid && !blur && n <= (size_t)max_floats) memcpy(grid, b.buf, n * sizeof(float));
  blur_line(b.buf, oz, oy, ox, b.size_z, b.size_y, b.size_x);
  blur_line(b.buf, oz, ox, oy, b.size_z, b.size_x, b.size_y);
  blur_line_z(b.buf, ox, oy, oz, b.size_x, b.size_y, b.size_z);
  if(grid && blur && n <= (size_t)max_floats) memcpy(grid, b.buf, n * sizeof(float));
  if(dims) dims[0] = (int)b.size_x, dims[1] = (int)b.size_y, dims[2] = (int)b.size_z;
  if(out)
  {
    const float norm = -detail * b.sigma_r * 0.04f;
    for(int j = 0; j < height; j++)
      for(int i = 0; i < width; i++)
      {
        const size_t index = 4 * ((size_t)j * width + i);
        float xf, yf, zf;
        const float L = in[index];
        const size_t gi = to_grid(&b, i, j, L, &xf, &yf, &zf);
        const float *g = b.buf + gi;
        out[index] = fmaxf(0.0f, L + norm * (g[0] * (1.0f - xf) * (1.0f - yf) * (1.0f - zf) + g[ox] * (xf) * (1.0f - yf) * (1.0f - zf)
                                             + g[oy] * (1.0f - xf) * (yf) * (1.0f - zf) + g[ox + oy] * (xf) * (yf) * (1.0f - zf)
                                             + g[oz] * (1.0f - xf) * (1.0f - yf) * (zf) + g[ox + oz] * (xf) * (1.0f - yf) * (zf)
                                             + g[oy + oz] * (1.0f - xf) * (yf) * (zf) + g[ox + oy + oz] * (xf) * (yf) * (zf)));
        out[index + 1] = in[index + 1];
        out[index + 2] = in[index + 2];
        out[index + 3] = in[index + 3];
      }
  }
  free(b.buf);
  return 0;
}
