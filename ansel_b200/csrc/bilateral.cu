// The bilateral grid behind local contrast's "bilateral grid" mode.
//
// Reference: pixel/bilateral.c dt_bilateral_grid_size :51-78, image_to_grid :131-144, dt_bilateral_splat :182-265,
// blur_line :302-338, blur_line_z :267-300, dt_bilateral_blur :341-353, dt_bilateral_slice :355-393; iop/bilat.c process()
// :346-353, tiling_callback :255-280.
//
// The reference splats with one horizontal slice of the frame per OpenMP thread and adds the slices' partial grids
// afterwards: which additions a grid cell sees in which order depends on the thread count (tests/test_cpu_bilateral.py
// counts the pixels that differ between 1 and 8 threads).  One slice is plain raster order, and that is what is computed
// here, deterministically and without atomics: a thread owns one (x, y) column of the grid, visits the pixels of its
// footprint (two cells wide in x and y) in raster order and accumulates its size_z cells privately -- every cell receives
// its contributions in the order the one-thread reference adds them.  4 W H pixel visits in all, independent of sigma.
// The three blurs run in place along lines exactly as the reference's (a thread per line, the two previous values in
// registers); the slice is a trilinear lookup per pixel.  The grid (<= 3000 x 3000 x 51 cells, usually a few MB) lives in L2.
#ifndef B200_KERNELS_ON_CPU // tests/emul compiles the kernels of this file with g++ to check them against the oracle without a GPU
#include "runtime.h"
#endif
#include <math.h>
#include <string.h>

namespace
{
constexpr int BNT = 128;
constexpr int MAX_Z = 52; // DT_COMMON_BILATERAL_MAX_RES_R (50) + 1, + 1 for the pair a pixel writes

struct bgrid_t
{
  int size_x, size_y, size_z, width, height;
  float sigma_s, sigma_r;
};

__device__ __forceinline__ float clamps(float a, float lo, float hi) { return a > lo ? (a < hi ? a : hi) : lo; } // CLAMPS, math/math.h
// image_to_grid :131-144 for one axis: cell index and fraction
__device__ __forceinline__ int cell_of(float v, int size, float *frac)
{
  const float c = clamps(v, 0.0f, (float)(size - 1));
  const int ci = (int)c < size - 2 ? (int)c : size - 2;
  *frac = c - (float)ci;
  return ci;
}

// dt_bilateral_splat :182-265 with one slice; buf is written, not accumulated into
__global__ void __launch_bounds__(BNT) bilateral_splat_kernel(const float4 *__restrict__ in, float *__restrict__ buf, const bgrid_t G)
{
  const int cell = blockIdx.x * BNT + threadIdx.x;
  if(cell >= G.size_x * G.size_y) return;
  const int gx = cell % G.size_x, gy = cell / G.size_x;
  float acc[MAX_Z];
  for(int z = 0; z < G.size_z; z++) acc[z] = 0.0f;
  const float sigma_s2 = G.sigma_s * G.sigma_s;
  // pixels whose cell index along an axis is g - 1 or g: a conservative range, checked exactly per pixel below
  const int j0 = max(0, (int)floorf((float)(gy - 1) * G.sigma_s) - 1), j1 = min(G.height - 1, (int)ceilf((float)(gy + 1) * G.sigma_s) + 1);
  const int i0 = max(0, (int)floorf((float)(gx - 1) * G.sigma_s) - 1), i1 = min(G.width - 1, (int)ceilf((float)(gx + 1) * G.sigma_s) + 1);
  for(int j = j0; j <= j1; j++)
  {
    float yf;
    const int yi = cell_of((float)j / G.sigma_s, G.size_y, &yf);
    if(yi != gy && yi + 1 != gy) continue;
    const float wy = yi == gy ? 1.0f - yf : yf;
    for(int i = i0; i <= i1; i++)
    {
      float xf, zf;
      const int xi = cell_of((float)i / G.sigma_s, G.size_x, &xf);
      if(xi != gx && xi + 1 != gx) continue;
      const float L = in[(size_t)j * G.width + i].x;
      const int zi = cell_of(L / G.sigma_r, G.size_z, &zf);
      // (1 - xf) (1 - yf) 100 / sigma_s^2 and its three siblings :231-237, evaluated left to right
      const float wx = xi == gx ? 1.0f - xf : xf;
      const float contrib = wx * wy * 100.0f / sigma_s2;
      acc[zi] += contrib * (1.0f - zf);
      acc[zi + 1] += contrib * zf;
    }
  }
  float *col = buf + (size_t)cell * G.size_z;
  for(int z = 0; z < G.size_z; z++) col[z] = acc[z];
}

// blur_line :302-338 (DERIVATIVE = false) and blur_line_z :267-300 (true): in place along the axis of stride o3, one
// thread per line (k over s1 with stride o1, j over s2 with stride o2)
template <bool DERIVATIVE>
__global__ void __launch_bounds__(BNT) bilateral_blur_kernel(float *__restrict__ buf, size_t o1, size_t o2, size_t o3, int s1, int s2, int s3)
{
  const size_t line = (size_t)blockIdx.x * BNT + threadIdx.x;
  if(line >= (size_t)s1 * s2) return;
  const int k = (int)(line % s1), j = (int)(line / s1);
  float *p = buf + (size_t)k * o1 + (size_t)j * o2;
  if(DERIVATIVE)
  {
    const float w1 = 4.f / 16.f, w2 = 2.f / 16.f;
    float tmp1 = p[0];
    p[0] = w1 * p[o3] + w2 * p[2 * o3];
    p += o3;
    float tmp2 = p[0];
    p[0] = w1 * (p[o3] - tmp1) + w2 * p[2 * o3];
    p += o3;
    for(int i = 2; i < s3 - 2; i++)
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
  else
  {
    const float w0 = 6.f / 16.f, w1 = 4.f / 16.f, w2 = 1.f / 16.f;
    float tmp1 = p[0];
    p[0] = p[0] * w0 + w1 * p[o3] + w2 * p[2 * o3];
    p += o3;
    float tmp2 = p[0];
    p[0] = p[0] * w0 + w1 * (p[o3] + tmp1) + w2 * p[2 * o3];
    p += o3;
    for(int i = 2; i < s3 - 2; i++)
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

// dt_bilateral_slice :355-393: L + norm * trilinear(grid), clipped at 0; the other three lanes are the input's
__global__ void __launch_bounds__(BNT) bilateral_slice_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, const float *__restrict__ buf, const bgrid_t G,
                                                              float norm)
{
  const int i = blockIdx.x * BNT + threadIdx.x, j = blockIdx.y;
  if(i >= G.width) return;
  const float4 p = in[(size_t)j * G.width + i];
  float xf, yf, zf;
  const int xi = cell_of((float)i / G.sigma_s, G.size_x, &xf), yi = cell_of((float)j / G.sigma_s, G.size_y, &yf), zi = cell_of(p.x / G.sigma_r, G.size_z, &zf);
  const size_t ox = G.size_z, oy = (size_t)G.size_x * G.size_z;
  const float *g = buf + ((size_t)xi + (size_t)yi * G.size_x) * G.size_z + zi;
  const float v = g[0] * (1.0f - xf) * (1.0f - yf) * (1.0f - zf) + g[ox] * (xf) * (1.0f - yf) * (1.0f - zf) + g[oy] * (1.0f - xf) * (yf) * (1.0f - zf)
                  + g[ox + oy] * (xf) * (yf) * (1.0f - zf) + g[1] * (1.0f - xf) * (1.0f - yf) * (zf) + g[ox + 1] * (xf) * (1.0f - yf) * (zf)
                  + g[oy + 1] * (1.0f - xf) * (yf) * (zf) + g[ox + oy + 1] * (xf) * (yf) * (zf);
  out[(size_t)j * G.width + i] = make_float4(fmaxf(0.0f, p.x + norm * v), p.y, p.z, p.w);
}

// dt_bilateral_grid_size :51-78 (host float arithmetic, as in the reference)
inline void bilateral_grid_size(bgrid_t *b, int width, int height, float L_range, float sigma_s, float sigma_r)
{
  if(sigma_s < 0.5) sigma_s = 0.5;
  const int ix = (int)roundf(width / sigma_s), iy = (int)roundf(height / sigma_s), iz = (int)roundf(L_range / sigma_r);
  const float _x = (float)(ix > 4 ? (ix < 3000 ? ix : 3000) : 4), _y = (float)(iy > 4 ? (iy < 3000 ? iy : 3000) : 4),
              _z = (float)(iz > 4 ? (iz < 50 ? iz : 50) : 4);
  const float sy = height / _y, sx = width / _x;
  b->sigma_s = sy > sx ? sy : sx;
  b->sigma_r = L_range / _z;
  b->size_x = (int)ceilf(width / b->sigma_s) + 1;
  b->size_y = (int)ceilf(height / b->sigma_s) + 1;
  b->size_z = (int)ceilf(L_range / b->sigma_r) + 1;
  b->width = width;
  b->height = height;
}
} // namespace

#ifndef B200_KERNELS_ON_CPU
namespace b200
{
// iop/bilat.c process() :346-353: init + splat + blur + slice
int bilateral_grid_dev(const float *d_in, float *d_out, int width, int height, float sigma_s, float sigma_r, float detail, cudaStream_t s)
{
  if(width < 1 || height < 1 || height > 65535) return fail(B200_ERR_ARG, "bilat: frame %d x %d", width, height);
  if(!(sigma_r > 0.0f) || !(sigma_s == sigma_s)) return fail(B200_ERR_ARG, "bilat: sigma_s %g sigma_r %g", sigma_s, sigma_r);
  bgrid_t G;
  bilateral_grid_size(&G, width, height, 100.0f, sigma_s, sigma_r);
  if(G.size_z > MAX_Z) return fail(B200_ERR_ARG, "bilat: grid %d x %d x %d", G.size_x, G.size_y, G.size_z);
  if(G.size_z < 4 || G.size_x < 4 || G.size_y < 4)
    return fail(B200_ERR_UNSUPPORTED, "bilat: a %d x %d x %d grid is narrower than the blur's five taps (the reference runs over the line ends)", G.size_x,
                G.size_y, G.size_z);
  const size_t cells = (size_t)G.size_x * G.size_y, n = cells * G.size_z;
  void *buf = nullptr;
  int rc = scratch(SLOT_TMP0, n * sizeof(float), &buf);
  if(rc) return rc;
  bilateral_splat_kernel<<<(unsigned)((cells + BNT - 1) / BNT), BNT, 0, s>>>((const float4 *)d_in, (float *)buf, G);
  B200_CUDA_TRY(cudaGetLastError());
  const size_t ox = (size_t)G.size_z, oy = (size_t)G.size_x * G.size_z, oz = 1;
  // dt_bilateral_blur :341-353: along x, along y, then the derivative of the Gaussian along z
  bilateral_blur_kernel<false><<<(unsigned)(((size_t)G.size_z * G.size_y + BNT - 1) / BNT), BNT, 0, s>>>((float *)buf, oz, oy, ox, G.size_z, G.size_y, G.size_x);
  bilateral_blur_kernel<false><<<(unsigned)(((size_t)G.size_z * G.size_x + BNT - 1) / BNT), BNT, 0, s>>>((float *)buf, oz, ox, oy, G.size_z, G.size_x, G.size_y);
  bilateral_blur_kernel<true><<<(unsigned)(((size_t)G.size_x * G.size_y + BNT - 1) / BNT), BNT, 0, s>>>((float *)buf, ox, oy, oz, G.size_x, G.size_y, G.size_z);
  B200_CUDA_TRY(cudaGetLastError());
  const float norm = -detail * G.sigma_r * 0.04f; // :358
  bilateral_slice_kernel<<<dim3((unsigned)((width + BNT - 1) / BNT), (unsigned)height), BNT, 0, s>>>((const float4 *)d_in, (float4 *)d_out, (const float *)buf, G, norm);
  B200_CUDA_TRY(cudaGetLastError());
  return B200_OK;
}
// dt_bilateral_memory_use / _singlebuffer_size :80-118 for the tiling callback (the CPU figures, which count three grid rows
// per OpenMP thread; the device needs the grid only)
size_t bilateral_grid_bytes(int width, int height, float sigma_s, float sigma_r)
{
  bgrid_t G;
  bilateral_grid_size(&G, width, height, 100.0f, sigma_s, sigma_r);
  return (size_t)G.size_x * G.size_y * G.size_z * sizeof(float);
}
} // namespace b200
#endif
