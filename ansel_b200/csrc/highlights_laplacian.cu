// highlights, mode "guided laplacians": clipped areas rebuilt at a quarter of the resolution by `iterations` rounds of two multi-scale
// reconstructions (colour from the best-exposed channel's wavelet details, then a diffusion of the colour ratios), blended back under
// a feathered clipping mask.
//
// Reference: src/iop/highlights/laplacian.c process_laplacian :433-575, wavelets_process :374-430, guide_laplacians :85-246,
// heat_PDE_diffusion :248-372; src/iop/highlights/gather.c _compute_laplacian_normalization :223-275, _interpolate_and_mask :67-221,
// _interpolate_and_mask_xtrans :317-422 (_build_xtrans_bilinear_lookup :277-315), _interpolate_and_mask_passthrough :424-455,
// _remosaic_and_replace :457-486, _remosaic_and_replace_xtrans :488-512, _remosaic_and_replace_passthrough :514-541;
// src/pixel/box_filters.c dt_box_mean_4ch :950-971; src/pixel/fast_guided_filter.h interpolate_bilinear :99-152;
// src/pixel/bspline.h decompose_2D_Bspline :351-377; src/iop/noise_generator.h poisson_noise_simd :174-200.
//
// Layout: the full-size [R, G, B, norm] frame and two copies of its clipping mask as RGBA float (3 x 16 B/px), seven quarter-size RGBA
// planes for the wavelets (7 x 1 B/px).  Every stage is one thread per pixel except the column pass of the box mean, whose running
// sum is a recurrence down each of the 4*width float columns (one thread per column, like the reference's vector lanes).
// Algorithmic bytes at the module boundary: 8 B/px on a mosaic; what the stages move per pixel: gather 4 + 32, box mean 16 + 16 twice,
// two reductions to a quarter 2 x (16 + 1), per iteration and scale at a sixteenth of the pixels 2 x (16 + 16 + 32 + 32 + 48 + 16),
// enlargement 1 + 16, composite 4 + 32 + 4: a few hundred launches on 2.8 MP planes, bounded by launch latency and L2, not by HBM.
//
// Arithmetic contract: the reference source under C float semantics (no contraction, IEEE division and square root, glibc's logf /
// sinf / cosf restated in flt32_math.cuh), as restated in oracle/restate/highlights_laplacian_oracle.c which is bit-identical to the
// lines above compiled in place.  ONE value is not the reference's: its normalization vector is an OpenMP float reduction over the
// whole frame, so it depends on the thread count and on the order the threads finish in (one thread's sum stops growing once the
// addends fall under half an ulp of it).  Here it is the sum in double, in a fixed order, rounded once; a caller that wants the
// reference's bits for a given run passes that run's vector (b200_highlights_laplacian_dev).
#ifndef B200_KERNELS_ON_CPU // tests/emul compiles the kernels and the launch sequence of this file with g++
#include "runtime.h"
#else
typedef void *cudaStream_t;
#endif
#include "flt32_math.cuh"
#include "bspline.cuh"
#include <math.h>
#include <string.h>

namespace
{
using namespace bsp; // NT, clip0, max_zero, the two B-spline kernels, splitmix32, xoshiro128plus
constexpr int DS_FACTOR = 4;                            // iop/highlights/common.h:617
constexpr int MAX_NUM_SCALES = 12;                      // :615
constexpr float B_SPLINE_SIGMA = 1.0553651328015339f;   // bspline.h:38
constexpr float B_SPLINE_TO_LAPLACIAN = 3.182727439285017f; // bspline.h:49
constexpr int FIRST_SCALE = 2, LAST_SCALE = 4;          // common.h:625-630
constexpr int NORM_BLOCKS = 592;                        // four per SM

__device__ __forceinline__ int fc(int row, int col, uint32_t filters)
{ // FC(), develop/imageop_math.h:190-193
  return (filters >> ((((row << 1) & 14) + (col & 1)) << 1)) & 3u;
}
__device__ __forceinline__ float sqf(float x) { return x * x; }
struct hl_xtrans_t
{ // the sensor's 6x6 table turned to the ROI origin: FCxtrans(row, col, roi_in, xtrans), develop/imageop_math.h:201-219
  unsigned char v[6][6];
};
__device__ __forceinline__ int fcx(int row, int col, const hl_xtrans_t &xt) { return xt.v[(row + 600) % 6][(col + 600) % 6]; }
// the colour of a site: filters == 9 reads the X-Trans table
__device__ __forceinline__ int hl_colour(int row, int col, uint32_t filters, const hl_xtrans_t &xt) { return filters == 9u ? fcx(row, col, xt) : fc(row, col, filters); }

// ---- normalization: the mean of every colour's samples over ALL sites, gather.c:223-275 ------------------------------------------------
#ifndef B200_KERNELS_ON_CPU // block reductions: not for the thread-by-thread harness, which is handed a vector
// partial[3 * block + c]: rows block, block + gridDim.x, ... of the frame; columns strided over the threads; a fixed tree per block
__global__ void __launch_bounds__(NT) hl_norm_partial_kernel(const float *__restrict__ in, int width, int height, uint32_t filters, hl_xtrans_t xt,
                                                             double *__restrict__ partial)
{
  __shared__ double sh[3][NT];
  double acc[3] = { 0.0, 0.0, 0.0 };
  for(int i = blockIdx.x; i < height; i += gridDim.x)
    for(int j = threadIdx.x; j < width; j += NT)
    {
      if(filters)
      {
        const double v = (double)__ldg(in + (size_t)i * width + j);
        const int c = hl_colour(i, j, filters, xt);
        acc[0] += c == 0 ? v : 0.0;
        acc[1] += c == 1 ? v : 0.0;
        acc[2] += c == 2 ? v : 0.0;
      }
      else
      {
        const float4 p = __ldg((const float4 *)in + (size_t)i * width + j);
        acc[0] += (double)p.x;
        acc[1] += (double)p.y;
        acc[2] += (double)p.z;
      }
    }
  for(int c = 0; c < 3; c++) sh[c][threadIdx.x] = acc[c];
  __syncthreads();
  for(int step = NT / 2; step > 0; step >>= 1)
  {
    if((int)threadIdx.x < step)
      for(int c = 0; c < 3; c++) sh[c][threadIdx.x] += sh[c][threadIdx.x + step];
    __syncthreads();
  }
  if(threadIdx.x < 3) partial[3 * blockIdx.x + threadIdx.x] = sh[threadIdx.x][0];
}
// one block: the partials in a fixed order, divided by the pixel count the way the reference holds it (a float)
__global__ void __launch_bounds__(NT) hl_norm_final_kernel(const double *__restrict__ partial, int blocks, float n_pixels, float *__restrict__ norm)
{
  __shared__ double sh[3][NT];
  double acc[3] = { 0.0, 0.0, 0.0 };
  for(int b = threadIdx.x; b < blocks; b += NT)
    for(int c = 0; c < 3; c++) acc[c] += partial[3 * b + c];
  for(int c = 0; c < 3; c++) sh[c][threadIdx.x] = acc[c];
  __syncthreads();
  for(int step = NT / 2; step > 0; step >>= 1)
  {
    if((int)threadIdx.x < step)
      for(int c = 0; c < 3; c++) sh[c][threadIdx.x] += sh[c][threadIdx.x + step];
    __syncthreads();
  }
  if(threadIdx.x < 3) norm[threadIdx.x] = (float)(sh[threadIdx.x][0] / (double)n_pixels);
  if(threadIdx.x == 3) norm[3] = 1.f;
}

#endif

// ---- gather: a bilinear [R, G, B, norm] frame and its binary clipping flags ----------------------------------------------------------
struct hl_clips_t
{
  float v[4];
};
// the colour c2 (red or blue) around a site of another colour, gather.c:156-180 / :189-214
__device__ __forceinline__ void hl_around(const float *__restrict__ in, size_t ic, size_t ip, size_t in_, int j, int jp, int jn, int i, uint32_t filters,
                                          int c2, float clip, float &value, bool &clipped)
{
  if(fc(i + 1, j, filters) == c2)
  {
    const float north = __ldg(in + ip + j), south = __ldg(in + in_ + j);
    value = (north + south) / 2.f;
    clipped = north > clip || south > clip;
  }
  else if(fc(i, j + 1, filters) == c2)
  {
    const float west = __ldg(in + ic + jp), east = __ldg(in + ic + jn);
    value = (west + east) / 2.f;
    clipped = west > clip || east > clip;
  }
  else
  {
    const float nw = __ldg(in + ip + jp), ne = __ldg(in + ip + jn), se = __ldg(in + in_ + jn), sw = __ldg(in + in_ + jp);
    value = (nw + ne + se + sw) / 4.f;
    clipped = nw > clip || ne > clip || sw > clip || se > clip;
  }
}
__device__ __forceinline__ void hl_store_gathered(float4 *__restrict__ interpolated, float4 *__restrict__ mask, size_t p, float R, float G, float B, bool kr,
                                                  bool kg, bool kb, const float *__restrict__ wb)
{ // :216-219 / :448-452: every lane divided by its normalization, clamped at zero; the flags as floats
  const float norm = sqrtf(sqf(R) + sqf(G) + sqf(B));
  interpolated[p] = make_float4(fmaxf(R / __ldg(wb + 0), 0.f), fmaxf(G / __ldg(wb + 1), 0.f), fmaxf(B / __ldg(wb + 2), 0.f), fmaxf(norm / __ldg(wb + 3), 0.f));
  mask[p] = make_float4(kr ? 1.f : 0.f, kg ? 1.f : 0.f, kb ? 1.f : 0.f, (kr || kg || kb) ? 1.f : 0.f);
}
// grid = (ceil(width / NT), height)
__global__ void __launch_bounds__(NT) hl_gather_bayer_kernel(const float *__restrict__ in, float4 *__restrict__ interpolated, float4 *__restrict__ mask, int width,
                                                             int height, uint32_t filters, hl_clips_t clips, const float *__restrict__ wb)
{ // _interpolate_and_mask, gather.c:67-221 (det_scale = 1): the border ring mirrors, which keeps every neighbour's colour
  const int j = blockIdx.x * NT + threadIdx.x, i = blockIdx.y;
  if(j >= width) return;
  const int c = fc(i, j, filters);
  const size_t ic = (size_t)i * width, ip = (size_t)(i == 0 ? 1 : i - 1) * width, in_ = (size_t)(i == height - 1 ? height - 2 : i + 1) * width;
  const int jp = j == 0 ? 1 : j - 1, jn = j == width - 1 ? width - 2 : j + 1;
  const float center = __ldg(in + ic + j);
  float R, G, B;
  bool kr, kg, kb;
  if(c == 1)
  {
    G = center;
    kg = center > clips.v[1];
  }
  else
  {
    const float north = __ldg(in + ip + j), south = __ldg(in + in_ + j), west = __ldg(in + ic + jp), east = __ldg(in + ic + jn);
    G = (north + south + east + west) / 4.f;
    kg = north > clips.v[1] || south > clips.v[1] || east > clips.v[1] || west > clips.v[1];
  }
  if(c == 0)
  {
    R = center;
    kr = center > clips.v[0];
  }
  else
    hl_around(in, ic, ip, in_, j, jp, jn, i, filters, 0, clips.v[0], R, kr);
  if(c == 2)
  {
    B = center;
    kb = center > clips.v[2];
  }
  else
    hl_around(in, ic, ip, in_, j, jp, jn, i, filters, 2, clips.v[2], B, kb);
  hl_store_gathered(interpolated, mask, ic + j, R, G, B, kr, kg, kb, wb);
}
__global__ void __launch_bounds__(NT) hl_gather_xtrans_kernel(const float *__restrict__ in, float4 *__restrict__ interpolated, float4 *__restrict__ mask, int width,
                                                              int height, hl_xtrans_t xt, hl_clips_t clips, const float *__restrict__ wb)
{ // _interpolate_and_mask_xtrans, gather.c:317-422.  The interior walks what _build_xtrans_bilinear_lookup :277-315 lists for the site's
  // position in the 6x6 tile: the eight neighbours row by row, those of the site's own colour skipped, weight 2 on the cross and 1 on the
  // corners; the border ring takes plain means of whatever its 3x3 window holds inside the frame, the site included.
  const int j = blockIdx.x * NT + threadIdx.x, i = blockIdx.y;
  if(j >= width) return;
  const size_t idx = (size_t)i * width + j;
  const float center = __ldg(in + idx);
  const int f = fcx(i, j, xt);
  float sum[3] = { 0.f, 0.f, 0.f }, rgb[3];
  bool used[3] = { false, false, false }, flags[3];
  if(i == 0 || j == 0 || i == height - 1 || j == width - 1)
  {
    int count[3] = { 0, 0, 0 };
    for(int y = max(i - 1, 0); y <= min(i + 1, height - 1); y++)
      for(int x = max(j - 1, 0); x <= min(j + 1, width - 1); x++)
      {
        const int color = fcx(y, x, xt);
        const float value = __ldg(in + (size_t)y * width + x);
#pragma unroll
        for(int c = 0; c < 3; c++)
          if(c == color)
          {
            sum[c] += value;
            count[c]++;
            used[c] = used[c] || value > clips.v[c];
          }
      }
#pragma unroll
    for(int c = 0; c < 3; c++)
    {
      const bool own = c == f || count[c] == 0;
      rgb[c] = own ? center : sum[c] / (float)count[c];
      flags[c] = own ? center > clips.v[c] : used[c];
    }
  }
  else
  {
    int total[3] = { 0, 0, 0 };
#pragma unroll
    for(int y = -1; y <= 1; y++)
#pragma unroll
      for(int x = -1; x <= 1; x++)
      {
        const int color = fcx(i + y, j + x, xt);
        if(color == f) continue;
        const int weight = 1 << ((y == 0) + (x == 0));
        const float value = __ldg(in + (size_t)(i + y) * width + (j + x));
#pragma unroll
        for(int c = 0; c < 3; c++)
          if(c == color)
          {
            sum[c] += value * (float)weight;
            total[c] += weight;
            used[c] = used[c] || value > clips.v[c];
          }
      }
#pragma unroll
    for(int c = 0; c < 3; c++)
    {
      rgb[c] = c == f ? center : (total[c] > 0 ? sum[c] / (float)total[c] : center);
      flags[c] = c == f ? center > clips.v[c] : used[c];
    }
  }
  hl_store_gathered(interpolated, mask, idx, rgb[0], rgb[1], rgb[2], flags[0], flags[1], flags[2], wb);
}
__global__ void __launch_bounds__(NT) hl_gather_rgba_kernel(const float4 *__restrict__ in, float4 *__restrict__ interpolated, float4 *__restrict__ mask, int width,
                                                            hl_clips_t clips, const float *__restrict__ wb)
{ // _interpolate_and_mask_passthrough, gather.c:424-455
  const int j = blockIdx.x * NT + threadIdx.x;
  if(j >= width) return;
  const size_t p = (size_t)blockIdx.y * width + j;
  const float4 v = __ldg(in + p);
  hl_store_gathered(interpolated, mask, p, v.x, v.y, v.z, v.x > clips.v[0], v.y > clips.v[1], v.z > clips.v[2], wb);
}

// ---- box mean, radius 2, dt_box_mean_4ch :950-971 -------------------------------------------------------------------------------
// Row pass (blur_horizontal_4ch :351-404): the running sums of the reference add and remove zeros and ones, which is exact in any
// order, so a pixel's sum is the count of set flags in its window and `hits` the number of columns of the window inside the row.
__global__ void __launch_bounds__(NT) hl_box_rows_kernel(const float4 *__restrict__ flags, float4 *__restrict__ out, int width, int radius)
{
  const int j = blockIdx.x * NT + threadIdx.x;
  if(j >= width) return;
  const size_t row = (size_t)blockIdx.y * width;
  const int lo = max(j - radius, 0), hi = min(j + radius, width - 1);
  float4 L = make_float4(0.f, 0.f, 0.f, 0.f);
  for(int x = lo; x <= hi; x++)
  {
    const float4 v = __ldg(flags + row + x);
    L.x += v.x;
    L.y += v.y;
    L.z += v.z;
    L.w += v.w;
  }
  const float hits = (float)(hi - lo + 1);
  out[row + j] = make_float4(L.x / hits, L.y / hits, L.z / hits, L.w / hits);
}
// Column pass (blur_vertical_1ch :891-913 and its 16-, 4- and 1-wide bodies): fractions enter and leave a float running sum, so
// the order is the reference's: down the column, the leaving sample before the entering one.  One thread per float column; the
// entering samples of eight rows are fetched ahead of the recurrence, the leaving ones are the entering ones of 2 * RADIUS + 1 rows
// earlier and wait in registers.  Frames are at least 8 rows high (the caller refuses smaller ones).
template <int RADIUS>
__global__ void __launch_bounds__(NT) hl_box_columns_kernel(const float *__restrict__ in, float *__restrict__ out, int height, size_t stride)
{
  const size_t x = (size_t)blockIdx.x * NT + threadIdx.x;
  if(x >= stride) return;
  constexpr int W = 2 * RADIUS + 1;
  const float *col = in + x;
  float *dst = out + x;
  float ring[W]; // before the step of row y: the samples of rows y - RADIUS - 1 .. y + RADIUS - 1, zeros outside the frame
#pragma unroll
  for(int q = 0; q < W; q++) ring[q] = 0.f;
  float L = 0.0f;
  int hits = 0;
#pragma unroll
  for(int y = 0; y < RADIUS; y++, hits++)
  {
    const float v = __ldg(col + (size_t)y * stride);
    L += v;
#pragma unroll
    for(int q = 0; q < W - 1; q++) ring[q] = ring[q + 1];
    ring[W - 1] = v;
  }
  for(int base = 0; base < height; base += 8)
  {
    float enter[8];
#pragma unroll
    for(int m = 0; m < 8; m++) enter[m] = base + m + RADIUS < height ? __ldg(col + (size_t)(base + m + RADIUS) * stride) : 0.f;
#pragma unroll
    for(int m = 0; m < 8; m++)
    {
      const int y = base + m;
      if(y < height)
      {
        if(y > RADIUS) L -= ring[0];
        if(y > RADIUS && y + RADIUS >= height) hits--;
        if(y + RADIUS < height)
        {
          L += enter[m];
          if(y <= RADIUS) hits++;
        }
        dst[(size_t)y * stride] = L / (float)hits;
      }
#pragma unroll
      for(int q = 0; q < W - 1; q++) ring[q] = ring[q + 1];
      ring[W - 1] = enter[m];
    }
  }
}

// ---- interpolate_bilinear, fast_guided_filter.h:99-152, four channels; grid = (ceil(width_out / NT), height_out) ------------------------
__global__ void __launch_bounds__(NT) hl_bilinear_kernel(const float4 *__restrict__ in, int width_in, int height_in, float4 *__restrict__ out, int width_out,
                                                         int height_out)
{
  const int j = blockIdx.x * NT + threadIdx.x, i = blockIdx.y;
  if(j >= width_out) return;
  const float x_out = (float)j / (float)width_out, y_out = (float)i / (float)height_out;
  const float x_in = x_out * (float)width_in, y_in = y_out * (float)height_in;
  const int x_floor = (int)floorf(x_in), y_floor = (int)floorf(y_in);
  const int x_prev = min(x_floor, width_in - 1), x_next = min(x_floor + 1, width_in - 1);
  const int y_prev = min(y_floor, height_in - 1), y_next = min(y_floor + 1, height_in - 1);
  const float4 nw = __ldg(in + (size_t)y_prev * width_in + x_prev), ne = __ldg(in + (size_t)y_prev * width_in + x_next);
  const float4 se = __ldg(in + (size_t)y_next * width_in + x_next), sw = __ldg(in + (size_t)y_next * width_in + x_prev);
  const float Dy_next = (float)y_next - y_in, Dy_prev = 1.f - Dy_next;
  const float Dx_next = (float)x_next - x_in, Dx_prev = 1.f - Dx_next;
  float4 o;
  o.x = Dy_prev * (sw.x * Dx_next + se.x * Dx_prev) + Dy_next * (nw.x * Dx_next + ne.x * Dx_prev);
  o.y = Dy_prev * (sw.y * Dx_next + se.y * Dx_prev) + Dy_next * (nw.y * Dx_next + ne.y * Dx_prev);
  o.z = Dy_prev * (sw.z * Dx_next + se.z * Dx_prev) + Dy_next * (nw.z * Dx_next + ne.z * Dx_prev);
  o.w = Dy_prev * (sw.w * Dx_next + se.w * Dx_prev) + Dy_next * (nw.w * Dx_next + ne.w * Dx_prev);
  out[(size_t)i * width_out + j] = o;
}

// ---- the two reconstructions of one wavelet scale; grid = (ceil(width / NT), height) --------------------------------------------------
struct hl_lane4
{
  float v[4];
  __device__ __forceinline__ hl_lane4() {}
  __device__ __forceinline__ hl_lane4(float4 p) : v{ p.x, p.y, p.z, p.w } {}
  __device__ __forceinline__ float4 pack() const { return make_float4(v[0], v[1], v[2], v[3]); }
};
// poisson_noise_simd, noise_generator.h:174-200, for the three colour lanes (the caller overwrites the fourth)
__device__ __forceinline__ void hl_poisson3(const f32m::tables_t &tb, const float mu[3], const float sigma[3], uint32_t (&st)[4], float out[3])
{
  float u1[3], u2[3];
  for(int c = 0; c < 3; c++)
  {
    u1[c] = fmaxf(xoshiro128plus(st), 1.17549435e-38f);
    u2[c] = xoshiro128plus(st);
  }
  for(int c = 0; c < 3; c++)
  {
    const float radius = sqrtf(-2.0f * f32m::logf_(tb, u1[c]));
    const float angle = (float)(6.283185307179586 * (double)u2[c]); // 2.f * M_PI * u2 is a double product in the source
    const float noise = (c != 1) ? radius * f32m::cosf_(angle) : radius * f32m::sinf_(angle); // flip = { 1, 0, 1, 0 }
    const float r = noise * sigma[c] + 2.0f * sqrtf(fmaxf(mu[c] + 3.f / 8.f, 0.0f));
    out[c] = (r * r - sigma[c] * sigma[c]) / 4.f - 3.f / 8.f;
  }
}
struct hl_guide_t
{
  int width, height, mult, scale, salt;
  float noise_level, scale_multiplier; // 1 / radius^2 of the scale
};
__global__ void __launch_bounds__(NT) hl_guide_kernel(const float4 *__restrict__ HF, const float4 *__restrict__ LF, const float4 *__restrict__ mask, float4 *out,
                                                      hl_guide_t a)
{ // guide_laplacians, laplacian.c:85-246
  const int j = blockIdx.x * NT + threadIdx.x, i = blockIdx.y;
  if(j >= a.width) return;
  const size_t index = (size_t)i * a.width + j;
  const hl_lane4 m(__ldg(mask + index));
  const float alpha = m.v[3], alpha_comp = 1.f - alpha;
  hl_lane4 hf(__ldg(HF + index));
  if(alpha > 0.f)
  { // a linear fit of every channel on the channel of the largest variance over the dilated 3x3 patch
    const size_t rows[3] = { (size_t)max(i - a.mult, 0) * a.width, (size_t)i * a.width, (size_t)min(i + a.mult, a.height - 1) * a.width };
    const int cols[3] = { max(j - a.mult, 0), j, min(j + a.mult, a.width - 1) };
    float sum[4] = { 0.f, 0.f, 0.f, 0.f }, sum_sq[4] = { 0.f, 0.f, 0.f, 0.f }, prod[3][4] = { { 0.f, 0.f, 0.f, 0.f }, { 0.f, 0.f, 0.f, 0.f }, { 0.f, 0.f, 0.f, 0.f } };
#pragma unroll
    for(int jj = 0; jj < 3; jj++)
#pragma unroll
      for(int ii = 0; ii < 3; ii++)
      {
        const hl_lane4 s(__ldg(HF + rows[jj] + cols[ii]));
#pragma unroll
        for(int c = 0; c < 4; c++)
        {
          sum[c] += s.v[c];
          sum_sq[c] += s.v[c] * s.v[c];
#pragma unroll
          for(int g = 0; g < 3; g++) prod[g][c] += s.v[c] * s.v[g];
        }
      }
    const float inv_patch = 1.f / 9.f;
    float means[4], variance[4];
#pragma unroll
    for(int c = 0; c < 4; c++)
    {
      means[c] = sum[c] * inv_patch;
      variance[c] = max_zero(sum_sq[c] * inv_patch - means[c] * means[c]);
    }
    int g = 0;
    float guide_variance = variance[0];
    if(variance[1] > guide_variance)
    {
      g = 1;
      guide_variance = variance[1];
    }
    if(variance[2] > guide_variance)
    {
      g = 2;
      guide_variance = variance[2];
    }
    if(guide_variance > 1e-12f)
    {
      const float guide_mean = g == 0 ? means[0] : (g == 1 ? means[1] : means[2]);
      const float guide = g == 0 ? hf.v[0] : (g == 1 ? hf.v[1] : hf.v[2]);
#pragma unroll
      for(int c = 0; c < 4; c++)
      {
        const float pr = g == 0 ? prod[0][c] : (g == 1 ? prod[1][c] : prod[2][c]);
        const float covariance = pr * inv_patch - means[c] * guide_mean;
        const float slope = max_zero(covariance / guide_variance);
        const float intercept = means[c] - slope * guide_mean;
        const float blend = m.v[c] * a.scale_multiplier;
        hf.v[c] = blend * (slope * guide + intercept) + (1.f - blend) * hf.v[c];
      }
    }
  }
  hl_lane4 px;
  if(a.scale & FIRST_SCALE)
    px = hf;
  else
  {
    const hl_lane4 o(out[index]);
#pragma unroll
    for(int c = 0; c < 4; c++) px.v[c] = hf.v[c] + o.v[c];
  }
  if(a.scale & LAST_SCALE)
  {
    const hl_lane4 lf(__ldg(LF + index));
#pragma unroll
    for(int c = 0; c < 4; c++) px.v[c] = max_zero(px.v[c] + lf.v[c]);
    if(a.salt && alpha > 0.f)
    { // noise on the last iteration, seeded by the position :198-227
      const f32m::tables_t tb = f32m::global_tables();
      uint32_t st[4] = { splitmix32((uint64_t)(j + 1)), splitmix32((uint64_t)((j + 1) * (i + 3))), splitmix32(1337), splitmix32(666) };
      xoshiro128plus(st);
      xoshiro128plus(st);
      xoshiro128plus(st);
      xoshiro128plus(st);
      const float sigma[3] = { px.v[0] * a.noise_level, px.v[1] * a.noise_level, px.v[2] * a.noise_level };
      float noise[3];
      hl_poisson3(tb, px.v, sigma, st, noise);
#pragma unroll
      for(int c = 0; c < 3; c++)
      {
        const float noisy = px.v[c] + fabsf(noise[c] - px.v[c]);
        px.v[c] = fmaxf(alpha * noisy + alpha_comp * px.v[c], 0.f);
      }
    }
    // ratios and their norm for the second reconstruction :229-235
    const float norm = fmaxf(sqrtf(sqf(px.v[0]) + sqf(px.v[1]) + sqf(px.v[2])), 1e-6f);
#pragma unroll
    for(int c = 0; c < 3; c++) px.v[c] /= norm;
    px.v[3] = norm;
  }
  out[index] = px.pack();
}
struct hl_heat_t
{
  int width, height, mult, scale;
  float first_order_factor;
};
__global__ void __launch_bounds__(NT) hl_heat_kernel(const float4 *__restrict__ HF, const float4 *__restrict__ LF, const float4 *__restrict__ mask, float4 *out,
                                                     hl_heat_t a)
{ // heat_PDE_diffusion, laplacian.c:248-372
  const int j = blockIdx.x * NT + threadIdx.x, i = blockIdx.y;
  if(j >= a.width) return;
  const size_t index = (size_t)i * a.width + j;
  const hl_lane4 alpha(__ldg(mask + index));
  hl_lane4 hf(__ldg(HF + index));
  const float norm_backup = hf.v[3];
  if(alpha.v[3] > 0.f)
  {
    const size_t rows[3] = { (size_t)max(i - a.mult, 0) * a.width, (size_t)i * a.width, (size_t)min(i + a.mult, a.height - 1) * a.width };
    const int cols[3] = { max(j - a.mult, 0), j, min(j + a.mult, a.width - 1) };
    const float kernel[9] = { 0.25f, 0.5f, 0.25f, 0.5f, -3.f, 0.5f, 0.25f, 0.5f, 0.25f };
    float lap[4] = { 0.f, 0.f, 0.f, 0.f };
#pragma unroll
    for(int k = 0; k < 9; k++)
    {
      const hl_lane4 s(__ldg(HF + rows[k / 3] + cols[k % 3]));
#pragma unroll
      for(int c = 0; c < 4; c++) lap[c] += s.v[c] * kernel[k];
    }
    const float mult3 = 1.f / B_SPLINE_TO_LAPLACIAN;
#pragma unroll
    for(int c = 0; c < 3; c++) hf.v[c] += alpha.v[c] * mult3 * (lap[c] - a.first_order_factor * hf.v[c]);
    hf.v[3] = norm_backup; // the norm is not diffused :334
  }
  hl_lane4 px;
  if(a.scale & FIRST_SCALE)
    px = hf;
  else
  {
    const hl_lane4 o(out[index]);
#pragma unroll
    for(int c = 0; c < 4; c++) px.v[c] = o.v[c] + hf.v[c];
  }
  if(a.scale & LAST_SCALE)
  {
    const hl_lane4 lf(__ldg(LF + index));
#pragma unroll
    for(int c = 0; c < 4; c++) px.v[c] = fmaxf(px.v[c] + lf.v[c], 0.f);
    if(alpha.v[3] > 0.f)
    {
      const float norm = sqrtf(sqf(px.v[0]) + sqf(px.v[1]) + sqf(px.v[2]));
      if(norm > 1e-4f)
      {
#pragma unroll
        for(int c = 0; c < 3; c++) px.v[c] /= norm;
      }
    }
#pragma unroll
    for(int c = 0; c < 3; c++) px.v[c] = px.v[c] * px.v[3]; // back from ratios to RGB; the norm stays in the fourth lane
  }
  out[index] = px.pack();
}

// ---- composite: the reconstruction, taken back to the input's scale, under the feathered mask -------------------------------------------
__global__ void __launch_bounds__(NT) hl_remosaic_mosaic_kernel(const float *__restrict__ in, const float4 *__restrict__ interpolated, const float4 *__restrict__ mask,
                                                                float *__restrict__ out, int width, uint32_t filters, hl_xtrans_t xt, const float *__restrict__ wb)
{ // _remosaic_and_replace, gather.c:457-486, and _remosaic_and_replace_xtrans :488-512, clip_is_floor = FALSE
  const int j = blockIdx.x * NT + threadIdx.x, i = blockIdx.y;
  if(j >= width) return;
  const size_t p = (size_t)i * width + j;
  const int c = hl_colour(i, j, filters, xt);
  const float4 v = __ldg(interpolated + p);
  const float opacity = __ldg(mask + p).w;
  const float rec = fmaxf((c == 0 ? v.x : (c == 1 ? v.y : v.z)) * __ldg(wb + c), 0.f);
  out[p] = opacity * rec + (1.f - opacity) * __ldg(in + p);
}
__global__ void __launch_bounds__(NT) hl_remosaic_rgba_kernel(const float4 *__restrict__ in, const float4 *__restrict__ interpolated, const float4 *__restrict__ mask,
                                                              float4 *__restrict__ out, int width, const float *__restrict__ wb)
{ // _remosaic_and_replace_passthrough, gather.c:514-541: each channel under its own mask, the fourth lane passed through
  const int j = blockIdx.x * NT + threadIdx.x;
  if(j >= width) return;
  const size_t p = (size_t)blockIdx.y * width + j;
  const float4 v = __ldg(interpolated + p), m = __ldg(mask + p), b = __ldg(in + p);
  float4 o;
  o.x = m.x * fmaxf(v.x * __ldg(wb + 0), 0.f) + (1.f - m.x) * b.x;
  o.y = m.y * fmaxf(v.y * __ldg(wb + 1), 0.f) + (1.f - m.y) * b.y;
  o.z = m.z * fmaxf(v.z * __ldg(wb + 2), 0.f) + (1.f - m.z) * b.z;
  o.w = b.w;
  out[p] = o;
}

// ---- the launch sequence ------------------------------------------------------------------------------------------------------------
#ifdef B200_KERNELS_ON_CPU
#define HL_LAUNCH(kernel, grid, ...) emulate(grid, NT, kernel, __VA_ARGS__)
#else
#define HL_LAUNCH(kernel, grid, ...)                                                                                                                           \
  do                                                                                                                                                           \
  {                                                                                                                                                            \
    kernel<<<grid, NT, 0, s>>>(__VA_ARGS__);                                                                                                                   \
    B200_CUDA_TRY(cudaGetLastError());                                                                                                                         \
  } while(0)
#endif

struct hl_job_t
{
  int width, height, ds_width, ds_height;
  uint32_t filters; // seen from the ROI origin; 0 = RGBA input, 9 = X-Trans (xt)
  hl_xtrans_t xt;
  hl_clips_t clips;
  int iterations, scales;
  float noise_level, solid_color;
};
struct hl_buffers_t
{
  float4 *interpolated, *mask_a, *mask_b;                                  // full size
  float4 *LF_odd, *LF_even, *temp, *HF, *ds_interpolated, *ds_mask, *vtmp; // a sixteenth of it
  float *norm;                                                             // 4 floats
};

float hl_sigma_at_step(unsigned s)
{ // equivalent_sigma_at_step, bspline.h:52-63
  if(s == 0) return B_SPLINE_SIGMA;
  const float prev = hl_sigma_at_step(s - 1), e = exp2f((float)s) * B_SPLINE_SIGMA;
  return sqrtf(prev * prev + e * e);
}
int hl_scales(int scales_param, float module_scale)
{ // laplacian.c:461-463
  const float scale = DS_FACTOR * module_scale;
  const float final_radius = (float)((int)(1 << scales_param)) / scale;
  const int n = (int)ceilf(log2f(final_radius));
  return n < 1 ? 1 : (n > MAX_NUM_SCALES ? MAX_NUM_SCALES : n);
}

// wavelets_process, laplacian.c:374-430
int hl_wavelets(const hl_job_t &J, const hl_buffers_t &B, const float4 *in, float4 *reconstructed, bool chroma, int salt, cudaStream_t s)
{
  (void)s;
  const dim3 grid((unsigned)((J.ds_width + NT - 1) / NT), (unsigned)J.ds_height);
  for(int k = 0; k < J.scales; k++)
  {
    const float4 *buffer_in = k == 0 ? in : ((k & 1) ? B.LF_odd : B.LF_even);
    float4 *buffer_out = k == 0 ? B.LF_odd : ((k & 1) ? B.LF_even : B.LF_odd);
    const int mult = 1 << k;
    HL_LAUNCH(bspline_vertical_kernel, grid, buffer_in, B.vtmp, J.ds_width, J.ds_height, mult);
    HL_LAUNCH(bspline_horizontal_kernel, grid, (const float4 *)B.vtmp, buffer_in, buffer_out, B.HF, (float4 *)nullptr, J.ds_width, mult);
    const int type = 1 | (k == 0 ? FIRST_SCALE : 0) | (k == J.scales - 1 ? LAST_SCALE : 0);
    if(!chroma)
    {
      const float sigma = hl_sigma_at_step((unsigned)(k * DS_FACTOR));
      const hl_guide_t a = { J.ds_width, J.ds_height, mult, type, salt, J.noise_level, 1.f / (sigma * sigma) };
      HL_LAUNCH(hl_guide_kernel, grid, (const float4 *)B.HF, (const float4 *)buffer_out, (const float4 *)B.ds_mask, reconstructed, a);
    }
    else
    {
      const hl_heat_t a = { J.ds_width, J.ds_height, mult, type, J.solid_color };
      HL_LAUNCH(hl_heat_kernel, grid, (const float4 *)B.HF, (const float4 *)buffer_out, (const float4 *)B.ds_mask, reconstructed, a);
    }
  }
  return 0;
}
// process_laplacian :433-575 past the normalization (B.norm holds it)
int hl_sequence(const hl_job_t &J, const hl_buffers_t &B, const void *d_in, void *d_out, cudaStream_t s)
{
  (void)s;
  const dim3 full((unsigned)((J.width + NT - 1) / NT), (unsigned)J.height), ds((unsigned)((J.ds_width + NT - 1) / NT), (unsigned)J.ds_height);
  if(J.filters == 9u)
    HL_LAUNCH(hl_gather_xtrans_kernel, full, (const float *)d_in, B.interpolated, B.mask_a, J.width, J.height, J.xt, J.clips, (const float *)B.norm);
  else if(J.filters)
    HL_LAUNCH(hl_gather_bayer_kernel, full, (const float *)d_in, B.interpolated, B.mask_a, J.width, J.height, J.filters, J.clips, (const float *)B.norm);
  else
    HL_LAUNCH(hl_gather_rgba_kernel, full, (const float4 *)d_in, B.interpolated, B.mask_a, J.width, J.clips, (const float *)B.norm);
  HL_LAUNCH(hl_box_rows_kernel, full, (const float4 *)B.mask_a, B.mask_b, J.width, 2);
  const size_t stride = (size_t)4 * J.width;
  HL_LAUNCH(hl_box_columns_kernel<2>, dim3((unsigned)((stride + NT - 1) / NT)), (const float *)B.mask_b, (float *)B.mask_a, J.height, stride);
  HL_LAUNCH(hl_bilinear_kernel, ds, (const float4 *)B.mask_a, J.width, J.height, B.ds_mask, J.ds_width, J.ds_height);
  HL_LAUNCH(hl_bilinear_kernel, ds, (const float4 *)B.interpolated, J.width, J.height, B.ds_interpolated, J.ds_width, J.ds_height);
  for(int i = 0; i < J.iterations; i++)
  {
    const int salt = i == J.iterations - 1; // noise on the last iteration only
    int rc = hl_wavelets(J, B, B.ds_interpolated, B.temp, false, salt, s);
    if(rc) return rc;
    if((rc = hl_wavelets(J, B, B.temp, B.ds_interpolated, true, salt, s))) return rc;
  }
  HL_LAUNCH(hl_bilinear_kernel, full, (const float4 *)B.ds_interpolated, J.ds_width, J.ds_height, B.interpolated, J.width, J.height);
  if(J.filters)
    HL_LAUNCH(hl_remosaic_mosaic_kernel, full, (const float *)d_in, (const float4 *)B.interpolated, (const float4 *)B.mask_a, (float *)d_out, J.width, J.filters,
              J.xt, (const float *)B.norm);
  else
    HL_LAUNCH(hl_remosaic_rgba_kernel, full, (const float4 *)d_in, (const float4 *)B.interpolated, (const float4 *)B.mask_a, (float4 *)d_out, J.width,
              (const float *)B.norm);
  return 0;
}
} // namespace

#ifndef B200_KERNELS_ON_CPU
namespace b200
{
// clips: what process() hands to process_laplacian (iop/highlights.c:764-766); normalization: NULL = computed here
int highlights_laplacian_dev(const b200_piece_t *piece, const b200_highlights_data_t *d, const void *d_in, void *d_out, const float clips[4],
                             const float *normalization, cudaStream_t s)
{
  const int width = piece->roi_in.width, height = piece->roi_in.height;
  if(!piece->filters && piece->channels != 4) return fail(B200_ERR_ARG, "highlights: guided laplacians on %d-channel input", piece->channels);
  if(width < 2 * DS_FACTOR || height < 2 * DS_FACTOR)
    return fail(B200_ERR_UNSUPPORTED, "highlights: guided laplacians on a %dx%d frame (the reference's quarter-size planes need 8 px each way)", width, height);
  if(d->iterations < 1 || d->scales < 0 || d->scales > 30) return fail(B200_ERR_ARG, "highlights: iterations %d, scales %d", d->iterations, d->scales);
  hl_job_t J;
  J.width = width;
  J.height = height;
  J.ds_width = width / DS_FACTOR;
  J.ds_height = height / DS_FACTOR;
  J.filters = b200_roi_filters(piece->filters, piece->roi_in.x, piece->roi_in.y);
  memset(&J.xt, 0, sizeof(J.xt));
  if(piece->filters == 9u)
    for(int r = 0; r < 6; r++)
      for(int c = 0; c < 6; c++)
      {
        const unsigned char v = piece->xtrans[(r + piece->roi_in.y + 600) % 6][(c + piece->roi_in.x + 600) % 6];
        if(v > 2) return fail(B200_ERR_ARG, "highlights: xtrans table holds %d", (int)v);
        J.xt.v[r][c] = v;
      }
  for(int c = 0; c < 4; c++) J.clips.v[c] = clips[c];
  const float module_scale = (float)((double)piece->iscale / piece->roi_in.scale); // dt_dev_get_module_scale, develop/imageop.c:134-137: a double quotient
  J.iterations = d->iterations;
  J.scales = hl_scales(d->scales, module_scale);
  J.noise_level = d->noise_level / (DS_FACTOR * module_scale);
  J.solid_color = d->solid_color;

  const size_t npx = (size_t)width * height, ds_npx = (size_t)J.ds_width * J.ds_height;
  void *full[3], *quarter = nullptr, *small = nullptr;
  int rc;
  for(int k = 0; k < 3; k++)
    if((rc = scratch(SLOT_TMP0 + k, npx * 16, &full[k]))) return rc;
  if((rc = scratch(SLOT_TMP3, ds_npx * 16 * 7, &quarter))) return rc;
  if((rc = scratch(SLOT_SMALL, 256 + sizeof(double) * 3 * NORM_BLOCKS, &small))) return rc;
  hl_buffers_t B;
  B.interpolated = (float4 *)full[0];
  B.mask_a = (float4 *)full[1];
  B.mask_b = (float4 *)full[2];
  float4 *q = (float4 *)quarter;
  B.LF_odd = q;
  B.LF_even = q + ds_npx;
  B.temp = q + 2 * ds_npx;
  B.HF = q + 3 * ds_npx;
  B.ds_interpolated = q + 4 * ds_npx;
  B.ds_mask = q + 5 * ds_npx;
  B.vtmp = q + 6 * ds_npx;
  B.norm = (float *)small;
  if(normalization)
    B200_CUDA_TRY(cudaMemcpyAsync(B.norm, normalization, 16, cudaMemcpyHostToDevice, s));
  else
  {
    double *partial = (double *)((char *)small + 256);
    const int blocks = height < NORM_BLOCKS ? height : NORM_BLOCKS;
    hl_norm_partial_kernel<<<blocks, NT, 0, s>>>((const float *)d_in, width, height, J.filters, J.xt, partial);
    B200_CUDA_TRY(cudaGetLastError());
    hl_norm_final_kernel<<<1, NT, 0, s>>>(partial, blocks, (float)(height * width), B.norm);
    B200_CUDA_TRY(cudaGetLastError());
  }
  return hl_sequence(J, B, d_in, d_out, s);
}
} // namespace b200
#endif
