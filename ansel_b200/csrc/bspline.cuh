// The a-trous B-spline decomposition (src/pixel/bspline.h decompose_2D_Bspline :351-377 with _bspline_vertical_pass :118-133 and
// _bspline_horizontal :136-151, clip at zero after EACH pass) and the counter-based noise source (src/iop/noise_generator.h splitmix32
// :36-43, xoshiro128plus :54-70) shared by diffuse.cu and highlights_laplacian.cu.  The kernels have internal linkage: each
// translation unit that includes this file launches its own copy.
#pragma once
#include <stdint.h>

namespace
{
namespace bsp
{
constexpr int NT = 256;
__device__ __forceinline__ float clip0(float v) { return 0.0f > v ? 0.0f : v; } // MAX(0.0f, v)
__device__ __forceinline__ float max_zero(float v)
{ // dt_simd_max_zero, system/simd.h:107-114
  const float t = fmaxf(v, 0.0f); // NaN -> 0, -inf -> 0, -0 -> +0 (PTX max), finite -> MAX(v, 0)
  return t == __int_as_float(0x7f800000) ? 0.0f : t;
}
// The two B-spline passes are pure streaming: one thread per PIXEL (float4), five 16-byte taps; grid = (ceil(width / NT), height).
__device__ __forceinline__ float bs5(float a, float b, float c, float d, float e)
{ // sparse_scalar_product(), bspline.h:83-117, clip_negatives
  return clip0(0.0625f * a + 0.25f * b + 0.375f * c + 0.25f * d + 0.0625f * e);
}
// _bspline_vertical_pass: rows clamped, clip at zero.  grid.y = row, x over the pixels of a row
__global__ void __launch_bounds__(NT) bspline_vertical_kernel(const float4 *__restrict__ in, float4 *__restrict__ tmp, int width, int height, int mult)
{
  const int j = blockIdx.x * NT + threadIdx.x, i = blockIdx.y;
  if(j >= width) return;
  const float4 *b = in + j;
  const float4 p0 = __ldg(b + (size_t)width * max(i - 2 * mult, 0)), p1 = __ldg(b + (size_t)width * max(i - mult, 0));
  const float4 p2 = __ldg(b + (size_t)width * i);
  const float4 p3 = __ldg(b + (size_t)width * min(i + mult, height - 1)), p4 = __ldg(b + (size_t)width * min(i + 2 * mult, height - 1));
  tmp[(size_t)width * i + j] = make_float4(bs5(p0.x, p1.x, p2.x, p3.x, p4.x), bs5(p0.y, p1.y, p2.y, p3.y, p4.y), bs5(p0.z, p1.z, p2.z, p3.z, p4.z),
                                           bs5(p0.w, p1.w, p2.w, p3.w, p4.w));
}
// _bspline_horizontal + the HF subtraction of decompose_2D_Bspline
__device__ __forceinline__ float ratio_sq(float hf, float lf)
{ // one term of the HF/LF energy of heat_PDE_diffusion(), diffuse.c:826-830
  const float safe_lf = max_zero(lf - 1e-8f) + 1e-8f;
  const float ratio = hf / safe_lf;
  return ratio * ratio;
}
// R (optional): the energy terms of the coarsest band, whose LF is this very output
__global__ void __launch_bounds__(NT) bspline_horizontal_kernel(const float4 *__restrict__ tmp, const float4 *__restrict__ in, float4 *__restrict__ LF,
                                                                float4 *__restrict__ HF, float4 *__restrict__ R, int width, int mult)
{
  const int j = blockIdx.x * NT + threadIdx.x;
  if(j >= width) return;
  const size_t row = (size_t)width * blockIdx.y;
  const float4 *t = tmp + row;
  const float4 p0 = __ldg(t + max(j - 2 * mult, 0)), p1 = __ldg(t + max(j - mult, 0)), p2 = __ldg(t + j);
  const float4 p3 = __ldg(t + min(j + mult, width - 1)), p4 = __ldg(t + min(j + 2 * mult, width - 1));
  const float4 lf = make_float4(bs5(p0.x, p1.x, p2.x, p3.x, p4.x), bs5(p0.y, p1.y, p2.y, p3.y, p4.y), bs5(p0.z, p1.z, p2.z, p3.z, p4.z),
                                bs5(p0.w, p1.w, p2.w, p3.w, p4.w));
  const float4 v = __ldg(in + row + j);
  const float4 hf = make_float4(v.x - lf.x, v.y - lf.y, v.z - lf.z, v.w - lf.w);
  LF[row + j] = lf;
  HF[row + j] = hf;
  if(R) R[row + j] = make_float4(ratio_sq(hf.x, lf.x), ratio_sq(hf.y, lf.y), ratio_sq(hf.z, lf.z), ratio_sq(hf.w, lf.w));
}

__device__ __forceinline__ uint32_t splitmix32(uint64_t seed)
{
  uint64_t r = (seed ^ (seed >> 33)) * 0x62a9d9ed799705f5ull;
  r = (r ^ (r >> 28)) * 0xcb24d0a5c88c35b3ull;
  return (uint32_t)(r >> 32);
}
__device__ __forceinline__ float xoshiro128plus(uint32_t (&st)[4])
{
  const uint32_t result = st[0] + st[3];
  const uint32_t t = st[1] << 9;
  st[2] ^= st[0];
  st[3] ^= st[1];
  st[1] ^= st[2];
  st[0] ^= st[3];
  st[2] ^= t;
  st[3] = (st[3] << 11) | (st[3] >> 21);
  return (float)(result >> 8) * 0x1.0p-24f;
}
} // namespace bsp
} // namespace
