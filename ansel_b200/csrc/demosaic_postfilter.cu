// The guided-Laplacian post-filter of the half-size demosaic ("downsample" method with data->color_smoothing iterations).
//
// Reference: iop/demosaic.c _downsample_guided_laplacian_fit :681-759, _apply :770-796, _postfilter :810-926 with
// DOWNSAMPLE_GUIDED_SCALES :117 = 1; pixel/bspline.h decompose_2D_Bspline :351-377, blur_2D_Bspline :330-350; dispatch :1108.
//
// One iteration, per a-trous scale: LF = clipped B-spline blur of the image, HF = (image - LF) / max(LF, 1e-8); a 5x5 patch of
// HF around every pixel gives the least-squares line channel = slope * guide + intercept over guide = (R+G+B)/3; slopes and
// intercepts are blurred (unclipped), the band (slope * guide + intercept) * LF is accumulated, and the result is
// max(bands + last LF, 0).  The reference runs ten passes over seven frame-sized buffers; here the horizontal half of every
// blur is computed by the kernel that consumes it, so an iteration is five launches and the blurred slopes / intercepts, the
// un-normalised HF and (with one scale) the band sum never reach memory:
//   vertical(image) -> decompose (LF, HF) -> fit (slope, intercept) -> vertical(slope, intercept) -> apply (image)
// about 290 bytes of traffic per half-size pixel and iteration against the reference's ~600.  Arithmetic and its order are the
// reference's (strict build), so results are bit-identical.
#ifndef B200_KERNELS_ON_CPU // tests/emul compiles the kernels of this file with g++ to check them against the oracle without a GPU
#include "runtime.h"
#endif
#include <math.h>

namespace
{
constexpr int FNT = 256;
constexpr int GUIDED_SCALES = 1; // demosaic.c:117

__device__ __forceinline__ float pf_div3(float a)
{ // a / 3.f as an IEEE division (nvcc would multiply by the rounded reciprocal)
#ifdef B200_KERNELS_ON_CPU
  return a / 3.0f;
#else
  float q;
  asm("div.rn.ftz.f32 %0, %1, %2;" : "=f"(q) : "f"(a), "f"(3.0f));
  return q;
#endif
}
__device__ __forceinline__ float pf_max_zero(float v)
{ // dt_simd_max_zero, system/simd.h:107-114: non-finite -> 0, else MAX(v, 0)
  const float t = v > 0.0f ? v : 0.0f; // NaN and -0 -> +0
  return t == __uint_as_float(0x7f800000u) ? 0.0f : t;
}
// sparse_scalar_product(), bspline.h:83-117
template <bool CLIP> __device__ __forceinline__ float pf_bs5(float a, float b, float c, float d, float e)
{
  const float v = 0.0625f * a + 0.25f * b + 0.375f * c + 0.25f * d + 0.0625f * e;
  return CLIP ? (0.0f > v ? 0.0f : v) : v;
}
template <bool CLIP> __device__ __forceinline__ float4 pf_bs5_4(const float4 &p0, const float4 &p1, const float4 &p2, const float4 &p3, const float4 &p4)
{
  return make_float4(pf_bs5<CLIP>(p0.x, p1.x, p2.x, p3.x, p4.x), pf_bs5<CLIP>(p0.y, p1.y, p2.y, p3.y, p4.y), pf_bs5<CLIP>(p0.z, p1.z, p2.z, p3.z, p4.z),
                     pf_bs5<CLIP>(p0.w, p1.w, p2.w, p3.w, p4.w));
}
template <bool CLIP> __device__ __forceinline__ float4 pf_row_blur(const float4 *__restrict__ t, int j, int width, int mult)
{ // _bspline_horizontal :136-151 on one vertically blurred row
  return pf_bs5_4<CLIP>(t[max(j - 2 * mult, 0)], t[max(j - mult, 0)], t[j], t[min(j + mult, width - 1)], t[min(j + 2 * mult, width - 1)]);
}
// _bspline_vertical_pass :118-133 for one image, or for two at once (blockIdx.z picks: the slopes and the intercepts)
template <bool CLIP>
__global__ void __launch_bounds__(FNT) pf_vertical_kernel(const float4 *__restrict__ in0, const float4 *__restrict__ in1, float4 *__restrict__ tmp0,
                                                          float4 *__restrict__ tmp1, int width, int height, int mult)
{
  const int j = blockIdx.x * FNT + threadIdx.x, i = blockIdx.y;
  if(j >= width) return;
  const float4 *b = (blockIdx.z ? in1 : in0) + j;
  float4 *o = blockIdx.z ? tmp1 : tmp0;
  o[(size_t)width * i + j] = pf_bs5_4<CLIP>(b[(size_t)width * max(i - 2 * mult, 0)], b[(size_t)width * max(i - mult, 0)], b[(size_t)width * i],
                                            b[(size_t)width * min(i + mult, height - 1)], b[(size_t)width * min(i + 2 * mult, height - 1)]);
}
// the horizontal half of decompose_2D_Bspline and the normalisation of the band, demosaic.c:868-884
__global__ void __launch_bounds__(FNT) pf_decompose_kernel(const float4 *__restrict__ tmp, const float4 *__restrict__ in, float4 *__restrict__ LF,
                                                           float4 *__restrict__ HF, int width, int mult)
{
  const int j = blockIdx.x * FNT + threadIdx.x;
  if(j >= width) return;
  const size_t row = (size_t)width * blockIdx.y;
  const float4 lf = pf_row_blur<true>(tmp + row, j, width, mult);
  const float4 v = in[row + j];
  LF[row + j] = lf;
  HF[row + j] = make_float4((v.x - lf.x) / fmaxf(lf.x, 1e-8f), (v.y - lf.y) / fmaxf(lf.y, 1e-8f), (v.z - lf.z) / fmaxf(lf.z, 1e-8f), 0.0f);
}
// _downsample_guided_laplacian_fit :681-759: the 25 taps in the reference's order (rows outer), clamped at the frame
__global__ void __launch_bounds__(FNT) pf_fit_kernel(const float4 *__restrict__ HF, float4 *__restrict__ coeff, float4 *__restrict__ bias, int width, int height)
{
  const int col = blockIdx.x * FNT + threadIdx.x, row = blockIdx.y;
  if(col >= width) return;
  float sr = 0.f, sg = 0.f, sb = 0.f, srg = 0.f, sgg = 0.f, sbg = 0.f, sum_guide = 0.f, sum_guide_sq = 0.f;
#pragma unroll
  for(int jj = -2; jj <= 2; jj++)
  {
    const float4 *r = HF + (size_t)min(max(row + jj, 0), height - 1) * width;
#pragma unroll
    for(int ii = -2; ii <= 2; ii++)
    {
      const float4 s = r[min(max(col + ii, 0), width - 1)];
      const float guide = pf_div3(s.x + s.y + s.z);
      sr += s.x, sg += s.y, sb += s.z;
      sum_guide += guide;
      sum_guide_sq += guide * guide;
      srg += s.x * guide, sgg += s.y * guide, sbg += s.z * guide;
    }
  }
  constexpr float inv_patch = 1.f / 25.f;
  const float guide_mean = sum_guide * inv_patch;
  float variance = sum_guide_sq * inv_patch - guide_mean * guide_mean;
  if(variance < 0.f) variance = 0.f;
  const bool fitted = variance > 1e-12f;
  const float mr = sr * inv_patch, mg = sg * inv_patch, mb = sb * inv_patch;
  const float cr = srg * inv_patch - mr * guide_mean, cg = sgg * inv_patch - mg * guide_mean, cb = sbg * inv_patch - mb * guide_mean;
  const float kr = fitted ? cr / variance : 0.f, kg = fitted ? cg / variance : 0.f, kb = fitted ? cb / variance : 0.f;
  const size_t p = (size_t)row * width + col;
  coeff[p] = make_float4(kr, kg, kb, 0.f);
  bias[p] = make_float4(mr - kr * guide_mean, mg - kg * guide_mean, mb - kb * guide_mean, 0.f);
}
// the horizontal half of the two blur_2D_Bspline calls (:888-891), _apply :770-796, and at the last scale the iteration's
// result :905-921 (the residual is this scale's LF).  `rec` carries the band sum between scales (unused with one scale).
template <bool LAST>
__global__ void __launch_bounds__(FNT) pf_apply_kernel(const float4 *__restrict__ tmp_coeff, const float4 *__restrict__ tmp_bias, const float4 *__restrict__ HF,
                                                       const float4 *__restrict__ LF, float4 *rec, float4 *__restrict__ out, int width, int reset)
{
  const int j = blockIdx.x * FNT + threadIdx.x;
  if(j >= width) return;
  const size_t row = (size_t)width * blockIdx.y;
  const float4 k = pf_row_blur<false>(tmp_coeff + row, j, width, 1), b = pf_row_blur<false>(tmp_bias + row, j, width, 1);
  const float4 hf = HF[row + j], lf = LF[row + j];
  const float guide = pf_div3(hf.x + hf.y + hf.z);
  float fr = (k.x * guide + b.x) * lf.x, fg = (k.y * guide + b.y) * lf.y, fb = (k.z * guide + b.z) * lf.z;
  if(!reset)
  {
    const float4 r = rec[row + j];
    fr += r.x, fg += r.y, fb += r.z;
  }
  if(LAST)
    out[row + j] = make_float4(pf_max_zero(fr + lf.x), pf_max_zero(fg + lf.y), pf_max_zero(fb + lf.z), 0.f);
  else
    rec[row + j] = make_float4(fr, fg, fb, 0.f);
}
// frame-sized RGBA buffers the filter needs besides the image: LF (two when scales alternate), HF, slopes, intercepts and their
// vertically blurred copies (the first doubles as the decomposition's row-blur buffer), the band sum between scales
constexpr int pf_buffers() { return GUIDED_SCALES > 1 ? 8 : 6; }
} // namespace

#ifndef B200_KERNELS_ON_CPU
namespace b200
{
size_t demosaic_postfilter_bytes(int width, int height) { return (size_t)pf_buffers() * width * height * sizeof(float4); }
// demosaic.c:1108: d_rgba = the half-size frame, filtered in place
int demosaic_postfilter_dev(float *d_rgba, int width, int height, int iterations, cudaStream_t s)
{
  if(iterations <= 0) return B200_OK;
  if(height > 65535) return fail(B200_ERR_ARG, "demosaic: post-filter frame height %d", height);
  void *base = nullptr;
  int rc = scratch(SLOT_TMP0, demosaic_postfilter_bytes(width, height), &base);
  if(rc) return rc;
  const size_t px = (size_t)width * height;
  float4 *buf = (float4 *)base, *out = (float4 *)d_rgba;
  float4 *LF_odd = buf, *HF = buf + px, *coeff = buf + 2 * px, *bias = buf + 3 * px, *tmp_coeff = buf + 4 * px, *tmp_bias = buf + 5 * px;
  float4 *LF_even = GUIDED_SCALES > 1 ? buf + 6 * px : nullptr, *rec = GUIDED_SCALES > 1 ? buf + 7 * px : nullptr;
  const dim3 grid((unsigned)((width + FNT - 1) / FNT), (unsigned)height), grid2(grid.x, grid.y, 2);
  for(int it = 0; it < iterations; it++)
    for(int sc = 0; sc < GUIDED_SCALES; sc++)
    {
      const float4 *bin = sc == 0 ? out : (sc % 2 ? LF_odd : LF_even);
      float4 *bout = (sc == 0 || sc % 2 == 0) ? LF_odd : LF_even;
      pf_vertical_kernel<true><<<grid, FNT, 0, s>>>(bin, bin, tmp_coeff, tmp_coeff, width, height, 1 << sc);
      pf_decompose_kernel<<<grid, FNT, 0, s>>>(tmp_coeff, bin, bout, HF, width, 1 << sc);
      pf_fit_kernel<<<grid, FNT, 0, s>>>(HF, coeff, bias, width, height);
      pf_vertical_kernel<false><<<grid2, FNT, 0, s>>>(coeff, bias, tmp_coeff, tmp_bias, width, height, 1);
      if(sc == GUIDED_SCALES - 1)
        pf_apply_kernel<true><<<grid, FNT, 0, s>>>(tmp_coeff, tmp_bias, HF, bout, rec, out, width, sc == 0);
      else
        pf_apply_kernel<false><<<grid, FNT, 0, s>>>(tmp_coeff, tmp_bias, HF, bout, rec, out, width, sc == 0);
      B200_CUDA_TRY(cudaGetLastError());
    }
  return B200_OK;
}
} // namespace b200
#endif
