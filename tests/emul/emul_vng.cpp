/* Test infrastructure: the kernels of ansel_b200/csrc/vng.cu compiled with g++ and run thread by thread on the CPU, in
 * the order vng_demosaic_dev() / dual_demosaic_dev() launch them.  The colour smoothing between VNG and the blend is a
 * kernel of demosaic_extra.cu (GPU-tested on its own); here the caller supplies it.  Not part of the product. */
#define B200_KERNELS_ON_CPU
#define __constant__ static const
#include "cuda_on_cpu.h"
#include "../../include/b200iop.h"
#include "../../ansel_b200/csrc/vng.cu"
#include <vector>

extern "C" int emul_vng_cfa(float *out, const float *in, int width, int height, int x0, int y0, unsigned filters, const unsigned char *xtrans36, int only_linear);
extern "C" int emul_vng(float *out, const float *in, int width, int height, int x0, int y0, unsigned filters, int only_linear)
{
  return emul_vng_cfa(out, in, width, height, x0, y0, filters, nullptr, only_linear);
}
extern "C" int emul_vng_cfa(float *out, const float *in, int width, int height, int x0, int y0, unsigned filters, const unsigned char *xtrans36, int only_linear)
{
  const dim3 grid((unsigned)((width + VNT - 1) / VNT), (unsigned)height);
  const cfa_t f4 = make_cfa(filters, xtrans36);
  std::vector<float4> lin((size_t)width * height);
  emulate(grid, VNT, lin_interpolate_kernel, in, only_linear ? (float4 *)out : lin.data(), width, height, x0, y0, f4);
  if(!only_linear) emulate(grid, VNT, vng_kernel, (const float4 *)lin.data(), (float4 *)out, width, height, x0, y0, f4);
  return 0;
}
typedef void (*smooth_fn)(float *, int, int, int);
extern "C" int emul_dual(float *rgb, const float *raw, int width, int height, int x0, int y0, unsigned filters, const float *wb, float dual_threshold, smooth_fn smooth)
{
  if(width < 16 || height < 16 || !(dual_threshold > 0.0f)) return 0;
  const size_t n = (size_t)width * height;
  std::vector<float> vng(4 * n), tmp(n), blend(n);
  emul_vng(vng.data(), raw, width, height, x0, y0, filters, 0);
  smooth(vng.data(), width, height, 2);
  const dim3 grid((unsigned)((width + VNT - 1) / VNT), (unsigned)height);
  emulate(dim3((unsigned)((n + 255) / 256)), 256, detail_luma_kernel, (const float4 *)rgb, tmp.data(), n, wb[0], wb[1], wb[2]);
  emulate(grid, VNT, detail_sigmoid_kernel, (const float *)tmp.data(), blend.data(), width, height, slider2contrast(dual_threshold));
  blur9_t B;
  blur9_coeff(&B, 2.0f);
  emulate(grid, VNT, detail_blur_kernel, (const float *)blend.data(), tmp.data(), width, height, B);
  emulate(dim3((unsigned)((n + 255) / 256)), 256, dual_blend_kernel, (float4 *)rgb, (const float4 *)vng.data(), (const float *)tmp.data(), n);
  return 0;
}
