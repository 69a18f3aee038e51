/* Test infrastructure: the kernels of ansel_b200/csrc/ppg.cu compiled with g++ and run thread by thread on the CPU, in
 * the order ppg_demosaic_dev() launches them.  Not part of the product. */
#define B200_KERNELS_ON_CPU
#include "cuda_on_cpu.h"
#include "../../include/b200iop.h"
#include "../../ansel_b200/csrc/ppg.cu"
#include <vector>

extern "C" int emul_demosaic_ppg(float *out, const float *in, int width, int height, unsigned filters, float median_thrs)
{
  ppg_frame_t F = { in, in, width, height, filters };
  const dim3 grid((unsigned)((width + PNT - 1) / PNT), (unsigned)height);
  std::vector<float> med;
  if(median_thrs > 0.0f)
  {
    med.resize((size_t)width * height);
    emulate(grid, PNT, pre_median_kernel, in, med.data(), width, height, filters, median_thrs);
    F.input = med.data();
  }
  emulate(grid, PNT, ppg_kernel, F, (float4 *)out);
  return 0;
}

extern "C" int emul_demosaic_passthrough(float *out, const float *in, int width, int height, int x, int y, unsigned filters, const unsigned char *xtrans36, int colour)
{
  emulate(dim3((unsigned)((width + PNT - 1) / PNT), (unsigned)height), PNT, passthrough_kernel, in, (float4 *)out, width, height, colour, filters, x, y, xtrans36);
  return 0;
}

extern "C" int emul_demosaic_downsample(float *out, const float *in, int width, int height, unsigned filters)
{
  const int ow = (width + 1) / 2, oh = (height + 1) / 2;
  emulate(dim3((unsigned)((ow + PNT - 1) / PNT), (unsigned)oh), PNT, downsample_kernel, in, (float4 *)out, width, height, ow, filters);
  return 0;
}

extern "C" int emul_demosaic_downsample_xtrans(float *out, const float *in, int width, int height, int x, int y, const unsigned char *xtrans36)
{
  const int ow = (width + 1) / 2, oh = (height + 1) / 2;
  emulate(dim3((unsigned)((ow + PNT - 1) / PNT), (unsigned)oh), PNT, downsample_xtrans_kernel, in, (float4 *)out, width, height, ow, pack_xtrans(xtrans36, x, y));
  return 0;
}

extern "C" int emul_demosaic_downsample4(float *out, const float *in, int width, int height, unsigned filters, const double *cam_to_rgb)
{
  const int ow = (width + 1) / 2, oh = (height + 1) / 2;
  cam_to_rgb_t M;
  for(int k = 0; k < 12; k++) M.m[k] = cam_to_rgb[k];
  emulate(dim3((unsigned)((ow + PNT - 1) / PNT), (unsigned)oh), PNT, downsample4_kernel, in, (float4 *)out, width, height, ow, filters, M);
  return 0;
}
