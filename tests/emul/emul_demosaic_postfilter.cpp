/* Test infrastructure: the kernels of ansel_b200/csrc/demosaic_postfilter.cu compiled with g++ and run thread by thread on the
 * CPU, in the order demosaic_postfilter_dev() launches them.  Not part of the product. */
#define B200_KERNELS_ON_CPU
#include "cuda_on_cpu.h"
#include "../../include/b200iop.h"
#include "../../ansel_b200/csrc/demosaic_postfilter.cu"
#include <vector>

extern "C" int emul_demosaic_downsample_postfilter(float *rgba, int width, int height, int iterations)
{
  static_assert(GUIDED_SCALES == 1, "the emulation follows the one-scale launch order");
  const size_t px = (size_t)width * height;
  std::vector<float4> LF(px), HF(px), coeff(px), bias(px), tc(px), tb(px);
  float4 *out = (float4 *)rgba;
  const dim3 grid((unsigned)((width + FNT - 1) / FNT), (unsigned)height), grid2(grid.x, grid.y, 2);
  for(int it = 0; it < iterations; it++)
  {
    emulate(grid, FNT, pf_vertical_kernel<true>, (const float4 *)out, (const float4 *)out, tc.data(), tc.data(), width, height, 1);
    emulate(grid, FNT, pf_decompose_kernel, (const float4 *)tc.data(), (const float4 *)out, LF.data(), HF.data(), width, 1);
    emulate(grid, FNT, pf_fit_kernel, (const float4 *)HF.data(), coeff.data(), bias.data(), width, height);
    emulate(grid2, FNT, pf_vertical_kernel<false>, (const float4 *)coeff.data(), (const float4 *)bias.data(), tc.data(), tb.data(), width, height, 1);
    emulate(grid, FNT, pf_apply_kernel<true>, (const float4 *)tc.data(), (const float4 *)tb.data(), (const float4 *)HF.data(), (const float4 *)LF.data(), (float4 *)nullptr, out,
            width, 1);
  }
  return 0;
}
