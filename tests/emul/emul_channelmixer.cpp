/* Test infrastructure: the kernel of ansel_b200/csrc/channelmixer.cu compiled with g++ and run thread by thread on the
 * CPU, with the argument block the product builds (make_cm_args).  Not part of the product. */
#define B200_KERNELS_ON_CPU
#define __constant__ static const
#include "cuda_on_cpu.h"
#include "../../include/b200iop.h"
#include "../../ansel_b200/csrc/channelmixer.cu"

extern "C" int emul_channelmixerrgb(const float *in, float *out, int width, int height, const b200_channelmixerrgb_piece_t *pc)
{
  cm_args_t a;
  if(!make_cm_args(pc, &a)) return 0;
  const size_t npx = (size_t)width * height;
  emulate(dim3((unsigned)((npx + CNT - 1) / CNT)), CNT, channelmixer_kernel, (const float4 *)in, (float4 *)out, npx, a);
  return 0;
}
