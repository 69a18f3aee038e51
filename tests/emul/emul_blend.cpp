/* Test infrastructure: the kernel of ansel_b200/csrc/blend.cu and its host-side plan compiled with g++, the kernel run thread by thread on the
 * CPU.  Not part of the product. */
#define B200_KERNELS_ON_CPU
#include "cuda_on_cpu.h"
#include "../../include/b200iop.h"
#include "../../ansel_b200/csrc/blend.cu"

extern "C" int emul_blend_process(const float *in, float *out, int iw, int ih, int ow, int oh, int xoffs, int yoffs, const b200_blend_params_t *bp,
                                  const float *form, float *mask_out)
{
  (void)ih;
  blend_plan_t pl;
  const int rc = bl_plan(pl, bp, form != nullptr);
  if(rc == 1) return 0;
  if(rc) return -1;
  pl.in = (const float4 *)in;
  pl.out = (float4 *)out;
  pl.form = form;
  pl.mask_out = mask_out;
  pl.iw = iw;
  pl.ow = ow;
  pl.oh = oh;
  pl.xoffs = xoffs;
  pl.yoffs = yoffs;
  if(pl.raw)
    emulate(dim3((unsigned)((ow + 255) / 256), (unsigned)oh), 256, blend_raw_kernel, pl);
  else
    emulate(dim3((unsigned)((ow + 255) / 256), (unsigned)oh), 256, blend_kernel, pl);
  return 0;
}
