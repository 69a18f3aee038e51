/* Test infrastructure: the plan builder and the kernels of ansel_b200/csrc/resample.cu compiled with g++ and run thread
 * by thread on the CPU, driven the way b200_finalscale_process_dev() drives them.  Not part of the product. */
#define B200_KERNELS_ON_CPU
#include "cuda_on_cpu.h"
#include "../../include/b200iop.h"
#include "../../ansel_b200/csrc/resample.cu"

extern "C" int emul_resampling_plan(int interpolator, int in, int in_x0, int out, int out_x0, float scale, int *lengths, float *kernel, int *index, int max_taps)
{
  axis_plan_t P;
  if(!build_axis_plan(interpolator, in, in_x0, out, out_x0, scale, P)) return -1;
  if((int)P.kernel.size() > max_taps) return -3;
  memcpy(lengths, P.length.data(), sizeof(int) * out);
  memcpy(kernel, P.kernel.data(), sizeof(float) * P.kernel.size());
  memcpy(index, P.index.data(), sizeof(int) * P.index.size());
  return (int)P.kernel.size();
}

extern "C" int emul_clip_and_zoom(const float *in, float *out, int in_x, int in_y, int in_w, int in_h, double in_scale, int out_x, int out_y, int out_w, int out_h,
                                  double out_scale, int itor);
extern "C" int emul_finalscale(const float *in, float *out, int in_w, int in_h, double in_scale, int out_w, int out_h, double out_scale, int itor)
{
  return emul_clip_and_zoom(in, out, 0, 0, in_w, in_h, in_scale, 0, 0, out_w, out_h, out_scale, itor);
}
extern "C" int emul_flip(const float *in, float *out, int ch, int width, int height, int orientation)
{
  emulate(dim3((unsigned)((width + RNT - 1) / RNT), (unsigned)height), RNT, flip_kernel, in, out, width, height, ch, orientation);
  return 0;
}
extern "C" int emul_clip_and_zoom(const float *in, float *out, int in_x, int in_y, int in_w, int in_h, double in_scale, int out_x, int out_y, int out_w, int out_h,
                                  double out_scale, int itor)
{
  const dim3 grid((unsigned)((out_w + RNT - 1) / RNT), (unsigned)out_h);
  if(out_scale == 1.f || out_scale == in_scale)
  {
    emulate(grid, RNT, crop_rows_kernel, (const float4 *)in, (float4 *)out, in_w, out_w, out_x - in_x, out_y - in_y);
    return 0;
  }
  const float resample_scale = (float)(out_scale / in_scale);
  axis_plan_t H, V;
  if(!build_axis_plan(itor, in_w, in_x, out_w, out_x, resample_scale, H) || !build_axis_plan(itor, in_h, in_y, out_h, out_y, resample_scale, V)) return 1;
  const plan_view_t P = { H.length.data(), H.offset.data(), H.index.data(), V.length.data(), V.offset.data(), V.index.data(), H.kernel.data(), V.kernel.data() };
  emulate(grid, RNT, resample_kernel, (const float4 *)in, (float4 *)out, in_w, out_w, P);
  return 0;
}
