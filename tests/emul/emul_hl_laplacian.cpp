/* Test infrastructure: the kernels and the launch sequence of ansel_b200/csrc/highlights_laplacian.cu compiled with g++, every kernel run thread
 * by thread on the CPU.  The normalization vector is handed in (the device sums it with block reductions, which this harness cannot run).
 * Not part of the product. */
#define B200_KERNELS_ON_CPU
#include "cuda_on_cpu.h"
#include "../../include/b200iop.h"
#include "../../ansel_b200/csrc/highlights_laplacian.cu"
#include <vector>

extern "C" int emul_hl_laplacian_scales(int scales_param, float iscale, float roi_scale) { return hl_scales(scales_param, iscale / roi_scale); }

extern "C" int emul_hl_laplacian(const float *in, float *out, int width, int height, uint32_t shifted_filters, const uint8_t shifted_xtrans[36], const float clips[4],
                                 int iterations,
                                 int scales_param, float noise_level, float solid_color, float iscale, float roi_scale, const float normalization[4])
{
  hl_job_t J;
  J.width = width;
  J.height = height;
  J.ds_width = width / DS_FACTOR;
  J.ds_height = height / DS_FACTOR;
  J.filters = shifted_filters;
  memset(&J.xt, 0, sizeof(J.xt));
  if(shifted_xtrans) memcpy(J.xt.v, shifted_xtrans, 36);
  for(int c = 0; c < 4; c++) J.clips.v[c] = clips[c];
  const float module_scale = iscale / roi_scale;
  J.iterations = iterations;
  J.scales = hl_scales(scales_param, module_scale);
  J.noise_level = noise_level / (DS_FACTOR * module_scale);
  J.solid_color = solid_color;
  const size_t npx = (size_t)width * height, ds_npx = (size_t)J.ds_width * J.ds_height;
  std::vector<float4> full[3], quarter[7];
  for(auto &v : full) v.resize(npx);
  for(auto &v : quarter) v.resize(ds_npx ? ds_npx : 1);
  float norm[4] = { normalization[0], normalization[1], normalization[2], normalization[3] };
  hl_buffers_t B = { full[0].data(),    full[1].data(),    full[2].data(),    quarter[0].data(), quarter[1].data(), quarter[2].data(),
                     quarter[3].data(), quarter[4].data(), quarter[5].data(), quarter[6].data(), norm };
  return hl_sequence(J, B, in, out, nullptr);
}
