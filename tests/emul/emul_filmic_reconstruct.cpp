/* Test infrastructure: the kernels of ansel_b200/csrc/filmic_reconstruct.cu compiled with g++ and run thread by thread
 * on the CPU, driven in the order filmic_reconstruct_dev() launches them.  Not part of the product. */
#define B200_KERNELS_ON_CPU
#include "cuda_on_cpu.h"
#include "../../include/b200iop.h"
#include "../../ansel_b200/csrc/filmic_reconstruct.cu"
#include <vector>

static void reconstruct(const float4 *in, const float *mask, float4 *rec, int variant, const b200_filmicrgb_data_t *d, int scales, int width, int height,
                        float4 *LF_even, float4 *LF_odd, float4 *HF_grey, float4 *vtmp)
{
  const size_t npx = (size_t)width * height;
  const dim3 lin((unsigned)((npx + NT - 1) / NT)), grid((width + NT - 1) / NT, height);
  emulate(lin, NT, init_kernel, in, mask, rec, npx);
  rec_args_t a;
  a.gamma = d->reconstruct_structure_vs_texture;
  a.gamma_comp = 1.0f - d->reconstruct_structure_vs_texture;
  a.beta = d->reconstruct_grey_vs_color;
  a.beta_comp = 1.f - d->reconstruct_grey_vs_color;
  a.delta = d->reconstruct_bloom_vs_details;
  for(int s = 0; s < scales; ++s)
  {
    const float4 *detail = s == 0 ? in : (s % 2 != 0 ? LF_odd : LF_even);
    float4 *LF = s == 0 ? LF_odd : (s % 2 != 0 ? LF_even : LF_odd);
    float4 *HF_temp = s == 0 ? LF_even : (s % 2 != 0 ? LF_odd : LF_even);
    emulate(grid, NT, blur_vertical_kernel<true>, detail, vtmp, width, height, 1 << s);
    emulate(grid, NT, blur_horizontal_detail_kernel, (const float4 *)vtmp, detail, LF, HF_temp, HF_grey, width, 1 << s);
    emulate(grid, NT, blur_vertical_kernel<false>, (const float4 *)HF_temp, vtmp, width, height, 1);
    a.last = (s == scales - 1) ? 1 : 0;
    if(variant == 0)
      emulate(grid, NT, blur_horizontal_reconstruct_kernel<0>, (const float4 *)vtmp, (const float4 *)LF, (const float4 *)HF_grey, mask, rec, width, a);
    else
      emulate(grid, NT, blur_horizontal_reconstruct_kernel<1>, (const float4 *)vtmp, (const float4 *)LF, (const float4 *)HF_grey, mask, rec, width, a);
  }
}

extern "C" int emul_filmic_reconstruct(const float *in, float *out, float *mask_out, int width, int height, const b200_filmicrgb_data_t *d, float iscale,
                                       double roi_scale, int scales)
{
  const size_t npx = (size_t)width * height;
  std::vector<float> mask(npx), norms(npx);
  std::vector<float4> inpainted(npx), rec(npx), LF_even(npx), LF_odd(npx), HF_grey(npx), vtmp(npx);
  const dim3 lin((unsigned)((npx + NT - 1) / NT)), grid((width + NT - 1) / NT, height);
  int count_unused = 0;
  emulate(lin, NT, mask_kernel, (const float4 *)in, mask.data(), &count_unused, npx, d->normalize, d->reconstruct_feather);
  memcpy(mask_out, mask.data(), npx * sizeof(float));
  const float module_scale = (float)((double)iscale / roi_scale);
  const float scale = fmaxf(module_scale, 1.f);
  emulate(grid, NT, inpaint_noise_kernel, (const float4 *)in, (const float *)mask.data(), inpainted.data(), width, height, d->noise_level / scale,
          d->reconstruct_threshold, d->noise_distribution);
  reconstruct(inpainted.data(), mask.data(), rec.data(), 0, d, scales, width, height, LF_even.data(), LF_odd.data(), HF_grey.data(), vtmp.data());
  float4 *ratios = inpainted.data();
  for(int i = 0; i < d->high_quality_reconstruction; i++)
  {
    emulate(lin, NT, ratios_kernel, (const float4 *)rec.data(), norms.data(), ratios, npx);
    reconstruct(ratios, mask.data(), rec.data(), 1, d, scales, width, height, LF_even.data(), LF_odd.data(), HF_grey.data(), vtmp.data());
    emulate(lin, NT, restore_kernel, rec.data(), (const float *)norms.data(), npx);
  }
  memcpy(out, rec.data(), npx * sizeof(float4));
  return 1;
}
