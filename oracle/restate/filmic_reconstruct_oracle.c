/* CPU restatement of filmic rgb's highlight reconstruction (the wavelet inpainting in front of the tone mapping;
 * off by default since the reference deprecated it, `hl_deprecated`).  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src/iop/filmicrgb.c: process() :2729-2838, mask_clipped_pixels :1201-1228,
 * inpaint_noise :1230-1270, wavelets_reconstruct_RGB :1272-1324, wavelets_reconstruct_ratios :1326-1384,
 * init_reconstruct :1387-1400, wavelets_detail_level :1403-1411, get_scales :1414-1431, reconstruct_highlights
 * :1434-1532, get_pixel_norm_simd (EUCLIDEAN_NORM_V1) :1012-1013, compute_ratios :2604-2619, restore_ratios
 * :2622-2639; iop/noise_generator.h: splitmix32 :36-43, xoshiro128plus :54-70, uniform_noise_simd :129-138,
 * gaussian_noise_simd :141-171, poisson_noise_simd :174-204; pixel/bspline.h: sparse_scalar_product :83-117,
 * _bspline_vertical_pass :118-133, _bspline_horizontal :136-151, blur_2D_Bspline :330-350; math/openmp_maths.h
 * fmaxabsf :110-115, clamp_simd :128-131; math/math.h NORM_MIN :37.
 * glibc exp2f / log2f / logf / sinf / cosf through flt32_math.h.
 *
 * Pinned bit-for-bit against those functions cut verbatim (oracle/_ref, ref_filmic.c: ref_filmic_reconstruct).
 */
#include "oracle_common.h"
#include "flt32_math.h"
#include "b200iop.h"
#include <float.h>
#include <stdlib.h>
#include <string.h>

#define REC_MAX_SCALES 10
#define BSPLINE_FSIZE 5
#define NORM_MIN 1.52587890625e-05f
#define MAXF(a, b) (((a) > (b)) ? (a) : (b)) /* glib MAX */

static inline float sqf(float x) { return x * x; }
static inline float clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
static inline float fmaxabsf(float a, float b) { return (fabsf(a) > fabsf(b) && !isnan(a)) ? a : (isnan(b) ? 0.f : b); }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

static inline uint32_t splitmix32(const uint64_t seed)
{
  uint64_t result = (seed ^ (seed >> 33)) * 0x62a9d9ed799705f5ul;
  result = (result ^ (result >> 28)) * 0xcb24d0a5c88c35b3ul;
  return (uint32_t)(result >> 32);
}
static inline float xoshiro128plus(uint32_t state[4])
{
  const uint32_t result = state[0] + state[3];
  const uint32_t t = state[1] << 9;
  state[2] ^= state[0];
  state[3] ^= state[1];
  state[1] ^= state[2];
  state[0] ^= state[3];
  state[2] ^= t;
  state[3] = (state[3] << 11) | (state[3] >> 21);
  return (float)(result >> 8) * 0x1.0p-24f;
}
/* Box-Muller term of lane c; lanes 0..2 draw u1, u2, lane 3 has u1 = u2 = 0 (the reference's zero-initialised arrays) */
static inline float box_muller(float u1, float u2, int flip)
{
  const float radius = sqrtf(-2.0f * f32m_logf(u1));
  const float angle = (float)(2.0 * 3.14159265358979323846 * (double)u2);
  return flip ? radius * f32m_cosf(angle) : radius * f32m_sinf(angle);
}
/* dt_noise_generator_simd(), noise_generator.h:207-236, all four lanes as the vectorised build computes them */
static void noise_simd(int distribution, const float mu[4], const float sigma[4], uint32_t state[4], float out[4])
{
  static const int flip[4] = { 1, 0, 1, 0 };
  float u1[4] = { 0.f, 0.f, 0.f, 0.f }, u2[4] = { 0.f, 0.f, 0.f, 0.f };
  if(distribution == 1)
  { /* gaussian: three u1 first, then three u2 (:152-158) */
    for(int c = 0; c < 3; c++) u1[c] = fmaxf(xoshiro128plus(state), FLT_MIN);
    for(int c = 0; c < 3; c++) u2[c] = xoshiro128plus(state);
    for(int c = 0; c < 4; c++) out[c] = box_muller(u1[c], u2[c], flip[c]) * sigma[c] + mu[c];
  }
  else if(distribution == 2)
  { /* poissonian: u1, u2 interleaved (:182-186), Anscombe transform on top */
    for(int c = 0; c < 3; c++)
    {
      u1[c] = fmaxf(xoshiro128plus(state), FLT_MIN);
      u2[c] = xoshiro128plus(state);
    }
    for(int c = 0; c < 4; c++)
    {
      const float noise = box_muller(u1[c], u2[c], flip[c]);
      const float r = noise * sigma[c] + 2.0f * sqrtf(fmaxf(mu[c] + 3.f / 8.f, 0.0f));
      out[c] = (r * r - sigma[c] * sigma[c]) / 4.f - 3.f / 8.f;
    }
  }
  else
  { /* uniform (and default), :129-138 */
    float noise[4] = { 0.f, 0.f, 0.f, 0.f };
    for(int c = 0; c < 3; c++) noise[c] = xoshiro128plus(state);
    for(int c = 0; c < 4; c++) out[c] = mu[c] + 2.0f * (noise[c] - 0.5f) * sigma[c];
  }
}

/* mask_clipped_pixels(), :1201-1228 */
static int mask_clipped(const float *in, float *mask, float normalize, float feathering, size_t npx)
{
  long clipped = 0;
  for(size_t k = 0; k < npx; k++)
  {
    const float *p = in + 4 * k;
    const float pix_max = fmaxf(sqrtf(sqf(p[0]) + sqf(p[1]) + sqf(p[2])), 0.f);
    const float argument = -pix_max * normalize + feathering;
    mask[k] = clamp01(1.0f / (1.0f + f32m_exp2f(argument)));
    clipped += (4.f > argument);
  }
  return (int)clipped > 9; /* the reference counts in an int */
}
/* inpaint_noise(), :1230-1270 */
static void inpaint_noise(const float *in, const float *mask, float *inpainted, float noise_level, float threshold, int distribution, size_t width,
                          size_t height)
{
#pragma omp parallel
  {
    orc_fp_fast_mode();
#pragma omp for
    for(size_t i = 0; i < height; i++)
      for(size_t j = 0; j < width; j++)
      {
        uint32_t state[4] = { splitmix32(j + 1), splitmix32((j + 1) * (i + 3)), splitmix32(1337), splitmix32(666) };
        xoshiro128plus(state);
        xoshiro128plus(state);
        xoshiro128plus(state);
        xoshiro128plus(state);
        const size_t idx = i * width + j;
        const float weight = mask[idx];
        const float *pix_in = in + 4 * idx;
        float noise[4], sigma[4];
        for(int c = 0; c < 4; c++) sigma[c] = pix_in[c] * noise_level / threshold;
        noise_simd(distribution, pix_in, sigma, state, noise);
        for(int c = 0; c < 4; c++) inpainted[4 * idx + c] = fmaxf(pix_in[c] * (1.0f - weight) + weight * noise[c], 0.f);
      }
  }
}
/* blur_2D_Bspline(), bspline.h:330-350 */
static void blur_bspline(const float *in, float *out, int width, int height, int mult, int clip)
{
  static const float f[5] = { 1.0f / 16.0f, 4.0f / 16.0f, 6.0f / 16.0f, 4.0f / 16.0f, 1.0f / 16.0f };
#pragma omp parallel
  {
    orc_fp_fast_mode();
    float *temp = malloc(sizeof(float) * 4 * (size_t)width);
#pragma omp for
    for(int i = 0; i < height; i++)
    {
      const size_t r[5] = { (size_t)4 * width * imax(i - 2 * mult, 0), (size_t)4 * width * imax(i - mult, 0), (size_t)4 * width * i,
                            (size_t)4 * width * imin(i + mult, height - 1), (size_t)4 * width * imin(i + 2 * mult, height - 1) };
      for(int j = 0; j < width; j++)
        for(int c = 0; c < 4; c++)
        {
          const float *b = in + 4 * (size_t)j + c;
          const float v = f[0] * b[r[0]] + f[1] * b[r[1]] + f[2] * b[r[2]] + f[3] * b[r[3]] + f[4] * b[r[4]];
          temp[4 * j + c] = clip ? MAXF(0.0f, v) : v;
        }
      for(int j = 0; j < width; j++)
      {
        const int x[5] = { 4 * imax(j - 2 * mult, 0), 4 * imax(j - mult, 0), 4 * j, 4 * imin(j + mult, width - 1), 4 * imin(j + 2 * mult, width - 1) };
        const size_t index = 4 * ((size_t)i * width + j);
        for(int c = 0; c < 4; c++)
        {
          const float v = f[0] * temp[x[0] + c] + f[1] * temp[x[1] + c] + f[2] * temp[x[2] + c] + f[3] * temp[x[3] + c] + f[4] * temp[x[4] + c];
          out[index + c] = clip ? MAXF(0.0f, v) : v;
        }
      }
    }
    free(temp);
  }
}
/* get_scales(), :1414-1431 */
int orc_filmic_reconstruct_scales(float iscale, double roi_scale, int buf_w, int buf_h)
{
  const float scale = 1.0f / (float)((double)iscale / roi_scale); /* dt_dev_get_module_scale: float / double */
  const size_t size = MAXF(buf_h * iscale, buf_w * iscale);
  const int scales = floorf(f32m_log2f((2.0f * size * scale / ((BSPLINE_FSIZE - 1) * BSPLINE_FSIZE)) - 1.0f));
  return scales > REC_MAX_SCALES ? REC_MAX_SCALES : (scales < 1 ? 1 : scales);
}
/* reconstruct_highlights(), :1434-1532; variant 0 = RGB, 1 = ratios */
static void reconstruct(const float *in, const float *mask, float *reconstructed, int variant, const b200_filmicrgb_data_t *d, int scales, int width,
                        int height)
{
  const size_t npx = (size_t)width * height, n = 4 * npx;
  float *LF_even = malloc(sizeof(float) * n), *LF_odd = malloc(sizeof(float) * n), *HF_RGB = malloc(sizeof(float) * n),
        *HF_grey = malloc(sizeof(float) * n);
  for(size_t k = 0; k < npx; k++) /* init_reconstruct */
    for(int c = 0; c < 4; c++) reconstructed[4 * k + c] = fmaxf(in[4 * k + c] * (1.f - mask[k]), 0.f);
  const float gamma = d->reconstruct_structure_vs_texture, gamma_comp = 1.0f - d->reconstruct_structure_vs_texture;
  const float beta = d->reconstruct_grey_vs_color, beta_comp = 1.f - d->reconstruct_grey_vs_color;
  const float delta = d->reconstruct_bloom_vs_details;
  for(int s = 0; s < scales; ++s)
  {
    const float *detail = s == 0 ? in : (s % 2 != 0 ? LF_odd : LF_even);
    float *LF = s == 0 ? LF_odd : (s % 2 != 0 ? LF_even : LF_odd);
    float *HF_RGB_temp = s == 0 ? LF_even : (s % 2 != 0 ? LF_odd : LF_even);
    blur_bspline(detail, LF, width, height, 1 << s, 1);
    for(size_t k = 0; k < n; k++) HF_RGB_temp[k] = HF_grey[k] = detail[k] - LF[k]; /* wavelets_detail_level */
    blur_bspline(HF_RGB_temp, HF_RGB, width, height, 1, 0);
    for(size_t k = 0; k < n; k += 4)
    {
      const float alpha = mask[k / 4];
      const float *HF_c = HF_RGB + k, *LF_c = LF + k, *TT_c = HF_grey + k;
      const float grey_texture = fmaxabsf(fmaxabsf(TT_c[0], TT_c[1]), TT_c[2]);
      const float grey_details = (HF_c[0] + HF_c[1] + HF_c[2]) / 3.f;
      if(variant == 0)
      {
        const float grey_HF = beta_comp * (gamma_comp * grey_details + gamma * grey_texture);
        const float grey_residual = beta_comp * (LF_c[0] + LF_c[1] + LF_c[2]) / 3.f;
        for(int c = 0; c < 4; c++)
        {
          const float details = (gamma_comp * HF_c[c] + gamma * TT_c[c]) * beta + grey_HF;
          const float residual = (s == scales - 1) ? (grey_residual + LF_c[c] * beta) : 0.f;
          reconstructed[k + c] += alpha * (delta * details + residual);
        }
      }
      else
      {
        const float grey_HF = (gamma_comp * grey_details + gamma * grey_texture);
        for(int c = 0; c < 4; c++)
        {
          const float details = 0.5f * ((gamma_comp * HF_c[c] + gamma * TT_c[c]) + grey_HF);
          const float residual = (s == scales - 1) ? LF_c[c] : 0.f;
          reconstructed[k + c] += alpha * (delta * details + residual);
        }
      }
    }
  }
  free(LF_even);
  free(LF_odd);
  free(HF_RGB);
  free(HF_grey);
}

/* process() :2729-2838.  Returns 1 with the reconstructed frame, 0 when nothing is recovered (out = in). */
int orc_filmic_reconstruct(const float *in, float *out, float *mask_out, size_t width, size_t height, const b200_filmicrgb_data_t *d, float iscale,
                           double roi_scale, int buf_w, int buf_h)
{
  orc_fp_fast_mode();
  const size_t npx = width * height;
  float *mask = malloc(sizeof(float) * npx);
  const float scale = fmaxf((float)((double)iscale / roi_scale), 1.f);
  const int recover = mask_clipped(in, mask, d->normalize, d->reconstruct_feather, npx);
  if(mask_out) memcpy(mask_out, mask, sizeof(float) * npx);
  if(!recover)
  {
    memcpy(out, in, sizeof(float) * 4 * npx);
    free(mask);
    return 0;
  }
  float *inpainted = malloc(sizeof(float) * 4 * npx);
  inpaint_noise(in, mask, inpainted, d->noise_level / scale, d->reconstruct_threshold, d->noise_distribution, width, height);
  const int scales = orc_filmic_reconstruct_scales(iscale, roi_scale, buf_w, buf_h);
  reconstruct(inpainted, mask, out, 0, d, scales, (int)width, (int)height);
  free(inpainted);
  if(d->high_quality_reconstruction > 0)
  {
    float *norms = malloc(sizeof(float) * npx), *ratios = malloc(sizeof(float) * 4 * npx);
    for(int i = 0; i < d->high_quality_reconstruction; i++)
    {
      for(size_t k = 0; k < npx; k++)
      { /* compute_ratios, euclidean norm v1 */
        const float *p = out + 4 * k;
        const float norm = fmaxf(sqrtf(sqf(p[0]) + sqf(p[1]) + sqf(p[2])), NORM_MIN);
        norms[k] = norm;
        for(int c = 0; c < 4; c++) ratios[4 * k + c] = p[c] / norm;
      }
      reconstruct(ratios, mask, out, 1, d, scales, (int)width, (int)height);
      for(size_t k = 0; k < npx; k++) /* restore_ratios */
        for(int c = 0; c < 4; c++) out[4 * k + c] = clamp01(out[4 * k + c]) * norms[k];
    }
    free(norms);
    free(ratios);
  }
  free(mask);
  return 1;
}
