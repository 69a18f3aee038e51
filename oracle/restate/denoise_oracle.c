/* CPU restatement of profiled denoise, wavelet mode.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src
 *   pixel/eaw.c            eaw_dn_decompose :242-326 (dn_weight :181-195, fast_mexp2f math/math.h:303-317),
 *                          eaw_synthesize :157-175
 *   iop/denoiseprofile.c   process_wavelets :1289-1447, variance_stabilizing_xform :1223-1286,
 *                          compute_wb_factors :1098-1129, set_up_conversion_matrices :1170-1221,
 *                          invert_matrix :1132-1166, precondition/backtransform{,_v2,_Y0U0V0} :852-1089
 *
 * Pinning: the two eaw functions are checked bit-for-bit against the reference's own eaw.c compiled
 * in place (oracle/_ref).  iop/denoiseprofile.c is one translation unit with its GTK GUI and cannot
 * be compiled whole, so its pixel functions (:852-1286: the VST pairs, compute_wb_factors,
 * set_up_conversion_matrices, variance_stabilizing_xform) are cut out verbatim at build time
 * (oracle/ref_shim/slice.py) and compiled; the restatements below match them bit for bit
 * (tests/test_cpu_oracle_pin.py).  Only the ~40 lines of glue in process_wavelets :1289-1447 that
 * derive p, compensate_p and the scale count are restated without a compiled counterpart.
 *
 * One deliberate difference: the per-channel sum of squared detail coefficients.  The reference
 * accumulates it in float with an OpenMP reduction (eaw.c:236,253,318-324), so its value depends
 * on the thread count and saturates near 2^24 for large frames on one thread.  Here it is
 * accumulated in double and rounded once to float -- the value the reference's sum approximates.
 */
#include "oracle_common.h"
#include "flt32_math.h"
#include "b200iop.h"
#include <stdlib.h>
#include <string.h>

#define MAXF(a, b) ((a) > (b) ? (a) : (b))

/* math/math.h:303-317 -- the "incorrect, reduced precision" variant eaw.c uses, kept as is */
static inline float mexp2_float(float x)
{
  const float i1 = (float)0x3f800000u, i2 = (float)0x3f000000u;
  const float k0 = i1 + x * (i2 - i1);
  const int32_t ki = k0 >= (float)0x800000u ? (int32_t)k0 : 0;
  float f;
  memcpy(&f, &ki, 4);
  return f;
}

/* eaw.c:181-195 */
static inline float dn_weight(const float *c1, const float *c2, float inv_sigma2)
{
  float sqr[4];
  for(int c = 0; c < 4; c++)
  {
    const float diff = c1[c] - c2[c];
    sqr[c] = diff * diff;
  }
  const float dot = (sqr[0] + sqr[1] + sqr[2]) * inv_sigma2;
  const float t = dot * 0.02f - 9.0f;
  return mexp2_float((0 > t) ? 0 : t); /* MAX(0, t): a NaN passes through, as in the macro */
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* eaw.c:242-326.  The reference special-cases interior pixels only to skip the clamping; the taps
 * and their order (rows outer, columns inner) are the same. */
void orc_eaw_dn_decompose(float *coarse, const float *in, float *detail, double sum_squared[4], int scale,
                          float inv_sigma2, int width, int height)
{
  static const float filter[5] = { 1.0f / 16.0f, 4.0f / 16.0f, 6.0f / 16.0f, 4.0f / 16.0f, 1.0f / 16.0f };
  const int mult = 1 << scale;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma omp parallel for schedule(static) reduction(+ : s0, s1, s2, s3)
  for(int j = 0; j < height; j++)
  {
    double r[4] = { 0, 0, 0, 0 };
    for(int i = 0; i < width; i++)
    {
      const float *px = in + 4 * ((size_t)j * width + i);
      float sum[4] = { 0, 0, 0, 0 }, wgt[4] = { 0, 0, 0, 0 };
      for(int jj = 0; jj < 5; jj++)
      {
        const int y = clampi(j + mult * (jj - 2), 0, height - 1);
        for(int ii = 0; ii < 5; ii++)
        {
          const int x = clampi(i + mult * (ii - 2), 0, width - 1);
          const float *px2 = in + 4 * ((size_t)y * width + x);
          const float f = filter[ii] * filter[jj];
          const float w = f * dn_weight(px, px2, inv_sigma2);
          for(int c = 0; c < 4; c++)
          {
            const float pd = w * px2[c];
            wgt[c] += w;
            sum[c] += pd;
          }
        }
      }
      for(int c = 0; c < 4; c++)
      {
        sum[c] /= wgt[c];
        coarse[4 * ((size_t)j * width + i) + c] = sum[c];
        const float det = px[c] - sum[c];
        detail[4 * ((size_t)j * width + i) + c] = det;
        r[c] += (double)(det * det);
      }
    }
    s0 += r[0];
    s1 += r[1];
    s2 += r[2];
    s3 += r[3];
  }
  sum_squared[0] = s0;
  sum_squared[1] = s1;
  sum_squared[2] = s2;
  sum_squared[3] = s3;
}

/* eaw.c:157-175 */
void orc_eaw_synthesize(float *out, const float *in, const float *detail, const float threshold[4], const float boost[4],
                        int width, int height)
{
  const size_t n = (size_t)width * height;
#pragma omp parallel for schedule(static)
  for(size_t k = 0; k < n; k++)
    for(int c = 0; c < 4; c++)
    {
      const float d = detail[4 * k + c];
      const float amount = MAXF(d - threshold[c], 0.0f) + ((d + threshold[c]) < 0.0f ? (d + threshold[c]) : 0.0f);
      out[4 * k + c] = in[4 * k + c] + boost[c] * amount;
    }
}

/* ---- host-side plan of process_wavelets ----------------------------------------------------- */
typedef struct
{
  int max_scale;
  float wb[4], p[4];
  float a_eff, b, bias_eff;
  float toY[3][4], toRGB[3][4];
  float aa[4], bb[4]; /* old VST */
  float sigma_band[B200_DENOISE_BANDS];
} wavelet_plan_t;

/* denoiseprofile.c:1098-1129 */
static void wb_factors(float wb[4], const b200_denoiseprofile_data_t *d, const float coeffs[4], const float pm[4],
                       const float weights[4])
{
  const float wb_mean = (coeffs[0] + coeffs[1] + coeffs[2]) / 3.0f;
  wb[0] = wb[1] = wb[2] = wb[3] = wb_mean;
  if(d->fix_anscombe_and_nlmeans_norm)
  {
    if(wb_mean != 0.0f && d->wb_adaptive_anscombe)
      for(int i = 0; i < 3; i++) wb[i] = coeffs[i];
    else if(wb_mean == 0.0f)
      for(int i = 0; i < 4; i++) wb[i] = 1.0f;
  }
  else
    for(int i = 0; i < 4; i++) wb[i] = weights[i] * pm[i];
}

/* denoiseprofile.c:1132-1166 */
static int invert3(float in[3][4], float out[3][4])
{
  const float biga = in[1][1] * in[2][2] - in[1][2] * in[2][1];
  const float bigb = -in[1][0] * in[2][2] + in[1][2] * in[2][0];
  const float bigc = in[1][0] * in[2][1] - in[1][1] * in[2][0];
  const float bigd = -in[0][1] * in[2][2] + in[0][2] * in[2][1];
  const float bige = in[0][0] * in[2][2] - in[0][2] * in[2][0];
  const float bigf = -in[0][0] * in[2][1] + in[0][1] * in[2][0];
  const float bigg = in[0][1] * in[1][2] - in[0][2] * in[1][1];
  const float bigh = -in[0][0] * in[1][2] + in[0][2] * in[1][0];
  const float bigi = in[0][0] * in[1][1] - in[0][1] * in[1][0];
  const float det = in[0][0] * biga + in[0][1] * bigb + in[0][2] * bigc;
  if(det == 0.0f) return 0;
  out[0][0] = 1.0f / det * biga;
  out[0][1] = 1.0f / det * bigd;
  out[0][2] = 1.0f / det * bigg;
  out[0][3] = 0.0f;
  out[1][0] = 1.0f / det * bigb;
  out[1][1] = 1.0f / det * bige;
  out[1][2] = 1.0f / det * bigh;
  out[1][3] = 0.0f;
  out[2][0] = 1.0f / det * bigc;
  out[2][1] = 1.0f / det * bigf;
  out[2][2] = 1.0f / det * bigi;
  out[2][3] = 0.0f;
  return 1;
}

/* denoiseprofile.c:1170-1221 */
static void conversion_matrices(float toY[3][4], float toRGB[3][4], const float wb[4])
{
  float sum_invwb = 1.0f / wb[0] + 1.0f / wb[1] + 1.0f / wb[2];
  sum_invwb *= sqrtf(3);
  toY[0][0] = sum_invwb / wb[0];
  toY[0][1] = sum_invwb / wb[1];
  toY[0][2] = sum_invwb / wb[2];
  toY[0][3] = 0.0f;
  const float sdU = sqrtf(0.5f * 0.5f * wb[0] * wb[0] + 0.5f * 0.5f * wb[2] * wb[2]);
  const float sdV = sqrtf(0.25f * 0.25f * wb[0] * wb[0] + 0.5f * 0.5f * wb[1] * wb[1] + 0.25f * 0.25f * wb[2] * wb[2]);
  for(int c = 0; c < 3; c++)
  {
    toY[1][c] /= sdU;
    toY[2][c] /= sdV;
  }
  toY[1][3] = toY[2][3] = 0.0f;
  if(!invert3(toY, toRGB))
  {
    const float sdY = sqrtf(1.0f / 9.0f * (wb[0] * wb[0] + wb[1] * wb[1] + wb[2] * wb[2]));
    toY[0][0] = toY[0][1] = toY[0][2] = 1.0f / (3.0f * sdY);
    toY[0][3] = 0.0f;
    invert3(toY, toRGB);
  }
}

/* denoiseprofile.c:1301-1317 (the same loop sizes the tiling overlap, :818-836) */
static int wavelet_max_scale(float roi_scale, int buf_w, int buf_h)
{
  int max_scale = 0;
  const float in_scale = fminf(roi_scale, 1.0f);
  const float big = (float)MAXF(buf_h, buf_w) * 0.2f;
  const float cap = (float)(2 * (2u << (B200_DENOISE_BANDS - 1)) + 1);
  const float supp0 = cap < big ? cap : big;
  const float i0 = f32m_log2f((supp0 - 1.0f) * .5f);
  for(; max_scale < B200_DENOISE_BANDS; max_scale++)
  {
    const float supp = (float)(2 * (2u << max_scale) + 1);
    const float supp_in = supp * (1.0f / in_scale);
    const float i_in = f32m_log2f((supp_in - 1) * .5f) - 1.0f;
    const float t = 1.0f - (i_in + .5f) / i0;
    if(t < 0.0f) break;
  }
  return max_scale;
}

static void make_plan_ex(wavelet_plan_t *pl, const b200_denoiseprofile_data_t *d, float roi_scale, int buf_w, int buf_h,
                         const float wb_coeffs[4], const float pm[4], int nlm);
static void make_plan(wavelet_plan_t *pl, const b200_denoiseprofile_data_t *d, float roi_scale, int buf_w, int buf_h,
                      const float wb_coeffs[4], const float pm[4])
{
  make_plan_ex(pl, d, roi_scale, buf_w, buf_h, wb_coeffs, pm, 0);
}
/* nlm = 1: nlmeans_precondition(), denoiseprofile.c:1500-1533 (unit wb weights, no Y0U0V0 matrices, strength*scale) */
static void make_plan_ex(wavelet_plan_t *pl, const b200_denoiseprofile_data_t *d, float roi_scale, int buf_w, int buf_h,
                         const float wb_coeffs[4], const float pm[4], int nlm)
{
  memset(pl, 0, sizeof(*pl));
  pl->max_scale = wavelet_max_scale(roi_scale, buf_w, buf_h);
  const float in_scale = fminf(roi_scale, 1.0f);
  const float wb_weights[4] = { 2.0f, 1.0f, 2.0f, 0.0f }, wb_unit[4] = { 1.0f, 1.0f, 1.0f, 0.0f };
  wb_factors(pl->wb, d, wb_coeffs, pm, nlm ? wb_unit : wb_weights);
  for(int c = 0; c < 3; c++)
  { /* MAX(d->shadows + 0.1 * logf(..), 0.0f): the 0.1 makes it a double expression, :1348-1351 */
    const double v = (double)d->shadows + 0.1 * (double)f32m_logf(in_scale / pl->wb[c]);
    pl->p[c] = (float)(v > 0.0 ? v : 0.0);
  }
  pl->p[3] = 0.0f;
  const float compensate_p = 0.05f / f32m_powf(0.05f, d->shadows);
  if(nlm)
  {
    for(int i = 0; i < 4; i++)
    {
      pl->wb[i] *= d->strength * in_scale;
      pl->aa[i] = d->a[1] * pl->wb[i];
      pl->bb[i] = d->b[1] * pl->wb[i];
    }
    pl->a_eff = d->a[1] * compensate_p;
    pl->b = d->b[1];
    pl->bias_eff = (float)((double)d->bias - 0.5 * (double)f32m_logf(in_scale));
    return;
  }
  float toY[3][4] = { { 1.0f / 3.0f, 1.0f / 3.0f, 1.0f / 3.0f, 0 }, { 0.5f, 0.0f, -0.5f, 0 }, { 0.25f, -0.5f, 0.25f, 0 } };
  float toRGB[3][4] = { { 0 } };
  conversion_matrices(toY, toRGB, pl->wb);
  const float cs = (d->wavelet_color_mode == B200_DENOISE_RGB) ? 1.0f : 2.5f;
  for(int k = 0; k < 3; k++)
    for(int c = 0; c < 4; c++)
    {
      toY[k][c] /= (d->strength * cs * in_scale);
      toRGB[k][c] *= (d->strength * cs * in_scale);
    }
  for(int i = 0; i < 4; i++) pl->wb[i] *= d->strength * cs * in_scale;
  memcpy(pl->toY, toY, sizeof(toY));
  memcpy(pl->toRGB, toRGB, sizeof(toRGB));
  for(int c = 0; c < 3; c++)
  {
    pl->aa[c] = d->a[1] * pl->wb[c];
    pl->bb[c] = d->b[1] * pl->wb[c];
  }
  pl->a_eff = d->a[1] * compensate_p;
  pl->b = d->b[1];
  pl->bias_eff = (float)((double)d->bias - 0.5 * (double)f32m_logf(in_scale));
  const float varf = sqrtf(2.0f + 2.0f * 4.0f * 4.0f + 6.0f * 6.0f) / 16.0f;
  for(int s = 0; s < B200_DENOISE_BANDS; s++) pl->sigma_band[s] = f32m_powf(varf, (float)s) * 1.0f;
}

/* ---- variance-stabilising transforms, pointwise, all four lanes ------------------------------ */
static void vst_forward(const wavelet_plan_t *pl, const b200_denoiseprofile_data_t *d, const float *in, float *buf, size_t npx)
{
  if(!d->use_new_vst)
  { /* precondition(), :852-870 */
    float s38[4] = { 0, 0, 0, 0 };
    for(int c = 0; c < 3; c++) s38[c] = (pl->bb[c] / pl->aa[c]) * (pl->bb[c] / pl->aa[c]) + 3.f / 8.f;
#pragma omp parallel for schedule(static)
    for(size_t k = 0; k < npx; k++)
      for(int c = 0; c < 4; c++)
      {
        const float v = fmaxf(0.0f, in[4 * k + c] / pl->aa[c] + s38[c]);
        buf[4 * k + c] = 2.0f * sqrtf(v);
      }
    return;
  }
  float expon[4], denom[4], scale[4];
  const float sa = sqrtf(pl->a_eff);
  for(int c = 0; c < 3; c++)
  {
    expon[c] = -pl->p[c] / 2 + 1;
    denom[c] = (-pl->p[c] + 2) * sa;
    scale[c] = 2.0f / ((-pl->p[c] + 2) * sa);
  }
  expon[3] = denom[3] = scale[3] = 1.0f;
  if(d->wavelet_color_mode == B200_DENOISE_RGB)
  { /* precondition_v2(), :925-941 */
#pragma omp parallel for schedule(static)
    for(size_t k = 0; k < npx; k++)
      for(int c = 0; c < 4; c++)
      {
        const float v = in[4 * k + c] / pl->wb[c] + pl->b;
        buf[4 * k + c] = 2.0f * f32m_powf(MAXF(v, 0.0f), expon[c]) / denom[c];
      }
  }
  else
  { /* precondition_Y0U0V0(), :1026-1054 */
#pragma omp parallel for schedule(static)
    for(size_t k = 0; k < npx; k++)
    {
      float tmp[4];
      for(int c = 0; c < 4; c++) tmp[c] = f32m_powf(MAXF(in[4 * k + c] + pl->b, 0.0f), expon[c]) * scale[c];
      for(int c = 0; c < 3; c++)
      {
        float sum = 0.0f;
        for(int j = 0; j < 4; j++) sum += pl->toY[c][j] * tmp[j];
        buf[4 * k + c] = sum;
      }
      buf[4 * k + 3] = 0;
    }
  }
}

static void vst_backward(const wavelet_plan_t *pl, const b200_denoiseprofile_data_t *d, float *buf, size_t npx)
{
  if(!d->use_new_vst)
  { /* backtransform(), :873-899 */
    float s18[4] = { 0, 0, 0, 0 };
    for(int c = 0; c < 3; c++) s18[c] = (pl->bb[c] / pl->aa[c]) * (pl->bb[c] / pl->aa[c]) + 1.f / 8.f;
    const float sqrt_3_2 = sqrtf(3.0f / 2.0f);
#pragma omp parallel for schedule(static)
    for(size_t k = 0; k < npx; k++)
      for(int c = 0; c < 4; c++)
      {
        const float x = buf[4 * k + c], x2 = x * x;
        buf[4 * k + c] = (x < 0.5f) ? 0.0f
                                     : pl->aa[c] * (1.f / 4.f * x2 + 1.f / 4.f * sqrt_3_2 / x - 11.f / 8.f / x2
                                                    + 5.f / 8.f * sqrt_3_2 / (x * x2) - s18[c]);
      }
    return;
  }
  float expon[4], denom[4], scale[4], bias_wb[4];
  const float sa = sqrtf(pl->a_eff);
  for(int c = 0; c < 3; c++)
  {
    expon[c] = 1.0f / (1.0f - pl->p[c] / 2.0f);
    denom[c] = 4.0f / (sa * (2.0f - pl->p[c]));
    scale[c] = (sa * (2.0f - pl->p[c])) / 4.0f;
    bias_wb[c] = pl->bias_eff * pl->wb[c];
  }
  expon[3] = denom[3] = scale[3] = 1.0f;
  bias_wb[3] = 0.0f;
  if(d->wavelet_color_mode == B200_DENOISE_RGB)
  { /* backtransform_v2(), :1003-1023 */
#pragma omp parallel for schedule(static)
    for(size_t k = 0; k < npx; k++)
      for(int c = 0; c < 4; c++)
      {
        const float x = MAXF(buf[4 * k + c], 0.0f);
        const float delta = x * x + pl->bias_eff;
        const float z1 = (x + sqrtf(MAXF(delta, 0.0f))) / denom[c];
        buf[4 * k + c] = pl->wb[c] * (f32m_powf(z1, expon[c]) - pl->b);
      }
  }
  else
  { /* backtransform_Y0U0V0(), :1057-1089 */
#pragma omp parallel for schedule(static)
    for(size_t k = 0; k < npx; k++)
    {
      float rgb[4] = { 0, 0, 0, 0 };
      for(int j = 0; j < 3; j++)
        for(int c = 0; c < 4; c++) rgb[j] += pl->toRGB[j][c] * buf[4 * k + c];
      for(int c = 0; c < 4; c++)
      {
        const float x = MAXF(rgb[c], 0.0f);
        const float delta = x * x + bias_wb[c];
        const float z1 = (x + sqrtf(MAXF(delta, 0.0f))) * scale[c];
        buf[4 * k + c] = f32m_powf(z1, expon[c]) - pl->b;
      }
    }
  }
}

/* denoiseprofile.c:1223-1286 */
void orc_wavelet_thresholds(float thrs[4], int scale, int max_scale, size_t npixels, const float sum_y2[4], float sigma_band,
                            const b200_denoiseprofile_data_t *d)
{
  const float sb2 = sigma_band * sigma_band;
  const float nm1 = (float)npixels - 1.0f;
  float std_x[4] = { 0, 0, 0, 1.0f }, adjt[4] = { 8.0f, 8.0f, 8.0f, 0.0f };
  for(int c = 0; c < 3; c++)
  {
    const float var_y = sum_y2[c] / nm1;
    std_x[c] = sqrtf(MAXF(1e-6f, var_y - sb2));
  }
  const int offset_scale = B200_DENOISE_BANDS - max_scale;
  const int band = B200_DENOISE_BANDS - (scale + offset_scale + 1);
  if(d->wavelet_color_mode == B200_DENOISE_RGB)
  {
    float f = d->force[B200_DENOISE_CH_ALL][band];
    f *= f;
    f *= 4;
    for(int c = 0; c < 4; c++) adjt[c] *= f;
    for(int c = 0; c < 3; c++)
    {
      f = d->force[B200_DENOISE_CH_R + c][band];
      f *= f;
      f *= 4;
      adjt[c] *= f;
    }
  }
  else
  {
    float f = d->force[B200_DENOISE_CH_Y0][band];
    f *= f;
    f *= 4;
    adjt[0] *= f;
    f = d->force[B200_DENOISE_CH_U0V0][band];
    f *= f;
    f *= 4;
    adjt[1] *= f;
    adjt[2] *= f;
  }
  for(int c = 0; c < 4; c++) thrs[c] = adjt[c] * sb2 / std_x[c];
}

/* process_wavelets(), :1289-1447.  Returns 0. */
int orc_denoiseprofile_wavelets(const float *in, float *out, int width, int height, const b200_denoiseprofile_data_t *d,
                                float roi_scale, int buf_w, int buf_h, const float wb_coeffs[4], const float pm[4])
{
  wavelet_plan_t pl;
  make_plan(&pl, d, roi_scale, buf_w, buf_h, wb_coeffs, pm);
  const size_t npx = (size_t)width * height;
  const int max_mult = 1 << (pl.max_scale - 1);
  if(width < 2 * max_mult || height < 2 * max_mult)
  {
    memcpy(out, in, sizeof(float) * 4 * npx);
    return 0;
  }
  float *precond = malloc(16 * npx), *tmp = malloc(16 * npx), *buf = malloc(16 * npx);
  if(!precond || !tmp || !buf)
  {
    free(precond);
    free(tmp);
    free(buf);
    return 1;
  }
  vst_forward(&pl, d, in, precond, npx);
  float *buf1 = precond, *buf2 = tmp;
  memset(out, 0, 16 * npx);
  for(int s = 0; s < pl.max_scale; s++)
  {
    double sums[4];
    orc_eaw_dn_decompose(buf2, buf1, buf, sums, s, 1.0f / (pl.sigma_band[s] * pl.sigma_band[s]), width, height);
    const float sum_y2[4] = { (float)sums[0], (float)sums[1], (float)sums[2], (float)sums[3] };
    const float boost[4] = { 1.0f, 1.0f, 1.0f, 1.0f };
    float thrs[4];
    orc_wavelet_thresholds(thrs, s, pl.max_scale, npx, sum_y2, pl.sigma_band[s], d);
    orc_eaw_synthesize(out, out, buf, thrs, boost, width, height);
    float *t = buf2;
    buf2 = buf1;
    buf1 = t;
  }
  for(size_t k = 0; k < 4 * npx; k++) out[k] += buf1[k];
  vst_backward(&pl, d, out, npx);
  free(precond);
  free(tmp);
  free(buf);
  return 0;
}

int orc_wavelet_max_scale(float roi_scale, int buf_w, int buf_h) { return wavelet_max_scale(roi_scale, buf_w, buf_h); }

/* ---- entry points used by the pinning tests --------------------------------------------------- */
/* out[0]=max_scale, [1..4]=wb, [5..8]=p, [9]=a_eff, [10]=b, [11]=bias_eff, [12..23]=toY, [24..35]=toRGB,
 * [36..39]=aa, [40..43]=bb, [44..50]=sigma_band */
void orc_dn_plan_export(const b200_denoiseprofile_data_t *d, float roi_scale, int buf_w, int buf_h, const float wb_coeffs[4],
                        const float pm[4], float out[51])
{
  wavelet_plan_t pl;
  make_plan(&pl, d, roi_scale, buf_w, buf_h, wb_coeffs, pm);
  out[0] = (float)pl.max_scale;
  memcpy(out + 1, pl.wb, 16);
  memcpy(out + 5, pl.p, 16);
  out[9] = pl.a_eff;
  out[10] = pl.b;
  out[11] = pl.bias_eff;
  memcpy(out + 12, pl.toY, 48);
  memcpy(out + 24, pl.toRGB, 48);
  memcpy(out + 36, pl.aa, 16);
  memcpy(out + 40, pl.bb, 16);
  memcpy(out + 44, pl.sigma_band, 28);
}
/* the same for the non-local-means mode (nlmeans_precondition(), denoiseprofile.c:1500-1533): what bench.py's CPU arm hands to the
 * reference's own precondition_v2 / backtransform_v2 around nlmeans_denoise() */
void orc_dn_plan_export_nlm(const b200_denoiseprofile_data_t *d, float roi_scale, int buf_w, int buf_h, const float wb_coeffs[4],
                            const float pm[4], float out[51])
{
  wavelet_plan_t pl;
  make_plan_ex(&pl, d, roi_scale, buf_w, buf_h, wb_coeffs, pm, 1);
  memset(out, 0, 51 * sizeof(float));
  out[0] = (float)pl.max_scale;
  memcpy(out + 1, pl.wb, 16);
  memcpy(out + 5, pl.p, 16);
  out[9] = pl.a_eff;
  out[10] = pl.b;
  out[11] = pl.bias_eff;
  memcpy(out + 36, pl.aa, 16);
  memcpy(out + 40, pl.bb, 16);
}
void orc_dn_vst(int forward, const b200_denoiseprofile_data_t *d, float roi_scale, int buf_w, int buf_h,
                const float wb_coeffs[4], const float pm[4], const float *in, float *out, size_t npx)
{
  wavelet_plan_t pl;
  make_plan(&pl, d, roi_scale, buf_w, buf_h, wb_coeffs, pm);
  if(forward)
    vst_forward(&pl, d, in, out, npx);
  else
  {
    memcpy(out, in, 16 * npx);
    vst_backward(&pl, d, out, npx);
  }
}
void orc_dn_wb_factors(float wb[4], const b200_denoiseprofile_data_t *d, const float coeffs[4], const float pm[4],
                       const float weights[4])
{
  wb_factors(wb, d, coeffs, pm, weights);
}


/* process_nlmeans_cpu(), denoiseprofile.c:1599-1648 */
int orc_nlmeans_denoise(const float *inbuf, float *outbuf, int width, int height, float scattering, float scale, float luma,
                        float chroma, float center_weight, float sharpness, int radius, int search_radius, int decimate,
                        const float norm[4]);
int orc_denoiseprofile_nlmeans(const float *in, float *out, int width, int height, const b200_denoiseprofile_data_t *d,
                               float roi_scale, int pipe_type, const float wb_coeffs[4], const float pm[4])
{
  const size_t npx = (size_t)width * height;
  const float scale = fminf(fminf(roi_scale, 2.0f), 1.0f);
  const int P = (int)ceilf(d->radius * scale);
  int K = (int)d->nbhood;
  float scattering = d->scattering;
  { /* nlmeans_scattering(), :1474-1499 */
    const int has_preview = pipe_type == B200_PIPE_PREVIEW;
    if(has_preview || pipe_type == B200_PIPE_THUMBNAIL)
    {
      const int maxk = (K * K * K + 7.0 * K * sqrt(K)) * scattering / 6.0 + K;
      K = K < 3 ? K : 3;
      scattering = (maxk - K) * 6.0 / (K * K * K + 7.0 * K * sqrt(K));
    }
    if(!has_preview)
    {
      const int maxk = (K * K * K + 7.0 * K * sqrt(K)) * scattering / 6.0 + K;
      K = MAXF((K < 4 ? K : 4), K * scale);
      scattering = (maxk - K) * 6.0 / (K * K * K + 7.0 * K * sqrt(K));
    }
  }
  float norm = .045f / ((2 * P + 1) * (2 * P + 1));
  if(!d->fix_anscombe_and_nlmeans_norm) norm = .015f / (2 * P + 1);
  const float cpw = d->central_pixel_weight * scale;
  wavelet_plan_t pl;
  make_plan_ex(&pl, d, roi_scale, width, height, wb_coeffs, pm, 1);
  b200_denoiseprofile_data_t dd = *d;
  dd.wavelet_color_mode = B200_DENOISE_RGB; /* the NLM path only has the RGB transforms */
  float *pre = malloc(16 * npx);
  if(!pre) return 1;
  vst_forward(&pl, &dd, in, pre, npx);
  const float norm2[4] = { 1.0f, 1.0f, 1.0f, 1.0f };
  orc_nlmeans_denoise(pre, out, width, height, scattering, scale, 1.0f, 1.0f, cpw, norm, P, K, 0, norm2);
  free(pre);
  vst_backward(&pl, &dd, out, npx);
  return 0;
}
