/* CPU restatement of the RGB <-> Lab glue and of the denoise (non-local means) iop's parameter derivation.
 * TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src/colorprofiles/iop_profile.c _transform_rgb_to_lab_matrix :376-420 and
 * _transform_lab_to_rgb_matrix :422-464, _apply_tonecurves :332-373 for profiles with tone curves
 * (colorprofiles/iop_profile.h extrapolate_lut :536-545, eval_exp :559-562, dt_ioppr_eval_trc :577-580), common/
 * colorspaces_inline_conversions.h cbrt_5f :51-56, cbrta_halleyf :59-64, lab_f :67-72, d50 :75, dt_XYZ_to_Lab
 * :78-86, lab_f_inv :89-94, dt_Lab_to_XYZ :98-106, system/simd.h dt_mat3x4_mul_vec4 :188-197;
 * and src/iop/nlmeans.c process_cpu :416-456.  Pinned bit-for-bit against those functions cut verbatim
 * (oracle/_ref: ref_labglue.c, ref_nlm.c).
 */
#include "oracle_common.h"
#include "b200iop.h"
#include "flt32_math.h"
#include <string.h>

static const float d50[3] = { 0.9642f, 1.0f, 0.8249f };

static inline float cbrt_5f(float f)
{
  uint32_t u;
  memcpy(&u, &f, 4);
  u = u / 3 + 709921077u;
  memcpy(&f, &u, 4);
  return f;
}
static inline float cbrta_halleyf(float a, float R)
{
  const float a3 = a * a * a;
  return a * (a3 + R + R) / (a3 + a3 + R);
}
static inline float lab_f(float x)
{
  const float epsilon = 216.0f / 24389.0f, kappa = 24389.0f / 27.0f;
  return (x > epsilon) ? cbrta_halleyf(cbrt_5f(x), x) : (kappa * x + 16.0f) / 116.0f;
}
static inline float lab_f_inv(float x)
{
  const float epsilon = 0.20689655172413796f, kappa = 24389.0f / 27.0f;
  return (x > epsilon) ? x * x * x : (116.0f * x - 16.0f) / kappa;
}
/* dt_mat3x4_mul_vec4 with the transposed matrix: out = col0*x; out = col1*y + out; out = col2*z + out */
static inline void mat(const float m[9], const float v[3], float o[3])
{
  for(int r = 0; r < 3; r++)
  {
    float acc = m[3 * r + 0] * v[0];
    acc = m[3 * r + 1] * v[1] + acc;
    acc = m[3 * r + 2] * v[2] + acc;
    o[r] = acc;
  }
}

/* RGB -> Lab; lane 3 is not written by the reference: the caller's output buffer keeps its own */
int orc_rgb_to_lab(const float *in, float *out, int width, int height, const float matrix_in[9])
{
  const size_t n = (size_t)width * height;
#pragma omp parallel for
  for(size_t k = 0; k < n; k++)
  {
    float xyz[3], f[3];
    mat(matrix_in, in + 4 * k, xyz);
    for(int i = 0; i < 3; i++) f[i] = lab_f(xyz[i] / d50[i]);
    out[4 * k + 0] = 116.0f * f[1] - 16.0f;
    out[4 * k + 1] = 500.0f * (f[0] - f[1]);
    out[4 * k + 2] = 200.0f * (f[1] - f[2]);
  }
  return 0;
}
/* Lab -> RGB; lane 3 = the input's alpha (:437,442) */
int orc_lab_to_rgb(const float *in, float *out, int width, int height, const float matrix_out[9])
{
  const size_t n = (size_t)width * height;
#pragma omp parallel for
  for(size_t k = 0; k < n; k++)
  {
    const float *lab = in + 4 * k;
    const float alpha = lab[3];
    const float fy = (lab[0] + 16.0f) / 116.0f;
    const float fx = lab[1] / 500.0f + fy;
    const float fz = fy - lab[2] / 200.0f;
    const float f[3] = { fx, fy, fz };
    float xyz[3];
    for(int c = 0; c < 3; c++) xyz[c] = d50[c] * lab_f_inv(f[c]);
    mat(matrix_out, xyz, out + 4 * k);
    out[4 * k + 3] = alpha;
  }
  return 0;
}

/* dt_ioppr_eval_trc(), iop_profile.h:536-580 */
#define GLUE_LUT 0x10000
static inline float glue_eval_trc(float x, const float *lut, const float co[3])
{
  if(!(x < 1.0f)) return co[1] * f32m_powf(x * co[0], co[2]);
  const float scaled = x * (float)(GLUE_LUT - 1);
  const float ft = scaled > 0.0f ? (scaled < (float)(GLUE_LUT - 1) ? scaled : (float)(GLUE_LUT - 1)) : 0.0f; /* CLAMPS */
  const int t = ft < (float)(GLUE_LUT - 2) ? (int)ft : GLUE_LUT - 2;
  const float f = ft - (float)t;
  return lut[t] * (1.0f - f) + lut[t + 1] * f;
}
/* RGB -> Lab through lut_in (:388-405).  _apply_tonecurves writes only the channels that have a curve into the
 * output buffer and the matrix loop then reads that buffer: a channel without a curve, and lane 3, are whatever the
 * output held -- in place (how the pipe calls it) that is the pixel itself, which is what this restatement returns. */
int orc_rgb_to_lab_trc(const float *in, float *out, int width, int height, const float matrix_in[9], const float *luts_in, const float co_in[9])
{
  const size_t n = (size_t)width * height;
#pragma omp parallel for
  for(size_t k = 0; k < n; k++)
  {
    float rgb[3], xyz[3], f[3];
    for(int c = 0; c < 3; c++)
    {
      const float *lut = luts_in + (size_t)c * GLUE_LUT;
      rgb[c] = lut[0] >= 0.0f ? glue_eval_trc(in[4 * k + c], lut, co_in + 3 * c) : in[4 * k + c];
    }
    mat(matrix_in, rgb, xyz);
    for(int i = 0; i < 3; i++) f[i] = lab_f(xyz[i] / d50[i]);
    out[4 * k + 0] = 116.0f * f[1] - 16.0f;
    out[4 * k + 1] = 500.0f * (f[0] - f[1]);
    out[4 * k + 2] = 200.0f * (f[1] - f[2]);
  }
  return 0;
}
/* Lab -> RGB, then lut_out in place on the channels that have a curve (:456-463) */
int orc_lab_to_rgb_trc(const float *in, float *out, int width, int height, const float matrix_out[9], const float *luts_out, const float co_out[9])
{
  orc_lab_to_rgb(in, out, width, height, matrix_out);
  const size_t n = (size_t)width * height;
#pragma omp parallel for
  for(size_t k = 0; k < n; k++)
    for(int c = 0; c < 3; c++)
    {
      const float *lut = luts_out + (size_t)c * GLUE_LUT;
      if(lut[0] >= 0.0f) out[4 * k + c] = glue_eval_trc(out[4 * k + c], lut, co_out + 3 * c);
    }
  return 0;
}

int orc_nlmeans_denoise(const float *inbuf, float *outbuf, int width, int height, float scattering, float scale, float luma,
                        float chroma, float center_weight, float sharpness, int radius, int search_radius, int decimate,
                        const float norm[4]);
/* iop/nlmeans.c process_cpu :416-456; decimate = thumbnail pipe or a pipe that feeds the preview */
int orc_nlmeans_iop(const float *in, float *out, int width, int height, const b200_nlmeans_data_t *d, double roi_scale, int decimate,
                    int mask_display)
{
  const float scale = fmin(roi_scale, 2.0f);
  const int P = ceilf(d->radius * scale);
  const int K = ceilf(7 * scale);
  const float sharpness = 3000.0f / (1.0f + d->strength);
  const float max_L = 120.0f, max_C = 512.0f;
  const float nL = 1.0f / max_L, nC = 1.0f / max_C;
  const float norm2[4] = { nL * nL, nC * nC, nC * nC, 1.0f };
  const int rc = orc_nlmeans_denoise(in, out, width, height, 0, scale, d->luma, d->chroma, -1, sharpness, P, K, decimate, norm2);
  if(rc) return rc;
  if(mask_display & 1)
    for(size_t k = 3; k < (size_t)4 * width * height; k += 4) out[k] = in[k];
  return 0;
}
