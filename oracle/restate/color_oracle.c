/* CPU restatement of the matrix + tone-curve colour conversion behind colorin and colorout.
 * TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src/colorprofiles/conversion.c
 *   dt_colorspaces_apply_conversion_hooked :744-760 -> _apply_matrix :593-682 -> _apply_target_curves :546-583
 * with dt_mat3x4_mul_vec4 (system/simd.h:188-197), _clamp_unit (conversion.c:536-543),
 * dt_ioppr_eval_trc / extrapolate_lut / eval_exp (colorprofiles/iop_profile.h:536-580), CLAMPS
 * (math/math.h:78), and the callers iop/colorin.c:711-734, iop/colorout.c:373-389.
 *
 * Two flavours of the same arithmetic:
 *   ORC_FP_STRICT    every multiply and add rounded on its own (C semantics, "ref-strict")
 *   ORC_FP_CONTRACT  multiply-adds fused the way the reference's release build fuses them on an
 *                    FMA-capable x86-64 (-ffp-contract=fast + target_clones x86-64-v3, gcc 13):
 *                    `r0*x; r1*y + out; r2*z + out` (simd.h:191-196) comes out as
 *                    fma(r2,z, fma(r0,x, r1*y)) -- GCC folds the FIRST product into the sum -- and
 *                    the LUT lerp l1*(1-f) + l2*f as fma(l1, 1-f, l2*f) in the target-curve pass
 *                    (conversion.c:546-583) but fma(l2, f, l1*(1-f)) inside the source-curve loop
 *                    (conversion.c:645-680).  Found by testing the
 *                    candidate orders against oracle/_ref/libref_fast.so, which pins this flavour.
 * powf is glibc's (flt32_math.h).
 */
#include "oracle_common.h"
#include "flt32_math.h"
#include <stdlib.h>

#define LUT_SAMPLES 0x10000 /* DT_CONVERSION_LUT_SAMPLES conversion.h:67 */
enum
{
  ORC_FP_STRICT = 0,
  ORC_FP_CONTRACT = 1
};

static inline float madd(float a, float b, float c, int fp)
{
  return fp == ORC_FP_CONTRACT ? fmaf(a, b, c) : a * b + c;
}

/* iop_profile.h:536-545 */
static inline float lut_lerp(const float *lut, float v, int fp, int decode_loop)
{
  const float scaled = v * (float)(LUT_SAMPLES - 1);
  const float ft = scaled > 0.0f ? (scaled < (float)(LUT_SAMPLES - 1) ? scaled : (float)(LUT_SAMPLES - 1)) : 0.0f;
  const int t = (ft < (float)(LUT_SAMPLES - 2)) ? (int)ft : LUT_SAMPLES - 2;
  const float f = ft - (float)t;
  const float l1 = lut[t], l2 = lut[t + 1];
  if(fp != ORC_FP_CONTRACT) return l1 * (1.0f - f) + l2 * f;
  /* which product GCC fuses differs between the two loops the lerp is inlined into */
  return decode_loop ? fmaf(l2, f, l1 * (1.0f - f)) : fmaf(l1, 1.0f - f, l2 * f);
}

/* iop_profile.h:559-562, 577-580 */
static inline float eval_trc(float x, const float *lut, const float coeff[3], int fp, int decode_loop)
{
  return (x < 1.0f) ? lut_lerp(lut, x, fp, decode_loop) : coeff[1] * f32m_powf(x * coeff[0], coeff[2]);
}

/* simd.h:188-197 on all four lanes; m is the row-major 3x3, lane 3 multiplies zeros */
static inline void mat_apply(const float m[9], const float in[4], float out[4], int fp)
{
  for(int i = 0; i < 4; i++)
  {
    const float a = i < 3 ? m[3 * i + 0] : 0.0f, b = i < 3 ? m[3 * i + 1] : 0.0f, c = i < 3 ? m[3 * i + 2] : 0.0f;
    float acc;
    if(fp == ORC_FP_CONTRACT)
      acc = fmaf(a, in[0], b * in[1]); /* GCC fuses the first product into the sum: fma(r0,x, r1*y) */
    else
      acc = a * in[0] + b * in[1];
    out[i] = madd(c, in[2], acc, fp);
  }
}

static inline float clamp01(float v) { return v > 1.0f ? 1.0f : (v < 0.0f ? 0.0f : v); } /* CLAMP, NaN passes */

/* lut_* : 3 x LUT_SAMPLES floats or NULL; a channel whose first entry is negative is linear.
 * Returns 0.  in may equal out. */
int orc_apply_matrix_conversion(const float *in, float *out, size_t width, size_t height, const float matrix[9],
                                const float clip_matrix[9], int has_clipping, const float *lut_source,
                                const float coeffs_source[9], const float *lut_target, const float coeffs_target[9],
                                int fp)
{
  const size_t npx = width * height;
  const float *ls[3] = { 0, 0, 0 }, *lt[3] = { 0, 0, 0 };
  int n_source = 0, n_target = 0;
  for(int k = 0; k < 3; k++)
  {
    if(lut_source) ls[k] = lut_source + (size_t)k * LUT_SAMPLES;
    if(lut_target) lt[k] = lut_target + (size_t)k * LUT_SAMPLES;
    if(lut_source && ls[k][0] >= 0.0f) n_source++;
    if(lut_target && lt[k][0] >= 0.0f) n_target++;
  }
  const int decode = lut_source && n_source > 0; /* conversion.c:610 */
  const int encode = lut_target && n_target > 0; /* conversion.c:611 */

#pragma omp parallel for schedule(static)
  for(size_t k = 0; k < npx; k++)
  {
    float px[4] = { in[4 * k], in[4 * k + 1], in[4 * k + 2], in[4 * k + 3] };
    if(decode)
    { /* conversion.c:652-664 */
      for(int c = 0; c < 3; c++)
        if(ls[c][0] >= 0.0f) px[c] = eval_trc(px[c], ls[c], coeffs_source + 3 * c, fp, 1);
      px[3] = 0.0f;
    }
    float v[4];
    mat_apply(matrix, px, v, fp);
    if(has_clipping)
    { /* conversion.c:536-543,625,669 */
      float c4[4] = { clamp01(v[0]), clamp01(v[1]), clamp01(v[2]), 0.0f };
      mat_apply(clip_matrix, c4, v, fp);
    }
    if(encode)
      for(int c = 0; c < 3; c++) /* conversion.c:546-583: a second pass upstream, pointwise all the same */
        if(lt[c][0] >= 0.0f) v[c] = eval_trc(v[c], lt[c], coeffs_target + 3 * c, fp, 0);
    out[4 * k] = v[0];
    out[4 * k + 1] = v[1];
    out[4 * k + 2] = v[2];
    out[4 * k + 3] = v[3];
  }
  return 0;
}

/* dt_iop_estimate_exp fit used by dt_ioppr_init_unbounded_coeffs (iop_profile.c:303-329) is host-side
 * set-up (commit_params); it is restated with the product's host code, see ansel_b200/iop/. */
