/* oracle/_ref wrapper for the reference colour-conversion apply path.  TEST INFRASTRUCTURE ONLY.
 *
 * Textually includes the reference's own, unmodified
 *   /root/reference/src/colorprofiles/conversion.c
 * to reach  dt_colorspaces_apply_conversion_hooked() :744-760 -> _apply_matrix() :593-682 and
 * _apply_target_curves() :546-583 (with dt_mat3x4_mul_vec4 system/simd.h:188-197 and
 * dt_ioppr_eval_trc colorprofiles/iop_profile.h:577-580), i.e. the arithmetic behind
 * colorin.c:711-734 and colorout.c:373-389.
 *
 * The profile-building half of conversion.c (dt_colorspaces_prepare_conversion, lcms2) is
 * compiled but never entered: the harness fills dt_colorspaces_conversion_t directly with the
 * matrix / LUT / extrapolation coefficients that half would have produced.  Its external
 * references are resolved to aborting stubs below.
 */
#include <glib.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef REF_STRICT
#define __DT_CLONE_TARGETS__
#endif

#include "colorprofiles/conversion.c"

/* colorin.c:690-709 is a static function of the iop file; the hook signature is what
 * conversion.c sees (dt_colorspaces_conversion_hook_t).  Not exercised by the pinned configs. */

/* Plain-C entry point: run the reference's matrix conversion over `npixels` RGBA pixels.
 *   matrix / clip_matrix : row-major 3x3 (source -> target | source -> clip, clip -> target)
 *   lut_source/lut_target: 3 x 65536 floats or NULL; lut[c][0] < 0 marks a linear channel
 *   coeffs_*             : 3 x {a, b, c} of eval_exp (iop_profile.h:559-562)
 * in and out must be 64-byte aligned (the reference asserts it, conversion.c:620). */
int ref_apply_matrix_conversion(const float *in, float *out, size_t width, size_t height, const float matrix[9],
                                const float clip_matrix[9], int has_clipping, const float *lut_source,
                                const float coeffs_source[9], const float *lut_target,
                                const float coeffs_target[9])
{
  dt_colorspaces_conversion_t c;
  memset(&c, 0, sizeof(c));
  c.magic = DT_CONVERSION_MAGIC_LIVE;
  c.is_matrix = TRUE;
  c.has_clipping = has_clipping;
  for(int i = 0; i < 3; i++)
    for(int j = 0; j < 3; j++)
    {
      c.matrix[i][j] = matrix[3 * i + j];
      c.clip_matrix[i][j] = clip_matrix ? clip_matrix[3 * i + j] : 0.0f;
    }
  for(int k = 0; k < 3; k++)
  {
    c.lut_source[k] = lut_source ? (float *)lut_source + (size_t)k * DT_CONVERSION_LUT_SAMPLES : NULL;
    c.lut_target[k] = lut_target ? (float *)lut_target + (size_t)k * DT_CONVERSION_LUT_SAMPLES : NULL;
    for(int j = 0; j < 3; j++)
    {
      c.coeffs_source[k][j] = coeffs_source ? coeffs_source[3 * k + j] : 0.0f;
      c.coeffs_target[k][j] = coeffs_target ? coeffs_target[3 * k + j] : 0.0f;
    }
    if(lut_source && c.lut_source[k][0] >= 0.0f) c.nonlinear_source++;
    if(lut_target && c.lut_target[k][0] >= 0.0f) c.nonlinear_target++;
  }
  dt_colorspaces_apply_conversion_hooked(&c, in, out, width, height, NULL);
  return 0;
}

int ref_conversion_lut_samples(void) { return DT_CONVERSION_LUT_SAMPLES; }
