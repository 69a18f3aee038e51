/* oracle/_ref wrapper: the RGB <-> Lab glue the pipe inserts between modules of different colour space
 * (dt_colorspaces_apply_profile -> dt_ioppr_transform_matrix).  TEST INFRASTRUCTURE ONLY.
 *
 * colorprofiles/iop_profile.c needs lcms2 and the profile registry; oracle/Makefile cuts the three pixel loops
 * verbatim into oracle/_ref/gen_iop_profile.c:  :332-373 _apply_tonecurves, :376-420 _transform_rgb_to_lab_matrix,
 * :422-464 _transform_lab_to_rgb_matrix.  The per-pixel maths they call (dt_XYZ_to_Lab, dt_Lab_to_XYZ,
 * dt_mat3x4_mul_vec4) comes from the reference's unmodified headers.
 */
#include <glib.h>
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include "system/macros.h"
#include "system/mem_alloc.h"
#include "system/openmp.h"
#ifdef REF_STRICT
#define __DT_CLONE_TARGETS__
#else
#include "system/target_clones.h"
#endif
#include "system/simd.h"
#include "math/math.h"
#include "math/matrices.h"
#include "pixel/format.h"
#include "common/colorspaces_inline_conversions.h"
#include "colorprofiles/iop_profile.h"

#include "gen_iop_profile.c"

static void fill(dt_iop_order_iccprofile_info_t *p, const float m_in[9], const float m_out[9])
{
  memset(p, 0, sizeof(*p));
  for(int i = 0; i < 3; i++)
    for(int j = 0; j < 3; j++)
    {
      p->matrix_in[i][j] = m_in[3 * i + j];
      p->matrix_out[i][j] = m_out[3 * i + j];
      p->matrix_in_transposed[j][i] = m_in[3 * i + j];
      p->matrix_out_transposed[j][i] = m_out[3 * i + j];
    }
  p->nonlinearlut = 0;
  p->lutsize = 0x10000;
}
int ref_rgb_to_lab(const float *in, float *out, int width, int height, const float m_in[9], const float m_out[9])
{
  dt_iop_order_iccprofile_info_t *p = aligned_alloc(64, ((sizeof(*p) + 63) / 64) * 64);
  fill(p, m_in, m_out);
  _transform_rgb_to_lab_matrix(in, out, width, height, p);
  free(p);
  return 0;
}
int ref_lab_to_rgb(const float *in, float *out, int width, int height, const float m_in[9], const float m_out[9])
{
  dt_iop_order_iccprofile_info_t *p = aligned_alloc(64, ((sizeof(*p) + 63) / 64) * 64);
  fill(p, m_in, m_out);
  _transform_lab_to_rgb_matrix(in, out, width, height, p);
  free(p);
  return 0;
}

/* the same two loops for a profile with tone curves: luts = 3 x 65536 floats per direction (lut[k][0] < 0 marks a
 * linear channel), coeffs = unbounded_coeffs_{in,out}; nonlinearlut as dt_ioppr_init_unbounded_coeffs counts it */
static void fill_curves(dt_iop_order_iccprofile_info_t *p, const float *luts_in, const float *luts_out, const float co_in[9], const float co_out[9])
{
  p->nonlinearlut = 0;
  for(int k = 0; k < 3; k++)
  {
    p->lut_in[k] = (float *)luts_in + (size_t)k * 0x10000;
    p->lut_out[k] = (float *)luts_out + (size_t)k * 0x10000;
    if(p->lut_in[k][0] >= 0.0f) p->nonlinearlut++;
    for(int j = 0; j < 3; j++)
    {
      p->unbounded_coeffs_in[k][j] = co_in[3 * k + j];
      p->unbounded_coeffs_out[k][j] = co_out[3 * k + j];
    }
  }
}
int ref_rgb_to_lab_trc(const float *in, float *out, int width, int height, const float m_in[9], const float m_out[9], const float *luts_in,
                       const float *luts_out, const float co_in[9], const float co_out[9])
{
  dt_iop_order_iccprofile_info_t *p = aligned_alloc(64, ((sizeof(*p) + 63) / 64) * 64);
  fill(p, m_in, m_out);
  fill_curves(p, luts_in, luts_out, co_in, co_out);
  _transform_rgb_to_lab_matrix(in, out, width, height, p);
  free(p);
  return 0;
}
int ref_lab_to_rgb_trc(const float *in, float *out, int width, int height, const float m_in[9], const float m_out[9], const float *luts_in,
                       const float *luts_out, const float co_in[9], const float co_out[9])
{
  dt_iop_order_iccprofile_info_t *p = aligned_alloc(64, ((sizeof(*p) + 63) / 64) * 64);
  fill(p, m_in, m_out);
  fill_curves(p, luts_in, luts_out, co_in, co_out);
  _transform_lab_to_rgb_matrix(in, out, width, height, p);
  free(p);
  return 0;
}
