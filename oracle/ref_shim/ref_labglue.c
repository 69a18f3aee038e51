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
