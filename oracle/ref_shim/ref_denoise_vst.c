/* oracle/_ref wrapper: the variance-stabilising transforms and wavelet-threshold arithmetic of
 * profiled denoise.  TEST INFRASTRUCTURE ONLY.
 *
 * iop/denoiseprofile.c is one translation unit with its GTK GUI and cannot be compiled whole here.
 * oracle/Makefile cuts its pixel functions out verbatim (oracle/ref_shim/slice.py) into
 * oracle/_ref/gen_denoiseprofile.c:
 *     :109-143   constants and enums            :352-371   dt_iop_denoiseprofile_data_t
 *     :852-1089  precondition/backtransform{,_v2,_Y0U0V0}
 *     :1098-1286 compute_wb_factors, invert_matrix, set_up_conversion_matrices, variance_stabilizing_xform
 * and this file compiles that cut after declaring the few names it expects.
 */
#include <glib.h>
#include <math.h>
#include <string.h>
#include "system/mem_alloc.h"
#include "system/openmp.h"
#ifdef REF_STRICT
#define __DT_CLONE_TARGETS__
#else
#include "system/target_clones.h"
#endif
#include "system/simd.h"
#include "math/matrices.h"
#include "pixel/format.h"

typedef struct dt_draw_curve_t dt_draw_curve_t;
typedef struct dt_dev_pixelpipe_iop_t
{
  dt_iop_buffer_dsc_t dsc_in;
} dt_dev_pixelpipe_iop_t;
#define DT_PIXEL_APPLY_DPI(x) (x)

#include "gen_denoiseprofile.c"

/* plain-C entry points over RGBA float buffers ----------------------------------------------- */
void ref_dn_precondition(const float *in, float *buf, int wd, int ht, const float a[4], const float b[4])
{
  dt_aligned_pixel_t aa = { a[0], a[1], a[2], a[3] }, bb = { b[0], b[1], b[2], b[3] };
  precondition(in, buf, wd, ht, aa, bb);
}
void ref_dn_backtransform(float *buf, int wd, int ht, const float a[4], const float b[4])
{
  dt_aligned_pixel_t aa = { a[0], a[1], a[2], a[3] }, bb = { b[0], b[1], b[2], b[3] };
  backtransform(buf, wd, ht, aa, bb);
}
void ref_dn_precondition_v2(const float *in, float *buf, int wd, int ht, float a, const float p[4], float b, const float wb[4])
{
  dt_aligned_pixel_t pp = { p[0], p[1], p[2], p[3] }, ww = { wb[0], wb[1], wb[2], wb[3] };
  precondition_v2(in, buf, wd, ht, a, pp, b, ww);
}
void ref_dn_backtransform_v2(float *buf, int wd, int ht, float a, const float p[4], float b, float bias, const float wb[4])
{
  dt_aligned_pixel_t pp = { p[0], p[1], p[2], p[3] }, ww = { wb[0], wb[1], wb[2], wb[3] };
  backtransform_v2(buf, wd, ht, a, pp, b, bias, ww);
}
void ref_dn_precondition_Y0U0V0(const float *in, float *buf, int wd, int ht, float a, const float p[4], float b, const float m[12])
{
  dt_aligned_pixel_t pp = { p[0], p[1], p[2], p[3] };
  dt_colormatrix_t M = { { 0 } };
  for(int i = 0; i < 3; i++)
    for(int j = 0; j < 4; j++) M[i][j] = m[4 * i + j];
  precondition_Y0U0V0(in, buf, wd, ht, a, pp, b, M);
}
void ref_dn_backtransform_Y0U0V0(float *buf, int wd, int ht, float a, const float p[4], float b, float bias, const float wb[4],
                                 const float m[12])
{
  dt_aligned_pixel_t pp = { p[0], p[1], p[2], p[3] }, ww = { wb[0], wb[1], wb[2], wb[3] };
  dt_colormatrix_t M = { { 0 } };
  for(int i = 0; i < 3; i++)
    for(int j = 0; j < 4; j++) M[i][j] = m[4 * i + j];
  backtransform_Y0U0V0(buf, wd, ht, a, pp, b, bias, ww, M);
}
/* wb: in/out; toY, toRGB: 12 floats each (3 rows of 4) */
void ref_dn_conversion_matrices(float toY[12], float toRGB[12], const float wb[4])
{
  dt_colormatrix_t A = { { 1.0f / 3.0f, 1.0f / 3.0f, 1.0f / 3.0f }, { 0.5f, 0.0f, -0.5f }, { 0.25f, -0.5f, 0.25f } };
  dt_colormatrix_t B = { { 0 } };
  dt_aligned_pixel_t ww = { wb[0], wb[1], wb[2], wb[3] };
  set_up_conversion_matrices(A, B, ww);
  for(int i = 0; i < 3; i++)
    for(int j = 0; j < 4; j++)
    {
      toY[4 * i + j] = A[i][j];
      toRGB[4 * i + j] = B[i][j];
    }
}
void ref_dn_wb_factors(float wb[4], int fix_norm, int adaptive, const float coeffs[4], const float pm[4], const float weights[4])
{
  dt_iop_denoiseprofile_data_t d;
  memset(&d, 0, sizeof(d));
  d.fix_anscombe_and_nlmeans_norm = fix_norm;
  d.wb_adaptive_anscombe = adaptive;
  dt_dev_pixelpipe_iop_t piece;
  memset(&piece, 0, sizeof(piece));
  for(int k = 0; k < 4; k++)
  {
    piece.dsc_in.temperature.coeffs[k] = coeffs[k];
    piece.dsc_in.processed_maximum[k] = pm[k];
  }
  dt_aligned_pixel_t w = { 0 }, ww = { weights[0], weights[1], weights[2], weights[3] };
  compute_wb_factors(w, &d, &piece, ww);
  for(int k = 0; k < 4; k++) wb[k] = w[k];
}
/* force: 6 x 7 floats */
void ref_dn_thresholds(float thrs[4], int scale, int max_scale, size_t npixels, const float sum_y2[4], int color_mode,
                       const float *force)
{
  dt_iop_denoiseprofile_data_t d;
  memset(&d, 0, sizeof(d));
  d.wavelet_color_mode = color_mode;
  memcpy(d.force, force, sizeof(d.force));
  dt_aligned_pixel_t t = { 0 };
  variance_stabilizing_xform(t, scale, max_scale, npixels, sum_y2, &d);
  for(int k = 0; k < 4; k++) thrs[k] = t[k];
}
