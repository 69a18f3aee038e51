/* oracle/_ref wrapper: blending of a module's output over its input with a mask, display-referred RGB space (the modules behind the tone
 * mapping).  TEST INFRASTRUCTURE ONLY.
 *
 * oracle/Makefile cuts verbatim into oracle/_ref/gen_blend_lab_*.c:
 *     develop/blend.h   :52-193 197-237 329-329, develop/develop.h :120-145      the enums and dt_develop_blend_params_t (as for ref_blend.c)
 *     develop/blend.c   :214-260    dt_develop_blendif_process_parameters      :321-353  dt_develop_blendif_init_masking_profile
 *                       :626-655    _develop_blend_process_mask_tone_curve
 *     colorprofiles/iop_profile.h :637-654  dt_ioppr_get_rgb_matrix_luminance
 *     develop/blends/blendif_rgb_hsl.c :34-1007   _blendif_compute_factor, the gray / R / G / B / H / S / L channels,
 *                                              _blendif_combine_channels, dt_develop_blendif_rgb_hsl_make_mask, the 27 blend operators, _choose_blend_func
 *                       :1197-1288  _copy_mask, dt_develop_blendif_rgb_hsl_blend
 * common/colorspaces_inline_conversions.h (dt_RGB_2_HSL, dt_RGB_2_HSV and back) is included unmodified.
 * Not cut: the GUI's channel display (:1010-1195), an aborting stub.
 * ref_blend_rgb_hsl_process() below is dt_develop_blend_process (develop/blend.c:657-860) for blend_cst == DEVELOP_BLEND_CS_RGB_DISPLAY
 * without feathering, blur and detail refinement, with the form mask handed in by the caller: the twin of ref_blend_process() in ref_blend.c.
 */
#include "ref_piece.h"
#undef DT_DEV_PIXELPIPE_DISPLAY_MASK /* ref_piece.h supplies it as a macro; here the enum of develop/develop.h is cut in */
#include <stdio.h>
#include "math/matrices.h"
#include "common/colorspaces_inline_conversions.h"
#include "math/openmp_maths.h"
typedef char dt_dev_operation_t[20]; /* history/history.h */
typedef struct dt_iop_order_iccprofile_info_t
{ /* the members the cut lines read (colorprofiles/iop_profile.h) */
  dt_colormatrix_t matrix_in, matrix_out, matrix_out_transposed;
  float *lut_in[3];
  float unbounded_coeffs_in[3][3];
  int lutsize, nonlinearlut;
} dt_iop_order_iccprofile_info_t;
static inline void _apply_trc(const float *rgb, float *out, float *const lut[3], const float c[3][3], int lutsize)
{
  (void)rgb; (void)out; (void)lut; (void)c; (void)lutsize;
  abort(); /* linear work profiles only */
}
#define dt_ioppr_get_rgb_matrix_luminance ref_rgb_hsl_matrix_luminance
static dt_iop_order_iccprofile_info_t g_blend_rgb_hsl_profile;
#define dt_ioppr_get_pipe_current_profile_info(module, pipe) (&g_blend_rgb_hsl_profile)
#define dt_ioppr_get_iop_work_profile_info(module, iop) (&g_blend_rgb_hsl_profile)
#define dt_develop_blendif_init_masking_profile ref_blend_rgb_hsl_init_masking_profile

typedef struct ref_blend_rgb_hsl_piece_t
{
  dt_iop_roi_t roi_in, roi_out;
  struct { int channels; } dsc_in;
  void *blendop_data;
} ref_blend_rgb_hsl_piece_t;
typedef struct ref_blend_rgb_hsl_pipe_t { int mask_display; } ref_blend_rgb_hsl_pipe_t;
#define dt_dev_pixelpipe_iop_t ref_blend_rgb_hsl_piece_t
#define dt_dev_pixelpipe_t ref_blend_rgb_hsl_pipe_t
static void dt_iop_image_fill(float *const buf, const float v, const size_t w, const size_t h, const size_t ch)
{
  for(size_t k = 0; k < w * h * ch; k++) buf[k] = v;
}
static void dt_iop_image_mul_const(float *const buf, const float v, const size_t w, const size_t h, const size_t ch)
{
  for(size_t k = 0; k < w * h * ch; k++) buf[k] *= v;
}
static void dt_iop_image_copy(float *const out, const float *const in, const size_t n) { memcpy(out, in, n * sizeof(float)); }
static float *dt_pixelpipe_cache_alloc_align_float_cache(size_t n, int id) { (void)id; return aligned_alloc(64, ((n * sizeof(float) + 63) / 64) * 64); }
#define dt_pixelpipe_cache_free_align(p) free((void *)(p))
#define dt_develop_blendif_process_parameters ref_blend_rgb_hsl_process_parameters /* ref_blend.c holds the same lines under their own name */
#include "gen_blend_hsl_a.c" /* enums, parameters, luminance, dt_develop_blendif_process_parameters, dt_develop_blendif_init_masking_profile */
#include "gen_blend_hsl_b.c" /* the channels of the parametric mask, make_mask, the blend operators */
static void _display_channel(const float *const restrict a, float *const restrict b, const float *const restrict mask, const size_t stride,
                             const int channel, const float *const restrict boost_factors, const dt_iop_order_iccprofile_info_t *const profile)
{
  (void)a; (void)b; (void)mask; (void)stride; (void)channel; (void)boost_factors; (void)profile;
  abort(); /* :1010-1195 not cut: a GUI request */
}
#include "gen_blend_hsl_c.c" /* _copy_mask, dt_develop_blendif_rgb_hsl_blend, _develop_blend_process_mask_tone_curve */

typedef struct ref_blend_rgb_hsl_params_t
{ /* == ref_blend_params_t of ref_blend.c */
  uint32_t mask_mode;
  int32_t blend_cst;
  uint32_t blend_mode;
  float blend_parameter, opacity;
  uint32_t mask_combine, blendif;
  float feathering_radius;
  uint32_t feathering_guide;
  float blur_radius, contrast, brightness, details;
  float blendif_parameters[4 * DEVELOP_BLENDIF_SIZE], blendif_boost_factors[DEVELOP_BLENDIF_SIZE];
  int32_t raster_used, drawn_used;
  float luminance[3];
  int32_t profile_nonlinear;
  uint32_t mask_display;
  float matrix_in[9];
} ref_blend_rgb_hsl_params_t;

/* dt_develop_blend_process(), develop/blend.c:657-860, for the display-referred RGB space; arguments as ref_blend_process().  Returns 0, or -1 for what the
 * wrapper does not reach (feathering, blur, detail refinement, other colour spaces). */
int ref_blend_rgb_hsl_process(const float *in, float *out, int iw, int ih, int ow, int oh, int xoffs, int yoffs, const ref_blend_rgb_hsl_params_t *bp,
                          const float *form, float *mask_out)
{
  dt_develop_blend_params_t d;
  memset(&d, 0, sizeof(d));
  d.mask_mode = bp->mask_mode;
  d.blend_cst = bp->blend_cst;
  d.blend_mode = bp->blend_mode;
  d.blend_parameter = bp->blend_parameter;
  d.opacity = bp->opacity;
  d.mask_combine = bp->mask_combine;
  d.blendif = bp->blendif;
  d.feathering_radius = bp->feathering_radius;
  d.feathering_guide = bp->feathering_guide;
  d.blur_radius = bp->blur_radius;
  d.contrast = bp->contrast;
  d.brightness = bp->brightness;
  d.details = bp->details;
  memcpy(d.blendif_parameters, bp->blendif_parameters, sizeof(d.blendif_parameters));
  memcpy(d.blendif_boost_factors, bp->blendif_boost_factors, sizeof(d.blendif_boost_factors));
  if(!(d.mask_mode & DEVELOP_MASK_ENABLED)) return 0; /* :673 */
  if(d.blend_cst != DEVELOP_BLEND_CS_RGB_DISPLAY || bp->profile_nonlinear) return -1;
  memset(&g_blend_rgb_hsl_profile, 0, sizeof(g_blend_rgb_hsl_profile));
  for(int r = 0; r < 3; r++)
    for(int k = 0; k < 3; k++) g_blend_rgb_hsl_profile.matrix_in[r][k] = bp->matrix_in[3 * r + k];
  for(int k = 0; k < 3; k++) g_blend_rgb_hsl_profile.matrix_in[1][k] = bp->luminance[k]; /* what the gray channel reads */
  if(d.feathering_radius > 0.1f || d.blur_radius > 0.1f || d.details != 0.0f) return -1;
  ref_blend_rgb_hsl_piece_t piece;
  memset(&piece, 0, sizeof(piece));
  piece.roi_in = (dt_iop_roi_t){ 0, 0, iw, ih, 1.0 };
  piece.roi_out = (dt_iop_roi_t){ xoffs, yoffs, ow, oh, 1.0 };
  piece.dsc_in.channels = 4;
  piece.blendop_data = &d;
  ref_blend_rgb_hsl_pipe_t pipe = { (int)bp->mask_display };
  const size_t buffsize = (size_t)ow * oh;
  int parametric = 0; /* parametric_used, :290-312 */
  if(d.mask_mode & DEVELOP_MASK_PARAMETRIC)
    for(uint32_t ch = 0; ch < DEVELOP_BLENDIF_SIZE; ch++)
    {
      const uint32_t bit = 1u << ch;
      if(!(DEVELOP_BLENDIF_RGB_MASK & bit) || !(d.blendif & bit)) continue;
      const float *c = &d.blendif_parameters[ch * 4];
      if(fabsf(c[0]) > 1e-6f || fabsf(c[1]) > 1e-6f || fabsf(c[2] - 1.0f) > 1e-6f || fabsf(c[3] - 1.0f) > 1e-6f) parametric = 1;
    }
  const int raster = bp->raster_used && form, drawn = bp->drawn_used && form;
  const float opacity = fminf(fmaxf(d.opacity / 100.0f, 0.0f), 1.0f);
  float *mask = aligned_alloc(64, ((buffsize * sizeof(float) + 63) / 64) * 64);
  if(!raster && !drawn && !parametric)
    dt_iop_image_fill(mask, opacity, ow, oh, 1); /* :731-735 */
  else if(raster && !drawn && !parametric)
  { /* :736-741 */
    memcpy(mask, form, buffsize * sizeof(float));
    dt_iop_image_mul_const(mask, opacity, ow, oh, 1);
  }
  else
  {
    if(!raster && !drawn)
      dt_iop_image_fill(mask, (d.mask_combine & DEVELOP_COMBINE_INCL) ? 0.0f : 1.0f, ow, oh, 1); /* :744-752 */
    else
      memcpy(mask, form, buffsize * sizeof(float));
    dt_develop_blendif_rgb_hsl_make_mask(&pipe, &piece, in, out, mask); /* :800-803 */
    const int tone_curve = fabsf(d.contrast) >= 0.01f || fabsf(d.brightness) >= 0.01f; /* :432-435, :463-466 */
    if(tone_curve && opacity > 1e-4f) _develop_blend_process_mask_tone_curve(mask, buffsize, d.contrast, d.brightness, opacity);
  }
  dt_develop_blendif_rgb_hsl_blend(&pipe, &piece, in, out, mask, DT_DEV_PIXELPIPE_DISPLAY_NONE); /* :885-888 */
  if(mask_out) memcpy(mask_out, mask, buffsize * sizeof(float));
  free(mask);
  return 0;
}
