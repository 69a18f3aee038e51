/* oracle/_ref wrapper: blending of a module's output over its input with a mask, raw space (one sample per site: the modules in front of the
 * demosaicer).  TEST INFRASTRUCTURE ONLY.
 *
 * oracle/Makefile cuts verbatim into oracle/_ref/gen_blend_raw*.c:
 *     develop/blend.h   :52-193 197-237 329-329, develop/develop.h :120-145      the enums and dt_develop_blend_params_t (as for ref_blend.c)
 *     develop/blend.c   :626-655    _develop_blend_process_mask_tone_curve
 *     develop/blends/blendif_raw.c :32-409   dt_develop_blendif_raw_make_mask, the 16 blend operators, _choose_blend_func, dt_develop_blendif_raw_blend
 * ref_blend_raw_process() below is dt_develop_blend_process (develop/blend.c:657-860) for blend_cst == DEVELOP_BLEND_CS_RAW without feathering,
 * blur and detail refinement, with the form mask handed in by the caller: the twin of ref_blend_process() in ref_blend.c on float buffers of one
 * channel.  The parametric mask has no channels in this space; a block that asks for one still takes the seeded path of :742-752, as there.
 */
#include "ref_piece.h"
#undef DT_DEV_PIXELPIPE_DISPLAY_MASK /* ref_piece.h supplies it as a macro; here the enum of develop/develop.h is cut in */
#include <stdio.h>
#include "math/matrices.h"
#include "math/openmp_maths.h"
typedef char dt_dev_operation_t[20]; /* history/history.h */
typedef struct ref_blend_raw_piece_t
{
  dt_iop_roi_t roi_in, roi_out;
  struct { int channels; } dsc_in;
  void *blendop_data;
} ref_blend_raw_piece_t;
typedef struct ref_blend_raw_pipe_t { int mask_display; } ref_blend_raw_pipe_t;
#define dt_dev_pixelpipe_iop_t ref_blend_raw_piece_t
#define dt_dev_pixelpipe_t ref_blend_raw_pipe_t
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
#include "gen_blend_raw_a.c" /* enums, parameters */
#include "gen_blend_raw_b.c" /* make_mask, the blend operators, the blend, _develop_blend_process_mask_tone_curve */

typedef struct ref_blend_raw_params_t
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
} ref_blend_raw_params_t;

/* in: iw x ih floats, out: ow x oh floats at (xoffs, yoffs) inside it, blended in place; form / mask_out as in ref_blend_process() */
int ref_blend_raw_process(const float *in, float *out, int iw, int ih, int ow, int oh, int xoffs, int yoffs, const ref_blend_raw_params_t *bp,
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
  if(d.blend_cst != DEVELOP_BLEND_CS_RAW) return -1;
  if(d.feathering_radius > 0.1f || d.blur_radius > 0.1f || d.details != 0.0f) return -1;
  ref_blend_raw_piece_t piece;
  memset(&piece, 0, sizeof(piece));
  piece.roi_in = (dt_iop_roi_t){ 0, 0, iw, ih, 1.0 };
  piece.roi_out = (dt_iop_roi_t){ xoffs, yoffs, ow, oh, 1.0 };
  piece.dsc_in.channels = 1;
  piece.blendop_data = &d;
  ref_blend_raw_pipe_t pipe = { (int)bp->mask_display };
  const size_t buffsize = (size_t)ow * oh;
  int parametric = 0; /* parametric_used, :290-312: the RGB channel set outside Lab */
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
    dt_develop_blendif_raw_make_mask(&piece, in, out, mask); /* :808-811 */
    const int tone_curve = fabsf(d.contrast) >= 0.01f || fabsf(d.brightness) >= 0.01f; /* :432-435, :463-466 */
    if(tone_curve && opacity > 1e-4f) _develop_blend_process_mask_tone_curve(mask, buffsize, d.contrast, d.brightness, opacity);
  }
  dt_develop_blendif_raw_blend(&pipe, &piece, in, out, mask, DT_DEV_PIXELPIPE_DISPLAY_NONE); /* :893-896 */
  if(mask_out) memcpy(mask_out, mask, buffsize * sizeof(float));
  free(mask);
  return 0;
}
