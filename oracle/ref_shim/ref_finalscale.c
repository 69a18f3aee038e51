/* oracle/_ref wrapper: finalscale (the export's final resampling) = dt_iop_clip_and_zoom_roi ->
 * dt_interpolation_resample_roi.  TEST INFRASTRUCTURE ONLY.
 *
 * pixel/interpolation.c carries the OpenCL host code and the configuration lookups; oracle/Makefile cuts verbatim into
 * oracle/_ref/gen_interpolation.c:
 *     pixel/interpolation.h :37-66    interpolation types, struct dt_interpolation
 *     pixel/interpolation.c :51-70    border modes                 :88-165   _clip, _prepare_tap_boundaries
 *                           :175-314  the three tap generators and the interpolator table
 *                           :320-387  up/downsampling kernels       :710-1027 _prepare_resampling_plan,
 *                                                                             _interpolation_resample_plain
 * iop/finalscale.c process() :117-131 zeroes the origins of both ROIs and calls the resampler with the user's
 * interpolator (plugins/lighttable/export/pixel_interpolator; default mitchell): restated in ref_finalscale().
 */
#include "ref_piece.h"
#include "system/mem_alloc.h"
#include <inttypes.h>
#include <stddef.h>
#include <sys/types.h>
static inline void *dt_pixelpipe_cache_alloc_align_cache(size_t size, int id) { return aligned_alloc(64, ((size + 63) / 64) * 64); }
static inline void dt_pixelpipe_cache_free_align(void *p) { free(p); }
#define dt_omploop_sfence() do { } while(0)
#include "gen_interpolation.c"

/* interpolator: 0 bilinear, 1 bicubic, 2 mitchell (enum dt_interpolation_type) */
int ref_finalscale(const float *in, float *out, int in_w, int in_h, double in_scale, int out_w, int out_h, double out_scale, int interpolator)
{
  dt_iop_roi_t roi_in = { 0, 0, in_w, in_h, in_scale }, roi_out = { 0, 0, out_w, out_h, out_scale };
  _interpolation_resample_plain(&dt_interpolator[interpolator], out, &roi_out, in, &roi_in);
  return 0;
}
/* the resampling plan of one axis, flattened for comparison: returns the number of taps in total, -1 for scale == 1 */
int ref_resampling_plan(int interpolator, int in, int in_x0, int out, int out_x0, float scale, int *lengths, float *kernel, int *index, int max_taps)
{
  int *l = NULL, *i = NULL, *m = NULL;
  float *k = NULL;
  if(_prepare_resampling_plan(&dt_interpolator[interpolator], in, in_x0, out, out_x0, scale, &l, &k, &i, &m)) return -2;
  if(!l) return -1;
  int n = 0;
  for(int x = 0; x < out; x++) n += l[x];
  if(n > max_taps) n = -3;
  else
  {
    memcpy(lengths, l, sizeof(int) * out);
    memcpy(kernel, k, sizeof(float) * n);
    memcpy(index, i, sizeof(int) * n);
  }
  free(l);
  return n;
}

/* dt_iop_clip_and_zoom_roi with the ROIs as given (initialscale, iop/initialscale.c:122-129) */
int ref_clip_and_zoom(const float *in, float *out, int in_x, int in_y, int in_w, int in_h, double in_scale, int out_x, int out_y, int out_w, int out_h,
                      double out_scale, int interpolator)
{
  dt_iop_roi_t roi_in = { in_x, in_y, in_w, in_h, in_scale }, roi_out = { out_x, out_y, out_w, out_h, out_scale };
  _interpolation_resample_plain(&dt_interpolator[interpolator], out, &roi_out, in, &roi_in);
  return 0;
}
