/* oracle/_ref wrapper: the VNG4 demosaicer and the dual (RCD/AMaZE + VNG4) blend.  TEST INFRASTRUCTURE ONLY.
 *
 * oracle/Makefile cuts verbatim into oracle/_ref/gen_demosaic_vng.c:
 *     imageio/imageio_core.h :48-55   FILTERS_ARE_CYGM / _RGBE / _4BAYER
 *     develop/masks/detail.c :91-120  dt_masks_extend_border      :159-234 the 9x9 blur (coefficients, FAST_BLUR_9, loop)
 *                            :282-346 dt_masks_calc_rawdetail_mask, calcBlendFactor, dt_masks_calc_detail_mask
 *     iop/demosaic/basic.c   :20-125  lin_interpolate             :129-134 SWAP   :188-246 color_smoothing
 *     iop/demosaic/vng.c     :33-202  vng_interpolate
 *     iop/demosaic.c         :250-257 intp
 *     iop/demosaic/dual.c    :34-112  slider2contrast, dual_demosaic
 * develop/imageop_math.h :175-219 (FC, FCxtrans) comes through gen_imageop_math.c; fcol :207-214 is restated below.
 */
/* This translation unit is compiled serially.  lin_interpolate's border loop (basic.c:28-55) jumps its inner loop
 * variable (`col = roi_out->width - 1`) inside an `omp parallel for collapse(2)`: with OpenMP the collapsed iteration
 * count no longer matches the iterations executed, threads run past their chunk and the last one past the end of both
 * buffers (AddressSanitizer: heap-buffer-overflow at basic.c:54; glibc aborts with "free(): invalid next size").  The
 * pixels inside the frame are the same either way; pinned is the serial (C) meaning of the loop. */
#undef _OPENMP
#include "ref_piece.h"
#include <limits.h>
#include <stdio.h>
#include "gen_imageop_math.c"
static inline int fcol(const int row, const int col, const uint32_t filters, const uint8_t (*const xtrans)[6])
{ /* develop/imageop_math.h:207-214 */
  if(filters == 9) return FCxtrans(row, col, NULL, xtrans);
  return FC(row, col, filters);
}
static inline void *dt_pixelpipe_cache_alloc_align_cache(size_t size, int id) { return aligned_alloc(64, ((size + 63) / 64) * 64); }
static inline float *dt_pixelpipe_cache_alloc_align_float(size_t n, const void *pipe) { return aligned_alloc(64, ((n * sizeof(float) + 63) / 64) * 64); }
static inline void dt_pixelpipe_cache_free_align(const void *p) { free((void *)p); }
#define dt_control_log(...) do { } while(0)
#define _(s) s
static inline unsigned dt_get_debug_flags(void) { return 0; }
#define DT_DEBUG_DEMOSAIC 1
#define DT_DEBUG_PERF 2
#define DT_DEV_PIXELPIPE_FULL 2
#define DT_DEV_PIXELPIPE_DISPLAY_PASSTHRU (1 << 12)
typedef struct dt_times_t { double clock, user; } dt_times_t;
#define dt_get_times(t) do { } while(0)
#include "gen_demosaic_vng.c"

int ref_vng_interpolate(float *out, const float *in, int width, int height, int x, int y, uint32_t filters, int only_linear)
{
  const dt_iop_roi_t roi_in = { x, y, width, height, 1.0 }, roi_out = { 0, 0, width, height, 1.0 };
  return vng_interpolate(out, in, &roi_out, &roi_in, filters, NULL, only_linear);
}
/* rgb: the high-frequency demosaicer's result, blended in place; mask != 0 -> the blend mask is written instead */
int ref_dual_demosaic(float *rgb, const float *raw, int width, int height, int x, int y, uint32_t filters, const float wb[4], float dual_threshold, int mask)
{
  dt_iop_roi_t roi_in = { x, y, width, height, 1.0 }, roi_out = { 0, 0, width, height, 1.0 };
  dt_dev_pixelpipe_t pipe = { 1, 0, 1.0f, 0 };
  dt_dev_pixelpipe_iop_t piece;
  memset(&piece, 0, sizeof(piece));
  for(int k = 0; k < 4; k++) piece.dsc_in.temperature.coeffs[k] = wb[k];
  return dual_demosaic(&pipe, &piece, rgb, raw, &roi_out, &roi_in, filters, NULL, mask, dual_threshold);
}
void ref_blur_9x9_coeff(float *c, float sigma) { dt_masks_blur_9x9_coeff(c, sigma); }
int ref_vng_interpolate_xtrans(float *out, const float *in, int width, int height, int x, int y, const uint8_t xtrans[36], int only_linear)
{
  const dt_iop_roi_t roi_in = { x, y, width, height, 1.0 }, roi_out = { 0, 0, width, height, 1.0 };
  return vng_interpolate(out, in, &roi_out, &roi_in, 9u, (const uint8_t(*)[6])xtrans, only_linear);
}
