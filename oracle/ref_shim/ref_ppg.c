/* oracle/_ref wrapper: the PPG demosaicer.  TEST INFRASTRUCTURE ONLY.
 *
 * iop/demosaic/ppg.c and basic.c are fragments that demosaic.c #includes; oracle/Makefile cuts verbatim into
 * oracle/_ref/gen_demosaic_ppg.c:  basic.c :129-186 (SWAP, pre_median_b, pre_median),  ppg.c :21-211 (demosaic_ppg).
 * demosaic.c:1218-1226 calls it with roi_out's origin zeroed and the ROI-shifted filters word.
 */
#include "ref_piece.h"
static inline int FC(const size_t row, const size_t col, const uint32_t filters)
{ /* develop/imageop_math.h:190-193 */
  return filters >> (((row << 1 & 14) + (col & 1)) << 1) & 3;
}
static inline void dt_iop_image_copy_by_size(float *const out, const float *const in, const size_t width, const size_t height, const size_t ch)
{ /* common/imagebuf.h:91-95 */
  memcpy(out, in, sizeof(float) * width * height * ch);
}
static inline float *dt_pixelpipe_cache_alloc_align_float_cache(size_t n, int id) { return aligned_alloc(64, ((n * sizeof(float) + 63) / 64) * 64); }
static inline void dt_pixelpipe_cache_free_align(const void *p) { free((void *)p); }
#include "gen_demosaic_ppg.c"

/* out: width * height * 4 floats, pre-filled by the caller (the alpha of the outer three pixels is left as found) */
int ref_demosaic_ppg(float *out, const float *in, int width, int height, uint32_t filters, float median_thrs)
{
  const dt_iop_roi_t roi = { 0, 0, width, height, 1.0 };
  return demosaic_ppg(out, in, &roi, &roi, filters, median_thrs);
}
