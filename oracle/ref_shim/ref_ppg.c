/* oracle/_ref wrapper: the PPG demosaicer.  TEST INFRASTRUCTURE ONLY.
 *
 * iop/demosaic/ppg.c and basic.c are fragments that demosaic.c #includes; oracle/Makefile cuts verbatim into
 * oracle/_ref/gen_demosaic_ppg.c:  basic.c :129-186 (SWAP, pre_median_b, pre_median),  ppg.c :21-211 (demosaic_ppg).
 * iop/demosaic/passthrough.c :21-87 (passthrough_monochrome, passthrough_color) and iop/demosaic.c :480-532
 * (_downsample_bayer_half_size) and :543-666 (_downsample_xtrans_missing_colour, _downsample_xtrans_half_size) ride along.
 * demosaic.c:1218-1226 calls it with roi_out's origin zeroed and the ROI-shifted filters word.
 */
#include "ref_piece.h"
#include "gen_imageop_math.c" /* develop/imageop_math.h:175-219: dt_iop_alpha_copy, FC, FCxtrans */
static inline void dt_iop_image_copy_by_size(float *const out, const float *const in, const size_t width, const size_t height, const size_t ch)
{ /* common/imagebuf.h:91-95 */
  memcpy(out, in, sizeof(float) * width * height * ch);
}
static inline float *dt_pixelpipe_cache_alloc_align_float_cache(size_t n, int id) { return aligned_alloc(64, ((n * sizeof(float) + 63) / 64) * 64); }
static inline void dt_pixelpipe_cache_free_align(const void *p) { free((void *)p); }
#define RED 0
#define GREEN 1
#define BLUE 2
#include "gen_demosaic_ppg.c"

/* out: width * height * 4 floats, pre-filled by the caller (the alpha of the outer three pixels is left as found) */
int ref_demosaic_ppg(float *out, const float *in, int width, int height, uint32_t filters, float median_thrs)
{
  const dt_iop_roi_t roi = { 0, 0, width, height, 1.0 };
  return demosaic_ppg(out, in, &roi, &roi, filters, median_thrs);
}

/* demosaic.c:1111-1118: roi_out arrives with its origin zeroed, roi_in keeps the ROI origin; filters = the sensor's word */
int ref_demosaic_passthrough(float *out, const float *in, int width, int height, int x, int y, uint32_t filters, const uint8_t xtrans[36], int colour)
{
  dt_iop_roi_t roi_in = { x, y, width, height, 1.0 }, roi_out = { 0, 0, width, height, 1.0 };
  if(colour)
    passthrough_color(out, in, &roi_out, &roi_in, filters, (const uint8_t(*)[6])xtrans);
  else
    passthrough_monochrome(out, in, &roi_out, &roi_in);
  return 0;
}

/* demosaic.c:1101-1108 for a Bayer sensor with the post-filter off (data->color_smoothing == 0): out is (w+1)/2 x (h+1)/2 */
int ref_demosaic_downsample(float *out, const float *in, int width, int height, uint32_t filters)
{
  const dt_iop_roi_t roi_in = { 0, 0, width, height, 1.0 }, roi_out = { 0, 0, (width + 1) / 2, (height + 1) / 2, 1.0 };
  const double cam_to_rgb[3][4] = { { 0 } };
  _downsample_bayer_half_size(out, in, &roi_out, &roi_in, filters, 0, cam_to_rgb);
  return 0;
}

/* demosaic.c:1103-1104 for an X-Trans sensor, post-filter off: roi_in keeps the ROI origin (FCxtrans adds it), xtrans = dsc_in.xtrans */
int ref_demosaic_downsample_xtrans(float *out, const float *in, int width, int height, int x, int y, const uint8_t xtrans[36])
{
  const dt_iop_roi_t roi_in = { x, y, width, height, 1.0 }, roi_out = { 0, 0, (width + 1) / 2, (height + 1) / 2, 1.0 };
  _downsample_xtrans_half_size(out, in, &roi_out, &roi_in, (const uint8_t(*)[6])xtrans);
  return 0;
}

/* the same for a four-colour Bayer sensor (img->flags & DT_IMAGE_4BAYER): the camera primaries go through data->CAM_to_RGB */
int ref_demosaic_downsample4(float *out, const float *in, int width, int height, uint32_t filters, const double cam_to_rgb[12])
{
  const dt_iop_roi_t roi_in = { 0, 0, width, height, 1.0 }, roi_out = { 0, 0, (width + 1) / 2, (height + 1) / 2, 1.0 };
  _downsample_bayer_half_size(out, in, &roi_out, &roi_in, filters, 1, (const double(*)[4])cam_to_rgb);
  return 0;
}
