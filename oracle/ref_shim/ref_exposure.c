/* oracle/_ref wrapper: exposure.  TEST INFRASTRUCTURE ONLY.
 * oracle/Makefile cuts verbatim from iop/exposure.c: :97-103 dt_iop_exposure_mode_t, :116-124 params, :151-157 data,
 * :501-544 process(); dt_iop_alpha_copy comes from the develop/imageop_math.h cut (gen_imageop_math.c). */
#include "ref_piece.h"
#include "gen_imageop_math.c"
#define process exposure_process
#include "gen_exposure.c"
#undef process

int ref_exposure(const float *in, float *out, int width, int height, int channels, float black, float scale, int mask_display)
{
  dt_iop_exposure_data_t d;
  memset(&d, 0, sizeof(d));
  d.black = black;
  d.scale = scale;
  dt_dev_pixelpipe_t pipe = { 1, mask_display, 1.0f, 0 };
  dt_dev_pixelpipe_iop_t piece;
  memset(&piece, 0, sizeof(piece));
  piece.data = &d;
  piece.roi_in = piece.roi_out = (dt_iop_roi_t){ 0, 0, width, height, 1.0 };
  piece.dsc_in.channels = channels;
  return exposure_process(NULL, &pipe, &piece, in, out);
}
size_t ref_exposure_sizeof_data(void) { return sizeof(dt_iop_exposure_data_t); }
size_t ref_exposure_offsetof_black(void) { return offsetof(dt_iop_exposure_data_t, black); }
