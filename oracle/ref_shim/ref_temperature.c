/* oracle/_ref wrapper: temperature (white balance multipliers on the mosaic).  TEST INFRASTRUCTURE ONLY.
 *
 * oracle/Makefile cuts verbatim: iop/temperature.c :150-153 (dt_iop_temperature_data_t), :486-608 (process());
 * develop/imageop_math.h :175-219 (dt_iop_alpha_copy, FC, FCxtrans) into gen_imageop_math.c -- the header itself
 * is shadowed for the amaze build.
 */
#include "ref_piece.h"
#include "gen_imageop_math.c"
#define process temperature_process
#include "gen_temperature.c"
#undef process

int ref_temperature(const float *in, float *out, int x, int y, int width, int height, uint32_t filters, const uint8_t xtrans[36],
                    int channels, const float coeffs[4], int mask_display)
{
  dt_iop_temperature_data_t d;
  for(int k = 0; k < 4; k++) d.coeffs[k] = coeffs[k];
  dt_dev_pixelpipe_t pipe = { 1, mask_display, 1.0f, 0 };
  dt_dev_pixelpipe_iop_t piece;
  memset(&piece, 0, sizeof(piece));
  piece.data = &d;
  piece.roi_in = piece.roi_out = (dt_iop_roi_t){ x, y, width, height, 1.0 };
  piece.dsc_in.filters = filters;
  piece.dsc_in.channels = channels;
  if(xtrans) memcpy(piece.dsc_in.xtrans, xtrans, 36);
  return temperature_process(NULL, &pipe, &piece, in, out);
}
