/* oracle/_ref wrapper: the two geometry modules of every pipe that only move or resample pixels: flip (orientation) and
 * initialscale.  TEST INFRASTRUCTURE ONLY.
 * oracle/Makefile cuts verbatim: common/image.h :213-231 (dt_image_orientation_t), imageio/imageio_core.c :258-297
 * (dt_imageio_flip_buffers, what iop/flip.c process() :388-400 calls) into gen_geometry.c.  initialscale's process()
 * (iop/initialscale.c:122-129) is dt_iop_clip_and_zoom_roi with both ROIs as they are: the resampler of ref_finalscale.c,
 * reached through ref_clip_and_zoom() there. */
#include "ref_piece.h"
#include "gen_geometry.c"

int ref_flip(const void *in, void *out, int bpp, int width, int height, int orientation)
{
  dt_imageio_flip_buffers((char *)out, (const char *)in, bpp, width, height, width, height, bpp * width, (dt_image_orientation_t)orientation);
  return 0;
}
