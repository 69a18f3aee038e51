/* oracle/_ref wrapper: colour calibration (channelmixerrgb), the pixel loop.  TEST INFRASTRUCTURE ONLY.
 *
 * iop/channelmixerrgb.c carries 5 k lines of GUI, colour-checker fitting and illuminant detection; oracle/Makefile cuts
 * its pixel path verbatim into oracle/_ref/gen_channelmixerrgb.c:
 *     :98-99    CHANNEL_SIZE, INVERSE_SQRT_3        :114-119  dt_iop_channelmixer_rgb_version_t
 *     :259-272  dt_iop_channelmixer_rbg_data_t      :641-706  gamut_mapping
 *     :707-763  luma_chroma                         :765-959  loop_switch
 * pixel/chromatic_adaptation.h (the CAT matrices and adaptations) is the reference's unmodified header.  process()
 * :1920-2078 fetches the work profile's two matrices, lets the GUI run its detections, re-derives the illuminant for
 * DT_ILLUMINANT_CAMERA from the image's metadata, then calls loop_switch() with the committed data: that call is what
 * ref_channelmixerrgb() makes.
 */
#include "ref_piece.h"
#include "math/matrices.h"
#include "pixel/chromatic_adaptation.h"
typedef int dt_illuminant_t; /* pixel/illuminants.h (pulls common/image.h); only stored in the data block */
#define dt_omploop_sfence() do { } while(0)
#include "gen_channelmixerrgb.c"

size_t ref_channelmixerrgb_sizeof_data(void) { return sizeof(dt_iop_channelmixer_rbg_data_t); }
size_t ref_channelmixerrgb_offsetof(int which)
{
  switch(which)
  {
    case 0: return offsetof(dt_iop_channelmixer_rbg_data_t, saturation);
    case 1: return offsetof(dt_iop_channelmixer_rbg_data_t, illuminant);
    case 2: return offsetof(dt_iop_channelmixer_rbg_data_t, p);
    case 3: return offsetof(dt_iop_channelmixer_rbg_data_t, adaptation);
    default: return offsetof(dt_iop_channelmixer_rbg_data_t, version);
  }
}
/* data: a dt_iop_channelmixer_rbg_data_t as commit_params() left it; the two matrices: rows of a dt_colormatrix_t (3x4) */
int ref_channelmixerrgb(const float *in, float *out, int width, int height, const void *data_blob, const float rgb_to_xyz[12], const float xyz_to_rgb[12])
{
  dt_iop_channelmixer_rbg_data_t *d = aligned_alloc(64, ((sizeof(*d) + 63) / 64) * 64);
  memcpy(d, data_blob, sizeof(*d));
  dt_colormatrix_t RGB_to_XYZ = { { 0 } }, XYZ_to_RGB = { { 0 } };
  memcpy(RGB_to_XYZ, rgb_to_xyz, sizeof(float) * 12);
  memcpy(XYZ_to_RGB, xyz_to_rgb, sizeof(float) * 12);
  if(d->adaptation >= DT_ADAPTATION_LINEAR_BRADFORD && d->adaptation <= DT_ADAPTATION_RGB)
    loop_switch(in, out, width, height, 4, XYZ_to_RGB, RGB_to_XYZ, d->MIX, d->illuminant, d->saturation, d->lightness, d->grey, d->p, d->gamut, d->clip,
                d->apply_grey, d->adaptation, d->version);
  free(d);
  return 0;
}
