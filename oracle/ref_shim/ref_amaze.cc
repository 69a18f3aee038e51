/* oracle/_ref wrapper: AMaZE demosaic.  TEST INFRASTRUCTURE ONLY.
 * Textually includes the unmodified /root/reference/src/iop/demosaic/amaze.cc (amaze_demosaic_RT :180-1419); its two
 * develop/ headers are shadowed by oracle/ref_shim/shadow/develop/ (they drag in the GUI and OpenCL). */
#ifdef REF_STRICT
#define __DT_CLONE_TARGETS__
#endif
#include "iop/demosaic/amaze.cc"
#include <string.h>

extern "C" int ref_amaze_demosaic(float *out, const float *in, int width, int height, uint32_t filters, const float processed_maximum[3])
{
  dt_dev_pixelpipe_iop_t piece;
  memset(&piece, 0, sizeof(piece));
  for(int k = 0; k < 3; k++) piece.dsc_in.processed_maximum[k] = processed_maximum[k];
  const dt_iop_roi_t roi = { 0, 0, width, height, 1.0 };
  amaze_demosaic_RT(&piece, in, out, &roi, &roi, (int)filters);
  return 0;
}
