/* oracle/_ref wrapper: the demosaic module's optional pre/post passes.  TEST INFRASTRUCTURE ONLY.
 *
 * iop/demosaic/basic.c is a fragment that demosaic.c #includes; oracle/Makefile cuts  :129-134 (SWAP),
 * :188-246 (color_smoothing), :247-329 (green_equilibration_lavg / _favg) verbatim into oracle/_ref/gen_demosaic_basic.c.
 */
#include <glib.h>
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <stdint.h>
#include "system/macros.h"
#include "system/openmp.h"
#ifdef REF_STRICT
#define __DT_CLONE_TARGETS__
#else
#include "system/target_clones.h"
#endif
#include "pixel/format.h"

static inline int FC(const size_t row, const size_t col, const uint32_t filters)
{ /* develop/imageop_math.h:190-193 */
  return filters >> (((row << 1 & 14) + (col & 1)) << 1) & 3;
}
static inline void dt_iop_image_copy_by_size(float *const out, const float *const in, const size_t width, const size_t height, const size_t ch)
{ /* develop/imageop_math.c */
  memcpy(out, in, sizeof(float) * width * height * ch);
}

#include "gen_demosaic_basic.c"

void ref_color_smoothing(float *out, int width, int height, int passes)
{
  const dt_iop_roi_t roi = { 0, 0, width, height, 1.0 };
  color_smoothing(out, &roi, passes);
}
void ref_green_eq_lavg(float *out, const float *in, int width, int height, uint32_t filters, int x, int y, float thr)
{
  green_equilibration_lavg(out, in, width, height, filters, x, y, thr);
}
void ref_green_eq_favg(float *out, const float *in, int width, int height, uint32_t filters, int x, int y)
{
  green_equilibration_favg(out, in, width, height, filters, x, y);
}
