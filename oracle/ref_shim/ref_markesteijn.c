/* oracle/_ref wrapper: Frank Markesteijn's demosaicer for X-Trans sensors, 1 and 3 passes.  TEST INFRASTRUCTURE ONLY.
 *
 * oracle/Makefile cuts verbatim into oracle/_ref/gen_markesteijn.c:
 *     iop/demosaic/markesteijn.c :25-523   SQR, TS, hexmap, xtrans_markesteijn_interpolate
 * develop/imageop_math.h :175-219 (FC, FCxtrans) comes through gen_imageop_math.c; the per-thread scratch allocator is the
 * reference's own header (caches/pixelpipe_cache_alloc.h), its two library functions are in ref_nlm.c.
 */
#include "ref_piece.h"
#include <limits.h>
#include <stdio.h>
#include "system/mem_alloc.h"
#include "caches/pixelpipe_cache_alloc.h"
#include "gen_imageop_math.c"
#include "gen_markesteijn.c"

/* in: the X-Trans mosaic of the region (width x height floats), xtrans: the sensor's 6x6 pattern, (x, y): the region's
 * origin on the sensor (roi_in.x/.y, which FCxtrans adds); out: 4 floats per pixel, lane 3 untouched */
void ref_markesteijn(float *out, const float *in, int width, int height, int x, int y, const uint8_t xtrans[36], int passes)
{
  const dt_iop_roi_t roi_in = { x, y, width, height, 1.0 }, roi_out = { 0, 0, width, height, 1.0 };
  xtrans_markesteijn_interpolate(out, in, &roi_out, &roi_in, (const uint8_t(*)[6])xtrans, passes);
}
