/* oracle/_ref wrapper: the guided-Laplacian post-filter of the half-size demosaic.  TEST INFRASTRUCTURE ONLY.
 *
 * oracle/Makefile cuts iop/demosaic.c :117 (DOWNSAMPLE_GUIDED_SCALES) and :681-926 (_downsample_guided_laplacian_fit, _apply,
 * _postfilter) verbatim into oracle/_ref/gen_demosaic_postfilter.c; pixel/bspline.h (decompose_2D_Bspline, blur_2D_Bspline)
 * is included unmodified.  demosaic.c:1108 calls it on the half-size RGBA frame with data->color_smoothing iterations.
 */
#include <glib.h>
#include <math.h>
#include <float.h>
#include <string.h>
#include <stdlib.h>
#include <stdint.h>
#include "system/macros.h"
#include "system/mem_alloc.h"
#include "system/openmp.h"
#ifdef REF_STRICT
#define __DT_CLONE_TARGETS__
#else
#include "system/target_clones.h"
#endif
#include "system/simd.h"
#include "math/math.h"
#include "math/openmp_maths.h"
#include "pixel/format.h"
#include "pixel/dwt.h"
#include "caches/pixelpipe_cache_alloc.h"
#include "pixel/bspline.h"
#ifndef RED
#define RED 0
#define GREEN 1
#define BLUE 2
#define ALPHA 3
#endif
static inline void postfilter_image_copy(float *const out, const float *const in, const size_t width, const size_t height, const size_t ch)
{ /* common/imagebuf.h:91-95 */
  memcpy(out, in, sizeof(float) * width * height * ch);
}
#define dt_iop_image_copy_by_size postfilter_image_copy
#include "gen_demosaic_postfilter.c"

/* rgba: width * height * 4 floats, filtered in place */
int ref_demosaic_downsample_postfilter(float *rgba, int width, int height, int iterations)
{
  return _downsample_guided_laplacian_postfilter(rgba, (size_t)width, (size_t)height, iterations);
}
