/* oracle/_ref wrapper for the reference RCD demosaic.  TEST INFRASTRUCTURE ONLY.
 *
 * This translation unit textually includes the reference's own, unmodified
 *   /root/reference/src/iop/demosaic/rcd.c          (rcd_demosaic :274-564, rcd_ppg_border :91-272)
 * (found through -I$(REF)/src) after declaring the handful of names that file expects from the
 * rest of lib_ansel.  Wherever a reference header is self-contained it is included as is
 * (pixel/format.h, system/{simd,openmp,target_clones,fp_mode}.h); the remaining names are
 * one-line stand-ins, each citing what it replaces.  No reference source is copied into this
 * repository: the build reads it where it lies, and the output goes to oracle/_ref/ only.
 *
 * Two builds are made from this file (oracle/Makefile):
 *   ref-fast   : the reference's release flags (CMakeLists.txt:261-272) + target_clones
 *   ref-strict : -O2 -fno-fast-math -ffp-contract=off (C-standard semantics of the same source)
 */
#include <glib.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdint.h>

#include "system/mem_alloc.h"
#include "system/openmp.h"
#ifdef REF_STRICT
#define __DT_CLONE_TARGETS__
#else
#include "system/target_clones.h"
#endif
#include "system/fp_mode.h"
#include "system/simd.h"
#include "pixel/format.h" /* dt_iop_roi_t, dt_iop_buffer_dsc_t */

#define INLINE inline

/* develop/pixelpipe_hb.h:101-166 -- only the member rcd.c dereferences */
typedef struct dt_dev_pixelpipe_iop_t
{
  dt_iop_buffer_dsc_t dsc_in;
} dt_dev_pixelpipe_iop_t;

/* develop/imageop_math.h:190-193 (that header drags in OpenCL and image.h) */
static inline int FC(const size_t row, const size_t col, const uint32_t filters)
{
  return filters >> (((row << 1 & 14) + (col & 1)) << 1) & 3;
}
/* math/math.h:199-202 */
static inline float sqf(const float x)
{
  return x * x;
}
/* iop/demosaic.c:250-257 */
static inline __attribute__((always_inline)) float intp(float a, float b, float c)
{
  return a * (b - c) + c;
}

#define _(s) (s)
#define dt_control_log(...) fprintf(stderr, __VA_ARGS__)

/* caches/pixelpipe_cache_alloc.h:59-156.  The reference hands out uninitialised memory here;
 * `ref_poison` lets the parity harness find the output pixels that depend on it
 * (rcd.c:339 "TODO: figure out what part of rgb is being accessed
 * without initialization"). */
static float ref_poison = 0.0f;
static float *ref_scratch_alloc(size_t n)
{
  float *p = aligned_alloc(64, ((n * sizeof(float) + 63) / 64) * 64);
  if(p)
    for(size_t k = 0; k < n; k++) p[k] = ref_poison;
  return p;
}
#define dt_pixelpipe_cache_alloc_align_float_cache(n, id) ref_scratch_alloc(n)
#define dt_pixelpipe_cache_free_align(p) free(p)

#include "iop/demosaic/rcd.c"

/* Plain-C entry point for the harness.  filters is the already ROI-shifted dcraw word
 * (iop/demosaic.c:1071).  Returns 0 on success, 1 when the reference would have logged
 * "too small area" and left the output untouched (rcd.c:280-284). */
int ref_rcd_demosaic(float *out, const float *in, int width, int height, uint32_t filters,
                     const float processed_maximum[3], float poison)
{
  dt_dev_pixelpipe_iop_t piece;
  memset(&piece, 0, sizeof(piece));
  for(int k = 0; k < 3; k++) piece.dsc_in.processed_maximum[k] = processed_maximum[k];
  dt_iop_roi_t roi_in = { 0, 0, width, height, 1.0 };
  dt_iop_roi_t roi_out = roi_in;
  ref_poison = poison;
  if(width < 16 || height < 16) return 1;
  rcd_demosaic(&piece, out, in, &roi_out, &roi_in, filters);
  return 0;
}

/* layout facts the product's include/b200iop.h mirrors */
size_t ref_sizeof_roi(void) { return sizeof(dt_iop_roi_t); }
size_t ref_sizeof_dsc(void) { return sizeof(dt_iop_buffer_dsc_t); }
size_t ref_offsetof_dsc_filters(void) { return offsetof(dt_iop_buffer_dsc_t, filters); }
size_t ref_offsetof_dsc_processed_maximum(void) { return offsetof(dt_iop_buffer_dsc_t, processed_maximum); }
size_t ref_offsetof_dsc_temperature_coeffs(void) { return offsetof(dt_iop_buffer_dsc_t, temperature.coeffs); }
