/* oracle/_ref: the operator-surface types the sliced process() bodies of the pointwise pipe modules dereference
 * (rawprepare, temperature, highlights, exposure, gamma).  Members are the reference's own names
 * (develop/pixelpipe_hb.h:101-166, develop/develop.h:123, pixel/format.h for roi and buffer descriptor, which is the
 * reference's unmodified header).  TEST INFRASTRUCTURE ONLY. */
#pragma once
#include <glib.h>
#include <math.h>
#include <float.h>
#include <string.h>
#include <stdlib.h>
#include <stdint.h>
#include <assert.h>
#include "system/macros.h"
#include "system/openmp.h"
#ifdef REF_STRICT
#define __DT_CLONE_TARGETS__
#else
#include "system/target_clones.h"
#endif
#include "system/simd.h"
#include "math/math.h"
#include "pixel/format.h"

#define DT_DEV_PIXELPIPE_DISPLAY_MASK 1 /* develop/develop.h:123 */
typedef struct dt_develop_t { int gui_attached; struct dt_dev_pixelpipe_t *pipe; } dt_develop_t;
typedef struct dt_iop_module_t { dt_develop_t *dev; void *gui_data; } dt_iop_module_t;
typedef struct dt_dev_pixelpipe_t { int type; int mask_display; float iscale; int bypass_blendif; } dt_dev_pixelpipe_t;
typedef struct dt_dev_pixelpipe_iop_t
{
  void *data;
  dt_iop_roi_t buf_in, roi_in, roi_out;
  dt_iop_buffer_dsc_t dsc_in;
} dt_dev_pixelpipe_iop_t;
#undef HAVE_OPENCL
