/* shadow of src/develop/pixelpipe_hb.h for the oracle/_ref build of iop/demosaic/amaze.cc: only what that file
 * dereferences (develop/pixelpipe_hb.h:101-166, pixel/format.h).  TEST INFRASTRUCTURE ONLY. */
#pragma once
#include <glib.h>
#include "pixel/format.h"
typedef struct dt_dev_pixelpipe_iop_t
{
  dt_iop_buffer_dsc_t dsc_in;
} dt_dev_pixelpipe_iop_t;
