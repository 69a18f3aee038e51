/* oracle/_ref wrapper: diffuse or sharpen (anisotropic multi-scale heat PDE).  TEST INFRASTRUCTURE ONLY.
 *
 * iop/diffuse.c carries GUI, presets and OpenCL host code in the same translation unit.  oracle/Makefile cuts
 * its CPU pixel path out verbatim (oracle/ref_shim/slice.py) into oracle/_ref/gen_diffuse.c:
 *     :75-109     dt_iop_diffuse_params_t          :132-162   data typedef, isotropy enum + check
 *     :612-962    init_reconstruct .. heat_PDE_diffusion, compute_anisotropy_factor
 *     :977-1259   wavelets_process, build_mask, inpaint_mask, process()
 * pixel/bspline.h (decompose_2D_Bspline) is included unmodified.
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
#include "iop/noise_generator.h"
#include "caches/pixelpipe_cache_alloc.h"
#include "pixel/bspline.h"

typedef void dt_iop_params_t;
typedef struct dt_iop_module_t { size_t params_size; } dt_iop_module_t;
typedef struct dt_dev_pixelpipe_t { int type; float iscale; } dt_dev_pixelpipe_t;
typedef struct dt_dev_pixelpipe_iop_t { void *data; dt_iop_roi_t roi_in, roi_out; int cache_output_on_ram; } dt_dev_pixelpipe_iop_t;
static inline float dt_dev_get_module_scale(const dt_dev_pixelpipe_t *pipe, const dt_iop_roi_t *roi) { return pipe->iscale / roi->scale; }
#define DIFFUSE_V3 0
#define DEBUG_DUMP_PFM 0
#undef HAVE_OPENCL

/* module entry points have plain names in every iop; keep this TU's private */
#define commit_params diffuse_commit_params
#define process diffuse_process
#define tiling_callback diffuse_tiling_callback
#include "gen_diffuse.c"
#undef process

size_t ref_diffuse_sizeof_params(void) { return sizeof(dt_iop_diffuse_params_t); }

int ref_diffuse_process(const float *in, float *out, int width, int height, const void *params, float iscale, float roi_scale)
{
  dt_dev_pixelpipe_t pipe = { 1, iscale };
  dt_dev_pixelpipe_iop_t piece;
  memset(&piece, 0, sizeof(piece));
  piece.data = (void *)params;
  piece.roi_in = (dt_iop_roi_t){ 0, 0, width, height, roi_scale };
  piece.roi_out = piece.roi_in;
  return diffuse_process(NULL, &pipe, &piece, in, out);
}
