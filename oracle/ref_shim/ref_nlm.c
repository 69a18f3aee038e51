/* oracle/_ref wrapper: the reference's non-local-means core.  TEST INFRASTRUCTURE ONLY.
 *
 * Textually includes the unmodified /root/reference/src/pixel/nlmeans_core.c (nlmeans_denoise() :315-532,
 * define_patches :107-145, init_column_sums :214-264, ...).  That file includes four headers of
 * develop/, common/ and iop/ which drag in GTK; it uses nothing of them beyond dt_iop_roi_t and
 * dt_dev_pixelpipe_type_t, so their include guards are pre-defined here and those two names supplied.
 */
#include <glib.h>
#include <stdlib.h>
#include <string.h>
#ifdef REF_STRICT
#define __DT_CLONE_TARGETS__
#endif
#include "pixel/format.h" /* dt_iop_roi_t */
typedef int dt_dev_pixelpipe_type_t;
#define DT_COMMON_OPENCL_H   /* common/opencl.h   */
#define DT_DEVELOP_IMAGEOP_H /* develop/imageop.h */
#define DT_IOP_PARAMS_T      /* iop/iop_api.h     */
#define DT_DEVELOP_DEVELOP_H /* develop/develop.h */
#define HAVE_CONFIG_H         /* nlmeans_core.c:25-28 only includes the allocator header under it */

#include "pixel/nlmeans_core.c"

/* caches/pixelpipe_cache_alloc.h and system/openmp.h leave these to lib_ansel */
void *dt_pixelpipe_cache_alloc_align_cache_impl(size_t size, int id, const char *name)
{
  (void)id;
  (void)name;
  return aligned_alloc(64, ((size + 63) / 64) * 64);
}
void dt_pixelpipe_cache_free_align_cache(void **mem, const char *message)
{
  (void)message;
  if(mem && *mem)
  {
    free(*mem);
    *mem = NULL;
  }
}
int dt_get_num_openmp_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* plain-C entry point: nlmeans_denoise() with the parameter block spelled out */
void ref_nlmeans_denoise(const float *in, float *out, int width, int height, float scattering, float scale, float luma,
                         float chroma, float center_weight, float sharpness, int patch_radius, int search_radius,
                         int decimate, const float norm[4])
{
  const dt_iop_roi_t roi = { 0, 0, width, height, 1.0 };
  const dt_nlmeans_param_t params = { .scattering = scattering, .scale = scale, .luma = luma, .chroma = chroma,
                                      .center_weight = center_weight, .sharpness = sharpness, .patch_radius = patch_radius,
                                      .search_radius = search_radius, .decimate = decimate, .norm = norm };
  nlmeans_denoise(in, out, &roi, &roi, &params);
}

/* ---- the denoise (non-local means) iop's parameter derivation: iop/nlmeans.c process_cpu() :416-456, cut
 * verbatim into oracle/_ref/gen_nlmeans_iop.c together with the params struct :81-88 -------------------- */
typedef enum { DT_DEV_PIXELPIPE_EXPORT = 1, DT_DEV_PIXELPIPE_THUMBNAIL = 4 } ref_pipe_type_names_t;
#define DT_DEV_PIXELPIPE_DISPLAY_MASK 1
typedef struct dt_develop_t { int dummy; } dt_develop_t;
typedef struct dt_iop_module_t { dt_develop_t *dev; } dt_iop_module_t;
typedef struct dt_dev_pixelpipe_t { int type; int mask_display; int has_preview_output; } dt_dev_pixelpipe_t;
typedef struct dt_dev_pixelpipe_iop_t { void *data; dt_iop_module_t *module; dt_iop_roi_t roi_in, roi_out; } dt_dev_pixelpipe_iop_t;
static int dt_dev_pixelpipe_has_preview_output(const dt_develop_t *dev, const dt_dev_pixelpipe_t *pipe, const dt_iop_roi_t *roi)
{
  (void)dev;
  (void)roi;
  return pipe->has_preview_output;
}
static void dt_iop_alpha_copy(const void *ivoid, void *ovoid, const size_t width, const size_t height)
{ /* develop/imageop_math.c: copies channel 3 */
  const float *in = ivoid;
  float *out = ovoid;
  for(size_t k = 3; k < 4 * width * height; k += 4) out[k] = in[k];
}
#include "gen_nlmeans_iop.c"

int ref_nlmeans_iop(const float *in, float *out, int width, int height, const float params[4], double roi_scale, int pipe_type,
                    int has_preview_output, int mask_display)
{
  dt_iop_nlmeans_params_t d = { params[0], params[1], params[2], params[3] };
  dt_develop_t dev = { 0 };
  dt_iop_module_t mod = { &dev };
  dt_dev_pixelpipe_t pipe = { pipe_type, mask_display, has_preview_output };
  dt_dev_pixelpipe_iop_t piece = { &d, &mod, { 0, 0, width, height, roi_scale }, { 0, 0, width, height, roi_scale } };
  process_cpu(&pipe, &piece, in, out, &piece.roi_in, &piece.roi_out, nlmeans_denoise);
  return 0;
}
