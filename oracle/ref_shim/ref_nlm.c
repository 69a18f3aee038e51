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
