/* oracle/_ref wrapper: the bilateral grid behind local contrast's "bilateral grid" mode.  TEST INFRASTRUCTURE ONLY.
 * Textually includes the unmodified /root/reference/src/pixel/bilateral.c (dt_bilateral_init :157-180, _splat :182-265,
 * _blur :341-353, _slice :355-393) and runs it the way iop/bilat.c process() :346-353 does.
 *
 * The splat accumulates in horizontal slices, one per OpenMP thread, and adds the slices' partial grids afterwards: the
 * rounding of a grid cell fed by two slices depends on the thread count.  `threads` selects it; 1 is plain raster order. */
#include <glib.h>
#ifdef REF_STRICT
#define __DT_CLONE_TARGETS__
#endif
static int ref_bilateral_threads = 1;
#define dt_get_num_openmp_threads ref_bilateral_num_threads
static int ref_bilateral_num_threads(void) { return ref_bilateral_threads; }
#include "pixel/bilateral.c"
#undef dt_get_num_openmp_threads

int ref_bilateral(const float *in, float *out, int width, int height, float sigma_s, float sigma_r, float detail, int threads)
{
  ref_bilateral_threads = threads < 1 ? 1 : threads;
  dt_bilateral_t *b = dt_bilateral_init(width, height, sigma_s, sigma_r);
  if(!b) return 1;
  dt_bilateral_splat(b, in);
  dt_bilateral_blur(b);
  dt_bilateral_slice(b, in, out, detail);
  dt_bilateral_free(b);
  return 0;
}
/* the grid after the splat (and optionally the blur), for a finer comparison: size_x * size_y * size_z floats */
int ref_bilateral_grid(const float *in, float *grid, int max_floats, int dims[3], int width, int height, float sigma_s, float sigma_r, int blur, int threads)
{
  ref_bilateral_threads = threads < 1 ? 1 : threads;
  dt_bilateral_t *b = dt_bilateral_init(width, height, sigma_s, sigma_r);
  if(!b) return 1;
  dt_bilateral_splat(b, in);
  if(blur) dt_bilateral_blur(b);
  dims[0] = (int)b->size_x, dims[1] = (int)b->size_y, dims[2] = (int)b->size_z;
  const size_t n = b->size_x * b->size_y * b->size_z;
  const int rc = n > (size_t)max_floats ? 2 : 0;
  if(!rc) memcpy(grid, b->buf, n * sizeof(float));
  dt_bilateral_free(b);
  return rc;
}
