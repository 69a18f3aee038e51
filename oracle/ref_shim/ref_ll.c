/* oracle/_ref wrapper: the reference's local Laplacian filter.  TEST INFRASTRUCTURE ONLY.
 * Textually includes the unmodified /root/reference/src/pixel/locallaplacian.c
 * (local_laplacian_internal() :354-563 and everything under it). */
#include <glib.h>
#ifdef REF_STRICT
#define __DT_CLONE_TARGETS__
#endif
#include "pixel/locallaplacian.c"

/* local_laplacian(), pixel/locallaplacian.h:71-83, as iop/bilat.c:354 calls it (no preview boundary) */
int ref_local_laplacian(const float *in, float *out, int wd, int ht, float sigma, float shadows, float highlights, float clarity)
{
  return local_laplacian_internal(in, out, wd, ht, sigma, shadows, highlights, clarity, 0, NULL);
}
