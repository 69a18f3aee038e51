/* CPU restatement of the local Laplacian filter (local contrast module).  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src/pixel/locallaplacian.c: dl :53-58, ll_expand_gaussian :80-118,
 * ll_fill_boundary1/2 :120-145, pad_by_replication :147-159, gauss_expand :160-171, gauss_reduce :173-200,
 * ll_pad_input :204-280 (replication branch), ll_laplacian :283-293, curve_scalar :295-327, apply_curve
 * :329-352, local_laplacian_internal :354-563; caller iop/bilat.c process :336-360.
 *
 * Every boundary-fill pass of the reference copies already-computed neighbours, so each buffer is a pure
 * function "value at clamped coordinates"; that is how it is written here (and in the CUDA kernels).
 * Mixed precision is kept: the 4./256., 24.0, 4.0 and 2.0 literals make those expressions double.
 * Pinned bit-for-bit against the reference file compiled in place (oracle/_ref, ref_ll.c).
 */
#include "oracle_common.h"
#include "flt32_math.h"
#include <stdlib.h>
#include <string.h>

#define NUM_GAMMA 6
#define MAX_LEVELS 30
#define CLAMPS(A, L, H) ((A) > (L) ? ((A) < (H) ? (A) : (H)) : (L)) /* math/math.h:78 */

static inline int dl(int size, int level)
{
  for(int l = 0; l < level; l++) size = (size - 1) / 2 + 1;
  return size;
}
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* :80-118 at interior coordinates */
static inline float expand_at(const float *coarse, int i, int j, int wd)
{
  const int cw = (wd - 1) / 2 + 1;
  const int ind = (j / 2) * cw + i / 2;
  switch((i & 1) + 2 * (j & 1))
  {
    case 0:
      return (float)(4. / 256.
                     * (6.0f * (coarse[ind - cw] + coarse[ind - 1] + 6.0f * coarse[ind] + coarse[ind + 1] + coarse[ind + cw])
                        + coarse[ind - cw - 1] + coarse[ind - cw + 1] + coarse[ind + cw - 1] + coarse[ind + cw + 1]));
    case 1:
      return (float)(4. / 256.
                     * (24.0 * (coarse[ind] + coarse[ind + 1])
                        + 4.0 * (coarse[ind - cw] + coarse[ind - cw + 1] + coarse[ind + cw] + coarse[ind + cw + 1])));
    case 2:
      return (float)(4. / 256.
                     * (24.0 * (coarse[ind] + coarse[ind + cw])
                        + 4.0 * (coarse[ind - 1] + coarse[ind + 1] + coarse[ind + cw - 1] + coarse[ind + cw + 1])));
    default:
      return .25f * (coarse[ind] + coarse[ind + 1] + coarse[ind + cw] + coarse[ind + cw + 1]);
  }
}
/* gauss_expand + ll_fill_boundary2 (:160-171,131-145) == ll_laplacian's clamp (:283-293) */
static inline float expand_clamped(const float *coarse, int i, int j, int wd, int ht)
{
  return expand_at(coarse, clampi(i, 1, ((wd - 1) & ~1) - 1), clampi(j, 1, ((ht - 1) & ~1) - 1), wd);
}

/* gauss_reduce + ll_fill_boundary1, :173-200,120-129 */
static void reduce(const float *input, float *coarse, int wd, int ht)
{
  const int cw = (wd - 1) / 2 + 1, ch = (ht - 1) / 2 + 1;
  const float w[5] = { 1.f / 16.f, 4.f / 16.f, 6.f / 16.f, 4.f / 16.f, 1.f / 16.f };
#pragma omp parallel for schedule(static)
  for(int j = 0; j < ch; j++)
    for(int i = 0; i < cw; i++)
    {
      const int cj = clampi(j, 1, ch - 2), ci = clampi(i, 1, cw - 2);
      float acc = 0.0f;
      if(ch > 2 && cw > 2)
        for(int jj = -2; jj <= 2; jj++)
          for(int ii = -2; ii <= 2; ii++) acc += input[(size_t)(2 * cj + jj) * wd + 2 * ci + ii] * w[ii + 2] * w[jj + 2];
      coarse[(size_t)j * cw + i] = acc;
    }
}

/* curve_scalar, :295-327 */
static inline float curve(float x, float g, float sigma, float shadows, float highlights, float clarity)
{
  const float c = x - g;
  float val;
  if(c > 2 * sigma)
    val = g + sigma + shadows * (c - sigma);
  else if(c < -2 * sigma)
    val = g - sigma + highlights * (c + sigma);
  else if(c > 0.0f)
  {
    const float t = CLAMPS(c / (2.0f * sigma), 0.0f, 1.0f);
    const float t2 = t * t;
    const float mt = 1.0f - t;
    val = g + sigma * 2.0f * mt * t + t2 * (sigma + sigma * shadows);
  }
  else
  {
    const float t = CLAMPS(-c / (2.0f * sigma), 0.0f, 1.0f);
    const float t2 = t * t;
    const float mt = 1.0f - t;
    val = g - sigma * 2.0f * mt * t + t2 * (-sigma - sigma * highlights);
  }
  val += clarity * c * f32m_expf((float)(-c * c / (2.0 * sigma * sigma / 3.0f)));
  return val;
}

/* local_laplacian_internal(), :354-563, b == NULL.  Channel 3 of `out` is not written (the reference leaves
 * it as found, :532-538).  Returns 0, 1 on allocation failure. */
int orc_local_laplacian(const float *input, float *out, int wd, int ht, float sigma, float shadows, float highlights, float clarity)
{
  if(wd <= 1 || ht <= 1) return 0;
  const int mn = wd < ht ? wd : ht;
  int num_levels = 31 - __builtin_clz((unsigned)mn);
  if(num_levels > MAX_LEVELS) num_levels = MAX_LEVELS;
  /* min(wd,ht) in {2,3}: one level, and the reference then reads padded[-1] (:417) -- undefined there,
   * refused here and by the CUDA path */
  if(num_levels < 2) return 2;
  const int last = num_levels - 1;
  const int max_supp = 1 << last;
  const int w = 2 * max_supp + wd, h = 2 * max_supp + ht;
  float *padded[MAX_LEVELS] = { 0 }, *output[MAX_LEVELS] = { 0 }, *buf[NUM_GAMMA][MAX_LEVELS] = { { 0 } };
  int err = 0;
  for(int l = 0; l <= last && !err; l++)
  {
    const size_t n = (size_t)dl(w, l) * dl(h, l);
    if(l < last || last == 0) padded[l] = malloc(sizeof(float) * n);
    output[l] = malloc(sizeof(float) * n);
    err |= (!output[l]) || ((l < last || last == 0) && !padded[l]);
    for(int k = 0; k < NUM_GAMMA; k++)
    {
      buf[k][l] = malloc(sizeof(float) * n);
      err |= !buf[k][l];
    }
  }
  if(err) goto done;

  /* ll_pad_input, replication: value of L*0.01 at clamped coordinates (:262-273 + pad_by_replication) */
#pragma omp parallel for schedule(static)
  for(int j = 0; j < h; j++)
    for(int i = 0; i < w; i++)
    {
      const int sj = clampi(j - max_supp, 0, ht - 1), si = clampi(i - max_supp, 0, wd - 1);
      padded[0][(size_t)j * w + i] = input[4 * ((size_t)sj * wd + si)] * 0.01f;
    }
  for(int l = 1; l < last; l++) reduce(padded[l - 1], padded[l], dl(w, l - 1), dl(h, l - 1));
  if(last >= 1) reduce(padded[last - 1], output[last], dl(w, last - 1), dl(h, last - 1));

  float gamma[NUM_GAMMA];
  for(int k = 0; k < NUM_GAMMA; k++) gamma[k] = (k + .5f) / (float)NUM_GAMMA;
  for(int k = 0; k < NUM_GAMMA; k++)
  {
    /* apply_curve :329-352: curve on the interior, replication outward */
#pragma omp parallel for schedule(static)
    for(int j = 0; j < h; j++)
      for(int i = 0; i < w; i++)
      {
        const int cj = clampi(j, max_supp, h - max_supp - 1), ci = clampi(i, max_supp, w - max_supp - 1);
        buf[k][0][(size_t)j * w + i] = curve(padded[0][(size_t)cj * w + ci], gamma[k], sigma, shadows, highlights, clarity);
      }
    for(int l = 1; l <= last; l++) reduce(buf[k][l - 1], buf[k][l], dl(w, l - 1), dl(h, l - 1));
  }

  for(int l = last - 1; l >= 0; l--)
  {
    const int pw = dl(w, l), ph = dl(h, l);
#pragma omp parallel for schedule(static)
    for(int j = 0; j < ph; j++)
      for(int i = 0; i < pw; i++)
      {
        float o = expand_clamped(output[l + 1], i, j, pw, ph);
        const float v = padded[l][(size_t)j * pw + i];
        int hi = 1;
        for(; hi < NUM_GAMMA - 1 && gamma[hi] <= v; hi++)
          ;
        const int lo = hi - 1;
        const float a = CLAMPS((v - gamma[lo]) / (gamma[hi] - gamma[lo]), 0.0f, 1.0f);
        const float l0 = buf[lo][l][(size_t)j * pw + i] - expand_clamped(buf[lo][l + 1], i, j, pw, ph);
        const float l1 = buf[hi][l][(size_t)j * pw + i] - expand_clamped(buf[hi][l + 1], i, j, pw, ph);
        o += l0 * (1.0f - a) + l1 * a;
        output[l][(size_t)j * pw + i] = o;
      }
  }
#pragma omp parallel for schedule(static)
  for(int j = 0; j < ht; j++)
    for(int i = 0; i < wd; i++)
    {
      out[4 * ((size_t)j * wd + i) + 0] = 100.0f * output[0][(size_t)(j + max_supp) * w + max_supp + i];
      out[4 * ((size_t)j * wd + i) + 1] = input[4 * ((size_t)j * wd + i) + 1];
      out[4 * ((size_t)j * wd + i) + 2] = input[4 * ((size_t)j * wd + i) + 2];
    }
done:
  for(int l = 0; l < MAX_LEVELS; l++)
  {
    free(padded[l]);
    free(output[l]);
    for(int k = 0; k < NUM_GAMMA; k++) free(buf[k][l]);
  }
  return err;
}
