/* CPU restatement of "diffuse or sharpen" (multi-scale anisotropic heat PDE).  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src/iop/diffuse.c: params :76-109, check_isotropy_mode :151-162, tiling_callback
 * :585-610, find_gradients :627-635, rotation matrices :648-677, build_matrix :679-707, isotrope_laplacian
 * :709-725, heat_PDE_diffusion :760-953, compute_anisotropy_factor :955-962, wavelets_process :978-1107,
 * process :1155-1259; pixel/bspline.h: B_SPLINE_SIGMA :38, equivalent_sigma_at_step :52-63,
 * num_steps_to_reach_equivalent_sigma :65-77, sparse_scalar_product :83-117, _bspline_vertical_pass :118-133,
 * _bspline_horizontal :136-151, decompose_2D_Bspline :351-377; math/math.h: dt_fast_hypotf :246-249,
 * dt_fast_expf :254-267; system/simd.h: dt_simd_max_zero :107-114; develop/imageop.c:134-137.
 *
 * Pinned bit-for-bit against the reference's own process() cut verbatim out of diffuse.c (oracle/_ref,
 * ref_diffuse.c), the luminance mask included (threshold > 0: build_mask :1109-1119, inpaint_mask :1122-1152 with
 * iop/noise_generator.h splitmix32 :36-43, xoshiro128plus :54-70, gaussian_noise :82-96 -- glibc logf/sinf/cosf
 * through flt32_math.h).
 */
#include "oracle_common.h"
#include "flt32_math.h"
#include "b200iop.h"
#include <stdlib.h>
#include <string.h>

#define MAX_NUM_SCALES 10
#define B_SPLINE_SIGMA 1.0553651328015339f
#define KAPPA 0.25f

static inline float sqf(float x) { return x * x; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline float max_zero(float v) { return isfinite(v) ? (v > 0.0f ? v : 0.0f) : 0.f; } /* simd.h:107-114 */
static inline float clip0(float v) { return 0.0f > v ? 0.0f : v; }                           /* MAX(0.0f, v) */
static inline float fast_expf(float x)
{ /* math/math.h:254-267: float arithmetic, then truncation */
  const int i1 = 0x3f800000, i2 = 0x402DF854;
  const int k0 = i1 + x * (i2 - i1);
  const int k = k0 > 0 ? k0 : 0;
  float f;
  memcpy(&f, &k, 4);
  return f;
}
float orc_diffuse_sigma_at_step(unsigned s)
{ /* bspline.h:52-63 */
  if(s == 0) return B_SPLINE_SIGMA;
  return sqrtf(sqf(orc_diffuse_sigma_at_step(s - 1)) + sqf(f32m_exp2f((float)s) * B_SPLINE_SIGMA));
}
int orc_diffuse_scales(const b200_diffuse_data_t *d, float zoom)
{ /* diffuse.c:1175-1183, bspline.h:65-77 */
  const float final_radius = (d->radius + d->radius_center) * 2.f / zoom;
  unsigned s = 0;
  float radius = B_SPLINE_SIGMA;
  while(radius < final_radius)
  {
    ++s;
    radius = sqrtf(sqf(radius) + sqf((float)(1 << s) * B_SPLINE_SIGMA));
  }
  const int n = (int)(s + 1);
  return n < 1 ? 1 : (n > MAX_NUM_SCALES ? MAX_NUM_SCALES : n);
}

/* decompose_2D_Bspline(), bspline.h:351-377 */
static void decompose(const float *in, float *HF, float *LF, int width, int height, int mult)
{
  static const float f[5] = { 1.0f / 16.0f, 4.0f / 16.0f, 6.0f / 16.0f, 4.0f / 16.0f, 1.0f / 16.0f };
#pragma omp parallel
  {
    orc_fp_fast_mode();
    float *temp = malloc(sizeof(float) * 4 * (size_t)width);
#pragma omp for
    for(int i = 0; i < height; i++)
    {
      const size_t r[5] = { (size_t)4 * width * imax(i - 2 * mult, 0), (size_t)4 * width * imax(i - mult, 0), (size_t)4 * width * i,
                            (size_t)4 * width * imin(i + mult, height - 1), (size_t)4 * width * imin(i + 2 * mult, height - 1) };
      for(int j = 0; j < width; j++)
        for(int c = 0; c < 4; c++)
        {
          const float *b = in + 4 * (size_t)j + c;
          temp[4 * j + c] = clip0(f[0] * b[r[0]] + f[1] * b[r[1]] + f[2] * b[r[2]] + f[3] * b[r[3]] + f[4] * b[r[4]]);
        }
      for(int j = 0; j < width; j++)
      {
        const int x[5] = { 4 * imax(j - 2 * mult, 0), 4 * imax(j - mult, 0), 4 * j, 4 * imin(j + mult, width - 1), 4 * imin(j + 2 * mult, width - 1) };
        const size_t index = 4 * ((size_t)i * width + j);
        for(int c = 0; c < 4; c++)
        {
          LF[index + c] = clip0(f[0] * temp[x[0] + c] + f[1] * temp[x[1] + c] + f[2] * temp[x[2] + c] + f[3] * temp[x[3] + c] + f[4] * temp[x[4] + c]);
          HF[index + c] = in[index + c] - LF[index + c];
        }
      }
    }
    free(temp);
  }
}

typedef struct
{
  float anisotropy[4];
  int isotropy[4]; /* 0 isotrope, 1 isophote, 2 gradient */
  float variance_threshold, normalized_regularization, ABCD[4], strength;
} pde_t;

/* compute_kernel(), :727-758 */
static inline void make_kernel(float c2, float cs, float cos2, float sin2, int type, float k[9])
{
  if(type == 0)
  {
    k[0] = k[2] = k[6] = k[8] = 0.25f;
    k[1] = k[3] = k[5] = k[7] = 0.5f;
    k[4] = -3.f;
    return;
  }
  float a00, a11, a01;
  if(type == 1)
  { /* rotation_matrix_isophote */
    a00 = cos2 + c2 * sin2;
    a11 = c2 * cos2 + sin2;
    a01 = (c2 - 1.f) * cs;
  }
  else
  { /* rotation_matrix_gradient */
    a00 = c2 * cos2 + sin2;
    a11 = cos2 + c2 * sin2;
    a01 = (1.f - c2) * cs;
  }
  const float b11 = a01 * 0.5f, b13 = -b11, b22 = -2.f * (a00 + a11);
  k[0] = b11;
  k[1] = a11;
  k[2] = b13;
  k[3] = a00;
  k[4] = b22;
  k[5] = a00;
  k[6] = b13;
  k[7] = a11;
  k[8] = b11;
}

/* iop/noise_generator.h */
static inline uint32_t splitmix32(const uint64_t seed)
{ /* :36-43 */
  uint64_t result = (seed ^ (seed >> 33)) * 0x62a9d9ed799705f5ul;
  result = (result ^ (result >> 28)) * 0xcb24d0a5c88c35b3ul;
  return (uint32_t)(result >> 32);
}
static inline float xoshiro128plus(uint32_t state[4])
{ /* :54-70 */
  const uint32_t result = state[0] + state[3];
  const uint32_t t = state[1] << 9;
  state[2] ^= state[0];
  state[3] ^= state[1];
  state[1] ^= state[2];
  state[0] ^= state[3];
  state[2] ^= t;
  state[3] = (state[3] << 11) | (state[3] >> 21);
  return (float)(result >> 8) * 0x1.0p-24f;
}
static inline float gaussian_noise(float mu, float sigma, int flip, uint32_t state[4])
{ /* :82-96: Box-Muller; `2.f * M_PI * u2` is a double product rounded to float at the call */
  const float u1 = fmaxf(xoshiro128plus(state), 1.17549435e-38f);
  const float u2 = xoshiro128plus(state);
  const float radius = sqrtf(-2.0f * f32m_logf(u1));
  const float angle = (float)(2.0 * 3.14159265358979323846 * (double)u2);
  const float noise = flip ? radius * f32m_cosf(angle) : radius * f32m_sinf(angle);
  return noise * sigma + mu;
}
/* build_mask(), :1109-1119 */
static void build_mask(const float *in, uint8_t *mask, float threshold, size_t npx)
{
  for(size_t k = 0; k < npx; k++) mask[k] = (in[4 * k] > threshold || in[4 * k + 1] > threshold || in[4 * k + 2] > threshold);
}
/* inpaint_mask(), :1122-1152: the seed mixes the FLOAT index k with k / width, as the reference writes it */
static void inpaint_mask(float *inpainted, const float *original, const uint8_t *mask, size_t width, size_t height)
{
#pragma omp parallel
  {
    orc_fp_fast_mode();
#pragma omp for
    for(size_t k = 0; k < height * width * 4; k += 4)
    {
      if(mask[k / 4])
      {
        const uint32_t i = (uint32_t)(k / width);
        const uint32_t j = (uint32_t)(k - i);
        uint32_t state[4] = { splitmix32(j + 1), splitmix32((uint64_t)(j + 1) * (i + 3)), splitmix32(1337), splitmix32(666) };
        xoshiro128plus(state);
        xoshiro128plus(state);
        xoshiro128plus(state);
        xoshiro128plus(state);
        for(int c = 0; c < 4; c++) inpainted[k + c] = fabsf(gaussian_noise(original[k + c], original[k + c], i % 2 || j % 2, state));
      }
      else
        for(int c = 0; c < 4; c++) inpainted[k + c] = original[k + c];
    }
  }
}

/* heat_PDE_diffusion(), :760-953; mask == NULL is has_mask == 0 */
static void heat_pde(const float *HF, const float *LF, const uint8_t *mask, float *out, int width, int height, int mult, const pde_t *p)
{
#pragma omp parallel
  {
    orc_fp_fast_mode();
#pragma omp for
    for(int i = 0; i < height; i++)
    {
      const size_t in[3] = { (size_t)imax(i - mult, 0) * width, (size_t)i * width, (size_t)imin(i + mult, height - 1) * width };
      for(int j = 0; j < width; j++)
      {
        const int jn[3] = { imax(j - mult, 0), j, imin(j + mult, width - 1) };
        const float *hf[9], *lf[9];
        for(int ii = 0; ii < 3; ii++)
          for(int jj = 0; jj < 3; jj++)
          {
            hf[3 * ii + jj] = HF + 4 * (in[ii] + jn[jj]);
            lf[3 * ii + jj] = LF + 4 * (in[ii] + jn[jj]);
          }
        float *o = out + 4 * ((size_t)i * width + j);
        if(mask && !mask[(size_t)i * width + j])
        { /* :938-947: outside the mask the scale is only recombined */
          for(int c = 0; c < 4; c++) o[c] = max_zero(hf[4][c] + lf[4][c]);
          continue;
        }
        for(int c = 0; c < 4; c++)
        {
          float energy = 0.f;
          for(int k = 0; k < 9; k++)
          {
            const float safe_lf = max_zero(lf[k][c] - 1e-8f) + 1e-8f;
            const float ratio = hf[k][c] / safe_lf;
            energy += ratio * ratio;
          }
          energy = max_zero(p->variance_threshold + energy * p->normalized_regularization - 1e-8f) + 1e-8f;
          float cs[2], cos2[2], sin2[2], mag[2];
          const float *src[2] = { NULL, NULL };
          for(int g = 0; g < 2; g++)
          { /* g = 0: gradient of LF; g = 1: gradient of HF (the reference's "lapl") */
            const float *const *px = g ? hf : lf;
            (void)src;
            float gx = (px[7][c] - px[1][c]) * 0.5f, gy = (px[5][c] - px[3][c]) * 0.5f;
            const float m = sqrtf(gx * gx + gy * gy);
            const float nonzero = (m != 0.f);
            const float inv_mag = 1.f / (m + (1.f - nonzero));
            gx = gx * inv_mag + (1.f - nonzero);
            gy = gy * inv_mag;
            mag[g] = m;
            cos2[g] = sqf(gx);
            sin2[g] = sqf(gy);
            cs[g] = gx * gy;
          }
          float kern[4][9];
          for(int k = 0; k < 4; k++)
          {
            const int g = k & 1; /* orders 1,3 follow the LF gradient, 2,4 the HF gradient */
            const float c2 = fast_expf(-mag[g] * p->anisotropy[k]);
            make_kernel(c2, cs[g], cos2[g], sin2[g], p->isotropy[k], kern[k]);
          }
          float d[4] = { 0.f, 0.f, 0.f, 0.f };
          for(int k = 0; k < 9; k++)
          {
            d[0] = kern[0][k] * lf[k][c] + d[0];
            d[1] = kern[1][k] * lf[k][c] + d[1];
            d[2] = kern[2][k] * hf[k][c] + d[2];
            d[3] = kern[3][k] * hf[k][c] + d[3];
          }
          float update = d[0] * p->ABCD[0];
          update = d[1] * p->ABCD[1] + update;
          update = d[2] * p->ABCD[2] + update;
          update = d[3] * p->ABCD[3] + update;
          const float acc = hf[4][c] * p->strength + update / energy;
          o[c] = max_zero(acc + lf[4][c]);
        }
      }
    }
  }
}

/* host-side plan of one wavelets_process(): also what the CUDA host code must reproduce */
void orc_diffuse_plan(const b200_diffuse_data_t *d, float zoom, int scales, float *out /* [scales][8] */)
{
  const float regularization = f32m_powf(10.f, d->regularization) - 1.f;
  const float variance_threshold = f32m_powf(10.f, d->variance_threshold);
  for(int s = 0; s < scales; s++)
  {
    const float real_radius = orc_diffuse_sigma_at_step(s) * zoom;
    const float nr = regularization / 9.f * sqf(real_radius);
    const float norm = f32m_expf(-sqf(real_radius - (float)d->radius_center) / sqf(d->radius));
    float *o = out + 8 * s;
    o[0] = variance_threshold;
    o[1] = nr;
    o[2] = d->first * KAPPA * norm;
    o[3] = d->second * KAPPA * norm;
    o[4] = d->third * KAPPA * norm;
    o[5] = d->fourth * KAPPA * norm;
    o[6] = d->sharpness * norm + 1.f;
    o[7] = norm;
  }
}

/* process(), :1155-1259 */
int orc_diffuse(const float *in, float *out, int width, int height, const b200_diffuse_data_t *d, float iscale, float roi_scale)
{
  const size_t n = (size_t)width * height * 4;
  const float zoom = iscale / roi_scale;
  const int iterations = imax((int)ceilf((float)d->iterations), 1);
  const int scales = orc_diffuse_scales(d, zoom);
  float *HF[MAX_NUM_SCALES] = { NULL };
  float *temp1 = malloc(sizeof(float) * n), *temp2 = malloc(sizeof(float) * n);
  float *LF_odd = malloc(sizeof(float) * n), *LF_even = malloc(sizeof(float) * n);
  for(int s = 0; s < scales; s++) HF[s] = malloc(sizeof(float) * n);
  float plan[MAX_NUM_SCALES * 8];
  orc_diffuse_plan(d, zoom, scales, plan);
  pde_t p;
  const float an[4] = { d->anisotropy_first, d->anisotropy_second, d->anisotropy_third, d->anisotropy_fourth };
  for(int k = 0; k < 4; k++)
  {
    p.anisotropy[k] = sqf(an[k]);
    p.isotropy[k] = an[k] == 0.f ? 0 : (an[k] > 0.f ? 1 : 2);
  }
  uint8_t *mask = NULL;
  if(d->threshold > 0.f)
  { /* :1207-1218 */
    mask = malloc((size_t)width * height);
    build_mask(in, mask, d->threshold, (size_t)width * height);
    inpaint_mask(temp1, in, mask, (size_t)width, (size_t)height);
    in = temp1;
  }
  for(int it = 0; it < iterations; it++)
  {
    const float *temp_in = it == 0 ? in : (it % 2 == 0 ? temp1 : temp2);
    float *temp_out = it == 0 ? temp2 : (it % 2 == 0 ? temp2 : temp1);
    if(it == iterations - 1) temp_out = out;
    /* wavelets_process() */
    float *residual = NULL;
    for(int s = 0; s < scales; s++)
    {
      const float *bin = s == 0 ? temp_in : (s % 2 != 0 ? LF_odd : LF_even);
      float *bout = s == 0 ? LF_odd : (s % 2 != 0 ? LF_even : LF_odd);
      decompose(bin, HF[s], bout, width, height, 1 << s);
      residual = bout;
    }
    float *temp = residual == LF_even ? LF_odd : LF_even;
    int count = 0;
    for(int s = scales - 1; s > -1; --s)
    {
      const float *pl = plan + 8 * s;
      p.variance_threshold = pl[0];
      p.normalized_regularization = pl[1];
      for(int k = 0; k < 4; k++) p.ABCD[k] = pl[2 + k];
      p.strength = pl[6];
      const float *bin = count == 0 ? residual : (count % 2 != 0 ? temp : residual);
      float *bout = count == 0 ? temp : (count % 2 != 0 ? residual : temp);
      if(s == 0) bout = temp_out;
      heat_pde(HF[s], bin, mask, bout, width, height, 1 << s, &p);
      count++;
    }
  }
  for(int s = 0; s < scales; s++) free(HF[s]);
  free(mask);
  free(temp1);
  free(temp2);
  free(LF_odd);
  free(LF_even);
  return 0;
}
