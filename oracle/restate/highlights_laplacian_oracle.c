/* oracle/restate -- highlights, mode "guided laplacians" (iop/highlights/laplacian.c process_laplacian :433-575).
 * TEST INFRASTRUCTURE ONLY (see oracle_common.h).
 *
 * Follows: iop/highlights/gather.c  _compute_laplacian_normalization :223-275, _interpolate_and_mask :67-221,
 *          _interpolate_and_mask_passthrough :424-455, _remosaic_and_replace :457-486, _remosaic_and_replace_passthrough :514-541;
 *          pixel/box_filters.c      dt_box_mean_4ch :950-971 (blur_horizontal_4ch :351-404, blur_vertical_1ch :891-913);
 *          pixel/fast_guided_filter.h interpolate_bilinear :99-152;  pixel/bspline.h decompose_2D_Bspline :351-377,
 *          equivalent_sigma_at_step :52-63;  iop/noise_generator.h splitmix32 :36-43, xoshiro128plus :54-70,
 *          poisson_noise_simd :174-200;  iop/highlights/laplacian.c scale_type :76-82, guide_laplacians :85-246,
 *          heat_PDE_diffusion :248-372, wavelets_process :374-430.
 *
 * Pinned against those lines compiled in place (oracle/ref_shim/ref_highlights_laplacian.c) by tests/test_cpu_hl_laplacian.py.
 *
 * The normalization vector is an OpenMP float reduction in the reference: its value depends on the thread count and on the order the
 * threads finish in.  Here it is the sum of ONE thread in row order (what the reference computes with OMP_NUM_THREADS=1) unless the
 * caller imposes a vector (`force`), which is how tests compare with a multi-threaded reference run.
 * Bayer and X-Trans mosaics (gather.c _interpolate_and_mask_xtrans :317-422 with _build_xtrans_bilinear_lookup :277-315,
 * _remosaic_and_replace_xtrans :488-512) and non-mosaic RGBA input.
 */
#include "oracle_common.h"
#include "flt32_math.h"
#include <stdlib.h>
#include <string.h>

#define DS_FACTOR 4
#define MAX_NUM_SCALES 12
#define B_SPLINE_SIGMA 1.0553651328015339f
#define B_SPLINE_TO_LAPLACIAN 3.182727439285017f

static inline float sqf(float x) { return x * x; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline float max_zero(float v) { return isfinite(v) ? (v > 0.0f ? v : 0.0f) : 0.f; } /* simd.h:107-114 */
static inline float clip0(float v) { return 0.0f > v ? 0.0f : v; }                           /* MAX(0.0f, v) */

/* ---- gather ------------------------------------------------------------------------------------------------------------ */
/* FCxtrans(), develop/imageop_math.h:201-219, on a table already turned to the ROI origin */
static inline int fcx(int row, int col, const uint8_t xt[6][6]) { return xt[(row + 600) % 6][(col + 600) % 6]; }

static void normalization_of(const float *in, int width, int height, uint32_t filters, const uint8_t xt[6][6], float norm[4])
{ /* gather.c:223-275, one thread */
  float sum[3] = { 0.f, 0.f, 0.f };
  const float n_pixels = (float)(height * width);
  for(int i = 0; i < height; i++)
    for(int j = 0; j < width; j++)
    {
      if(!filters)
      {
        const float *p = in + 4 * ((size_t)i * width + j);
        for(int c = 0; c < 3; c++) sum[c] += p[c] / n_pixels;
      }
      else
      {
        const int c = filters == 9u ? fcx(i, j, xt) : orc_fc(i, j, filters);
        if(c < 0 || c > 2) continue;
        sum[c] += in[(size_t)i * width + j] / n_pixels;
      }
    }
  norm[0] = sum[0];
  norm[1] = sum[1];
  norm[2] = sum[2];
  norm[3] = 1.f;
}

/* one site of _interpolate_and_mask: the colour c2 (0 or 2) read around a site of another colour */
static inline void around(const float *in, size_t ic, size_t ip, size_t in_, int j, int jp, int jn, int i, uint32_t filters, int c2, float clip,
                          float *value, int *clipped)
{
  const float north = in[ip + j], south = in[in_ + j], west = in[ic + jp], east = in[ic + jn];
  if(orc_fc(i + 1, j, filters) == c2)
  {
    *value = (north + south) / 2.f;
    *clipped = north > clip || south > clip;
  }
  else if(orc_fc(i, j + 1, filters) == c2)
  {
    *value = (west + east) / 2.f;
    *clipped = west > clip || east > clip;
  }
  else
  {
    const float nw = in[ip + jp], ne = in[ip + jn], se = in[in_ + jn], sw = in[in_ + jp];
    *value = (nw + ne + se + sw) / 4.f;
    *clipped = nw > clip || ne > clip || sw > clip || se > clip;
  }
}
static void gather_bayer(const float *in, float *interpolated, float *mask, const float clips[4], const float wb[4], uint32_t filters, int width,
                         int height)
{ /* gather.c:67-221 with det_scale = 1 (:510) */
  float cl[4];
  for(int c = 0; c < 4; c++) cl[c] = clips[c] * 1.f;
#pragma omp parallel for
  for(int i = 0; i < height; i++)
    for(int j = 0; j < width; j++)
    {
      const int c = orc_fc(i, j, filters);
      const size_t ic = (size_t)i * width, ip = (size_t)(i == 0 ? 1 : i - 1) * width, in_ = (size_t)(i == height - 1 ? height - 2 : i + 1) * width;
      const int jp = j == 0 ? 1 : j - 1, jn = j == width - 1 ? width - 2 : j + 1;
      const float center = in[ic + j];
      float v[3];
      int k[3];
      if(c == 1)
      {
        v[1] = center;
        k[1] = center > cl[1];
      }
      else
      {
        const float north = in[ip + j], south = in[in_ + j], west = in[ic + jp], east = in[ic + jn];
        v[1] = (north + south + east + west) / 4.f;
        k[1] = north > cl[1] || south > cl[1] || east > cl[1] || west > cl[1];
      }
      for(int c2 = 0; c2 < 3; c2 += 2)
        if(c == c2)
        {
          v[c2] = center;
          k[c2] = center > cl[c2];
        }
        else
          around(in, ic, ip, in_, j, jp, jn, i, filters, c2, cl[c2], &v[c2], &k[c2]);
      const float rgb[4] = { v[0], v[1], v[2], sqrtf(sqf(v[0]) + sqf(v[1]) + sqf(v[2])) };
      const float flags[4] = { (float)k[0], (float)k[1], (float)k[2], (float)(k[0] || k[1] || k[2]) };
      const size_t idx = 4 * (ic + j);
      for(int q = 0; q < 4; q++)
      {
        interpolated[idx + q] = fmaxf(rgb[q] / wb[q], 0.f);
        mask[idx + q] = flags[q];
      }
    }
}
static void gather_xtrans(const float *in, float *interpolated, float *mask, const float clips[4], const float wb[4], const uint8_t xt[6][6],
                          int width, int height)
{ /* gather.c:317-422; the interior walks what _build_xtrans_bilinear_lookup :277-315 lists: the eight neighbours row by row, those of the
   * site's own colour skipped, weights 2 on the cross and 1 on the corners */
#pragma omp parallel for
  for(int i = 0; i < height; i++)
    for(int j = 0; j < width; j++)
    {
      const size_t idx = (size_t)i * width + j;
      const float center = in[idx];
      const int f = fcx(i, j, xt);
      float rgb[4] = { 0.f, 0.f, 0.f, 0.f }, flags[4] = { 0.f, 0.f, 0.f, 0.f };
      float sum[3] = { 0.f, 0.f, 0.f };
      int used_clipped[3] = { 0, 0, 0 };
      if(i == 0 || j == 0 || i == height - 1 || j == width - 1)
      { /* the border ring: plain means of whatever the 3x3 window holds inside the frame, the site itself included */
        int count[3] = { 0, 0, 0 };
        for(int y = imax(i - 1, 0); y <= imin(i + 1, height - 1); y++)
          for(int x = imax(j - 1, 0); x <= imin(j + 1, width - 1); x++)
          {
            const int color = fcx(y, x, xt);
            const float value = in[(size_t)y * width + x];
            sum[color] += value;
            count[color]++;
            used_clipped[color] |= value > clips[color];
          }
        for(int c = 0; c < 3; c++)
        {
          const int own = c == f || count[c] == 0;
          rgb[c] = own ? center : sum[c] / count[c];
          flags[c] = own ? (float)(center > clips[c]) : (float)used_clipped[c];
        }
      }
      else
      {
        int total[3] = { 0, 0, 0 };
        for(int y = -1; y <= 1; y++)
          for(int x = -1; x <= 1; x++)
          {
            const int color = fcx(i + y, j + x, xt);
            if(color == f) continue;
            const int weight = 1 << ((y == 0) + (x == 0));
            const float value = in[(size_t)(i + y) * width + (j + x)];
            sum[color] += value * weight;
            total[color] += weight;
            used_clipped[color] |= value > clips[color];
          }
        for(int c = 0; c < 3; c++)
          if(c != f)
          {
            rgb[c] = total[c] > 0 ? sum[c] / total[c] : center;
            flags[c] = (float)used_clipped[c];
          }
        rgb[f] = center;
        flags[f] = (float)(center > clips[f]);
      }
      rgb[3] = sqrtf(sqf(rgb[0]) + sqf(rgb[1]) + sqf(rgb[2]));
      flags[3] = (float)(flags[0] || flags[1] || flags[2]);
      for(int q = 0; q < 4; q++)
      {
        interpolated[4 * idx + q] = fmaxf(rgb[q] / wb[q], 0.f);
        mask[4 * idx + q] = flags[q];
      }
    }
}
static void gather_rgba(const float *in, float *interpolated, float *mask, const float clips[4], const float wb[4], size_t npx)
{ /* gather.c:424-455 */
#pragma omp parallel for
  for(size_t p = 0; p < npx; p++)
  {
    const float R = in[4 * p], G = in[4 * p + 1], B = in[4 * p + 2];
    const int kr = R > clips[0], kg = G > clips[1], kb = B > clips[2];
    const float rgb[4] = { R, G, B, sqrtf(sqf(R) + sqf(G) + sqf(B)) };
    const float flags[4] = { (float)kr, (float)kg, (float)kb, (float)(kr || kg || kb) };
    for(int q = 0; q < 4; q++)
    {
      interpolated[4 * p + q] = fmaxf(rgb[q] / wb[q], 0.f);
      mask[4 * p + q] = flags[q];
    }
  }
}

/* ---- box mean of radius 2 over four interleaved channels, in place (box_filters.c:950-971) ---------------------------- */
static void box_rows(float *buf, int height, int width, int radius)
{ /* blur_horizontal_4ch :351-404: a running sum per channel, entering and leaving samples in row order */
#pragma omp parallel
  {
    float *scratch = malloc(sizeof(float) * 4 * (size_t)width);
#pragma omp for
    for(int y = 0; y < height; y++)
    {
      float *row = buf + (size_t)4 * y * width;
      float L[4] = { 0, 0, 0, 0 };
      size_t hits = 0;
      memcpy(scratch, row, sizeof(float) * 4 * (size_t)width);
      for(int x = 0; x < imin(radius, width); x++, hits++)
        for(int c = 0; c < 4; c++) L[c] += scratch[4 * x + c];
      for(int x = 0; x < width; x++)
      {
        if(x > radius)
          for(int c = 0; c < 4; c++) L[c] -= scratch[4 * (x - radius - 1) + c];
        if(x > radius && x + radius >= width) hits--;
        if(x + radius < width)
        {
          for(int c = 0; c < 4; c++) L[c] += scratch[4 * (x + radius) + c];
          if(x <= radius) hits++;
        }
        for(int c = 0; c < 4; c++) row[4 * x + c] = L[c] / (float)hits;
      }
    }
    free(scratch);
  }
}
static void box_columns(float *buf, int height, int width, int radius)
{ /* blur_vertical_1ch :891-913 over 4*width float columns; every lane of the 16-, 4- and 1-wide variants (:509-574, :646-702, :767-825)
   * runs the same recurrence.  `hits` is a float in the 16-wide variant and a size_t in the others: the quotient is the same */
  const size_t stride = (size_t)4 * width;
#pragma omp parallel for
  for(size_t x = 0; x < stride; x++)
  {
    float *col = buf + x;
    float *keep = malloc(sizeof(float) * (size_t)height);
    for(int y = 0; y < height; y++) keep[y] = col[(size_t)y * stride];
    float L = 0.0f;
    int hits = 0;
    for(int y = 0; y < imin(radius, height); y++, hits++) L += keep[y];
    for(int y = 0; y < height; y++)
    {
      if(y > radius) L -= keep[y - radius - 1];
      if(y > radius && y + radius >= height) hits--;
      if(y + radius < height)
      {
        L += keep[y + radius];
        if(y <= radius) hits++;
      }
      col[(size_t)y * stride] = L / (float)hits;
    }
    free(keep);
  }
}

/* ---- interpolate_bilinear(), fast_guided_filter.h:99-152, four channels ------------------------------------------------ */
static void bilinear(const float *in, int width_in, int height_in, float *out, int width_out, int height_out)
{
#pragma omp parallel for
  for(int i = 0; i < height_out; i++)
    for(int j = 0; j < width_out; j++)
    {
      const float x_out = (float)j / (float)width_out, y_out = (float)i / (float)height_out;
      const float x_in = x_out * (float)width_in, y_in = y_out * (float)height_in;
      size_t x_prev = (size_t)floorf(x_in), y_prev = (size_t)floorf(y_in);
      size_t x_next = x_prev + 1, y_next = y_prev + 1;
      x_prev = x_prev < (size_t)width_in ? x_prev : (size_t)width_in - 1;
      x_next = x_next < (size_t)width_in ? x_next : (size_t)width_in - 1;
      y_prev = y_prev < (size_t)height_in ? y_prev : (size_t)height_in - 1;
      y_next = y_next < (size_t)height_in ? y_next : (size_t)height_in - 1;
      const float *nw = in + 4 * (y_prev * width_in + x_prev), *ne = in + 4 * (y_prev * width_in + x_next);
      const float *se = in + 4 * (y_next * width_in + x_next), *sw = in + 4 * (y_next * width_in + x_prev);
      const float Dy_next = (float)y_next - y_in, Dy_prev = 1.f - Dy_next;
      const float Dx_next = (float)x_next - x_in, Dx_prev = 1.f - Dx_next;
      float *o = out + 4 * ((size_t)i * width_out + j);
      for(int c = 0; c < 4; c++) o[c] = Dy_prev * (sw[c] * Dx_next + se[c] * Dx_prev) + Dy_next * (nw[c] * Dx_next + ne[c] * Dx_prev);
    }
}

/* ---- decompose_2D_Bspline(), bspline.h:351-377 ------------------------------------------------------------------------ */
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
      for(int j = 0; j < 4 * width; j++)
      {
        const float *b = in + j;
        temp[j] = clip0(f[0] * b[r[0]] + f[1] * b[r[1]] + f[2] * b[r[2]] + f[3] * b[r[3]] + f[4] * b[r[4]]);
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
static float sigma_at_step(unsigned s)
{ /* bspline.h:52-63 */
  if(s == 0) return B_SPLINE_SIGMA;
  return sqrtf(sqf(sigma_at_step(s - 1)) + sqf(f32m_exp2f((float)s) * B_SPLINE_SIGMA));
}

/* ---- iop/noise_generator.h ------------------------------------------------------------------------------------------------ */
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
/* poisson_noise_simd :174-200 for the three colour lanes (the fourth is overwritten by the caller's norm) */
static void poisson3(const float mu[3], const float sigma[3], uint32_t state[4], float out[3])
{
  float u1[3], u2[3];
  for(int c = 0; c < 3; c++)
  {
    u1[c] = fmaxf(xoshiro128plus(state), 1.17549435e-38f);
    u2[c] = xoshiro128plus(state);
  }
  for(int c = 0; c < 3; c++)
  {
    const float radius = sqrtf(-2.0f * f32m_logf(u1[c]));
    const float angle = (float)(2.0 * 3.14159265358979323846 * (double)u2[c]); /* `2.f * M_PI * u2` is a double product */
    const float noise = (c != 1) ? radius * f32m_cosf(angle) : radius * f32m_sinf(angle); /* flip = { 1, 0, 1, 0 } */
    const float r = noise * sigma[c] + 2.0f * sqrtf(fmaxf(mu[c] + 3.f / 8.f, 0.0f));
    out[c] = (r * r - sigma[c] * sigma[c]) / 4.f - 3.f / 8.f;
  }
}

/* ---- the two reconstructions of one wavelet scale ----------------------------------------------------------------------- */
enum { FIRST_SCALE = 2, LAST_SCALE = 4 };
static void guide_laplacians(const float *HF, const float *LF, const float *mask, float *out, int width, int height, int mult, float noise_level,
                             int salt, int scale, float radius_sq)
{ /* laplacian.c:85-246 */
  const float inv_patch = 1.f / 9.f, scale_multiplier = 1.f / radius_sq, eps = 1e-12f;
#pragma omp parallel
  {
    orc_fp_fast_mode();
#pragma omp for
    for(int i = 0; i < height; i++)
    {
      const size_t rows[3] = { (size_t)imax(i - mult, 0) * width, (size_t)i * width, (size_t)imin(i + mult, height - 1) * width };
      for(int j = 0; j < width; j++)
      {
        const size_t index = 4 * ((size_t)i * width + j);
        const float alpha = mask[index + 3], alpha_comp = 1.f - alpha;
        float hf[4] = { HF[index], HF[index + 1], HF[index + 2], HF[index + 3] };
        if(alpha > 0.f)
        {
          const int cols[3] = { imax(j - mult, 0), j, imin(j + mult, width - 1) };
          float sum[4] = { 0 }, sum_sq[4] = { 0 }, prod[3][4] = { { 0 } };
          for(int jj = 0; jj < 3; jj++)
            for(int ii = 0; ii < 3; ii++)
            {
              const float *s = HF + 4 * (rows[jj] + cols[ii]);
              for(int c = 0; c < 4; c++)
              {
                sum[c] += s[c];
                sum_sq[c] += s[c] * s[c];
                for(int g = 0; g < 3; g++) prod[g][c] += s[c] * s[g];
              }
            }
          float means[4], variance[4];
          for(int c = 0; c < 4; c++)
          {
            means[c] = sum[c] * inv_patch;
            variance[c] = max_zero(sum_sq[c] * inv_patch - means[c] * means[c]);
          }
          variance[3] = 0.f;
          int g = 0;
          float guide_variance = variance[0];
          if(variance[1] > guide_variance)
          {
            g = 1;
            guide_variance = variance[1];
          }
          if(variance[2] > guide_variance)
          {
            g = 2;
            guide_variance = variance[2];
          }
          if(guide_variance > eps)
          {
            const float guide_mean = means[g], guide = hf[g];
            for(int c = 0; c < 4; c++)
            {
              const float covariance = prod[g][c] * inv_patch - means[c] * guide_mean;
              const float slope = max_zero(covariance / guide_variance);
              const float intercept = means[c] - slope * guide_mean;
              const float blend = mask[index + c] * scale_multiplier;
              hf[c] = blend * (slope * guide + intercept) + (1.f - blend) * hf[c];
            }
          }
        }
        float px[4];
        for(int c = 0; c < 4; c++)
        {
          px[c] = (scale & FIRST_SCALE) ? hf[c] : hf[c] + out[index + c];
          if(scale & LAST_SCALE) px[c] = max_zero(px[c] + LF[index + c]);
        }
        if((scale & LAST_SCALE) && salt && alpha > 0.f)
        {
          uint32_t state[4] = { splitmix32((uint64_t)(j + 1)), splitmix32((uint64_t)(j + 1) * (uint64_t)(i + 3)), splitmix32(1337), splitmix32(666) };
          for(int k = 0; k < 4; k++) xoshiro128plus(state);
          const float sigma[3] = { px[0] * noise_level, px[1] * noise_level, px[2] * noise_level };
          float noise[3];
          poisson3(px, sigma, state, noise);
          for(int c = 0; c < 3; c++)
          {
            const float noisy = px[c] + fabsf(noise[c] - px[c]);
            px[c] = fmaxf(alpha * noisy + alpha_comp * px[c], 0.f);
          }
        }
        if(scale & LAST_SCALE)
        { /* ratios and norm for the second reconstruction */
          const float norm = fmaxf(sqrtf(sqf(px[0]) + sqf(px[1]) + sqf(px[2])), 1e-6f);
          for(int c = 0; c < 3; c++) px[c] /= norm;
          px[3] = norm;
        }
        for(int c = 0; c < 4; c++) out[index + c] = px[c];
      }
    }
  }
}
static void heat_pde(const float *HF, const float *LF, const float *mask, float *out, int width, int height, int mult, int scale,
                     float first_order_factor)
{ /* laplacian.c:248-372 */
  static const float kernel[9] = { 0.25f, 0.5f, 0.25f, 0.5f, -3.f, 0.5f, 0.25f, 0.5f, 0.25f };
  const float multipliers[4] = { 1.f / B_SPLINE_TO_LAPLACIAN, 1.f / B_SPLINE_TO_LAPLACIAN, 1.f / B_SPLINE_TO_LAPLACIAN, 0.f };
#pragma omp parallel
  {
    orc_fp_fast_mode();
#pragma omp for
    for(int i = 0; i < height; i++)
    {
      const size_t rows[3] = { (size_t)imax(i - mult, 0) * width, (size_t)i * width, (size_t)imin(i + mult, height - 1) * width };
      for(int j = 0; j < width; j++)
      {
        const size_t index = 4 * ((size_t)i * width + j);
        const float *alpha = mask + index;
        float hf[4] = { HF[index], HF[index + 1], HF[index + 2], HF[index + 3] };
        const float norm_backup = hf[3];
        if(alpha[3] > 0.f)
        {
          const int cols[3] = { imax(j - mult, 0), j, imin(j + mult, width - 1) };
          float lap[4] = { 0.f, 0.f, 0.f, 0.f };
          for(int k = 0; k < 9; k++)
          {
            const float *s = HF + 4 * (rows[k / 3] + cols[k % 3]);
            for(int c = 0; c < 4; c++) lap[c] += s[c] * kernel[k];
          }
          for(int c = 0; c < 4; c++) hf[c] += alpha[c] * multipliers[c] * (lap[c] - first_order_factor * hf[c]);
          hf[3] = norm_backup;
        }
        float px[4];
        for(int c = 0; c < 4; c++) px[c] = (scale & FIRST_SCALE) ? hf[c] : out[index + c] + hf[c];
        if(scale & LAST_SCALE)
        {
          for(int c = 0; c < 4; c++) px[c] = fmaxf(px[c] + LF[index + c], 0.f);
          if(alpha[3] > 0.f)
          {
            const float norm = sqrtf(sqf(px[0]) + sqf(px[1]) + sqf(px[2]));
            if(norm > 1e-4f)
              for(int c = 0; c < 3; c++) px[c] /= norm;
          }
          for(int c = 0; c < 3; c++) px[c] = px[c] * px[3];
        }
        for(int c = 0; c < 4; c++) out[index + c] = px[c];
      }
    }
  }
}
static void wavelets(const float *in, float *reconstructed, const float *mask, int width, int height, int scales, float *HF, float *LF_odd,
                     float *LF_even, int chroma, float noise_level, int salt, float first_order_factor)
{ /* laplacian.c:374-430 */
  for(int s = 0; s < scales; s++)
  {
    const float *buffer_in = s == 0 ? in : ((s & 1) ? LF_odd : LF_even);
    float *buffer_out = s == 0 ? LF_odd : ((s & 1) ? LF_even : LF_odd);
    decompose(buffer_in, HF, buffer_out, width, height, 1 << s);
    const int type = 1 | (s == 0 ? FIRST_SCALE : 0) | (s == scales - 1 ? LAST_SCALE : 0);
    const float radius = sqf(sigma_at_step((unsigned)(s * DS_FACTOR)));
    if(!chroma)
      guide_laplacians(HF, buffer_out, mask, reconstructed, width, height, 1 << s, noise_level, salt, type, radius);
    else
      heat_pde(HF, buffer_out, mask, reconstructed, width, height, 1 << s, type, first_order_factor);
  }
}

/* the number of wavelet scales of a run, laplacian.c:461-463 */
int orc_hl_laplacian_scales(int scales_param, float iscale, float roi_scale)
{
  const float scale = DS_FACTOR * (iscale / roi_scale);
  const float final_radius = (float)((int)(1 << scales_param)) / scale;
  const int n = (int)ceilf(f32m_log2f(final_radius));
  return n < 1 ? 1 : (n > MAX_NUM_SCALES ? MAX_NUM_SCALES : n);
}

int orc_hl_laplacian(const float *in, float *out, int x, int y, int width, int height, uint32_t filters, const uint8_t xtrans[36],
                     const float clips[4], int iterations, int scales_param, float noise_level_param, float solid_color, float iscale,
                     float roi_scale, float normalization[4], int force)
{
  const uint32_t shifted = orc_roi_filters(filters, x, y);
  uint8_t xt[6][6] = { { 0 } };
  if(filters == 9u)
  {
    if(!xtrans) return 2;
    for(int r = 0; r < 6; r++)
      for(int c = 0; c < 6; c++) xt[r][c] = xtrans[6 * ((r + y + 600) % 6) + (c + x + 600) % 6];
  }
  const size_t npx = (size_t)width * height;
  const int ds_width = width / DS_FACTOR, ds_height = height / DS_FACTOR;
  const size_t ds_npx = (size_t)ds_width * ds_height;
  const float scale = DS_FACTOR * (iscale / roi_scale);
  const int scales = orc_hl_laplacian_scales(scales_param, iscale, roi_scale);
  const float noise_level = noise_level_param / scale;

  float *interpolated = malloc(sizeof(float) * 4 * npx), *mask = malloc(sizeof(float) * 4 * npx);
  float *ds[6];
  for(int k = 0; k < 6; k++) ds[k] = malloc(sizeof(float) * 4 * (ds_npx ? ds_npx : 1));
  float *LF_odd = ds[0], *LF_even = ds[1], *temp = ds[2], *HF = ds[3], *ds_interpolated = ds[4], *ds_mask = ds[5];

  float wb[4];
  if(force)
    memcpy(wb, normalization, sizeof(wb));
  else
  {
    normalization_of(in, width, height, shifted, xt, wb);
    memcpy(normalization, wb, sizeof(wb));
  }
  if(shifted == 9u)
    gather_xtrans(in, interpolated, mask, clips, wb, xt, width, height);
  else if(shifted)
    gather_bayer(in, interpolated, mask, clips, wb, shifted, width, height);
  else
    gather_rgba(in, interpolated, mask, clips, wb, npx);
  box_rows(mask, height, width, 2);
  box_columns(mask, height, width, 2);
  bilinear(mask, width, height, ds_mask, ds_width, ds_height);
  bilinear(interpolated, width, height, ds_interpolated, ds_width, ds_height);
  for(int i = 0; i < iterations; i++)
  {
    const int salt = i == iterations - 1;
    wavelets(ds_interpolated, temp, ds_mask, ds_width, ds_height, scales, HF, LF_odd, LF_even, 0, noise_level, salt, solid_color);
    wavelets(temp, ds_interpolated, ds_mask, ds_width, ds_height, scales, HF, LF_odd, LF_even, 1, noise_level, salt, solid_color);
  }
  bilinear(ds_interpolated, ds_width, ds_height, interpolated, width, height);
#pragma omp parallel for
  for(size_t p = 0; p < npx; p++)
  { /* gather.c:457-486 / :488-512 / :514-541 with clip_is_floor = FALSE */
    const float opacity = mask[4 * p + 3];
    if(shifted)
    {
      const int c = shifted == 9u ? fcx((int)(p / (size_t)width), (int)(p % (size_t)width), xt) : orc_fc(p / (size_t)width, p % (size_t)width, shifted);
      const float reconstructed = fmaxf(interpolated[4 * p + c] * wb[c], 0.f);
      out[p] = opacity * reconstructed + (1.f - opacity) * in[p];
    }
    else
      for(int c = 0; c < 4; c++)
      {
        if(c == 3)
        {
          out[4 * p + 3] = in[4 * p + 3];
          continue;
        }
        const float own = mask[4 * p + c]; /* a pixel carries all three colours: each blends under its own mask */
        const float reconstructed = fmaxf(interpolated[4 * p + c] * wb[c], 0.f);
        out[4 * p + c] = own * reconstructed + (1.f - own) * in[4 * p + c];
      }
  }
  for(int k = 0; k < 6; k++) free(ds[k]);
  free(interpolated);
  free(mask);
  return 0;
}
