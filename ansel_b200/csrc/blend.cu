// Blending of a module's output over its input (mask + blend operator): scene-referred RGB, display-referred RGB, Lab and raw, for B200 / sm_100a.
//
// What the reference computes: develop/blend.c dt_develop_blend_process :657-860 with blend_cst == DEVELOP_BLEND_CS_RGB_SCENE,
// DEVELOP_BLEND_CS_RGB_DISPLAY (develop/blends/blendif_rgb_hsl.c: the same masks with H, S, L for Jz, Cz, hz, its 27 operators) or
// DEVELOP_BLEND_CS_LAB: the mask (uniform opacity | the raster / drawn mask the host rasterised | the parametric mask of
// develop/blends/blendif_rgb_jzczhz.c :42-325 on the gray, red, green, blue, Jz, Cz and hz channels -- of develop/blends/blendif_lab.c :56-298 on
// the L, a, b, chroma and hue channels -- of the module's input and output, combined exclusively or inclusively, inverted or not | the mask tone
// curve :626-655), then one of the sixteen blend operators of the RGB space (blendif_rgb_jzczhz.c :328-649) or one of the twenty-six
// of the Lab space (blendif_lab.c :302-1068), the result in place of the module's output with the mask in its alpha
// lane.  Parity contract: bit-identical to those lines under C float semantics (oracle/restate/blend_oracle.c, pinned against them
// compiled in place).
//
// The reference makes up to seven passes over full buffers (seed, one per parametric channel set, opacity, tone curve, a copy of
// the output, the operator, the alpha copy).  Every one of them is pointwise, so the kernel is ONE pass: 16 B of input, 16 B of
// output and 4 B of form mask in, 16 B out (+ 4 B when the caller wants the mask, e.g. to publish it as a raster mask) --
// 52 B/px algorithmic.  What the host decides once per call (which of the reference's branches a parameter block takes, the
// slopes of the parametric channels, exp2f / expf of the parameters) arrives in the plan; what depends on the pixel is evaluated
// here.  Not built (B200_ERR_UNSUPPORTED, the caller falls back to the reference's own path): feathering (guided filter), Gaussian
// blur and detail refinement of the mask; the GUI's channel display.  The raw space (develop/blends/blendif_raw.c, one float per site, no
// parametric channels, 16 operators) has a kernel of its own, blend_raw_kernel.  The Jz, Cz, hz channels of
// the RGB space, the chroma and hue channels and the four LCh operators of the Lab space go through glibc's powf / atan2f / hypotf / cosf /
// sinf as restated in flt32_math.cuh.
#ifndef B200_KERNELS_ON_CPU // tests/emul compiles the kernel of this file with g++ to check it against the oracle without a GPU
#include "runtime.h"
#endif
#include "flt32_math.cuh"
#include <float.h>
#include <math.h>
#include <string.h>

namespace
{
enum
{
  MASK_ENABLED = 1, MASK_SHAPE = 2, MASK_PARAMETRIC = 4, MASK_RASTER = 8, // dt_develop_mask_mode_t, blend.h:110-118
  COMBINE_INV = 1, COMBINE_INCL = 2,                                      // dt_develop_mask_combine_mode_t :120-131
  BLENDIF_SIZE = 16, BLENDIF_ITEMS = 6, BLENDIF_RGB_MASK = 0x77FF, BLENDIF_LAB_MASK = 0x3377, // :188-191, :329
  CS_RAW = 1, CS_LAB = 2, CS_RGB_DISPLAY = 3, CS_RGB_SCENE = 4             // :52-59
};
constexpr unsigned BLEND_REVERSE = 0x80000000u; // blend.h:106

struct blend_plan_t
{
  const float4 *in;
  float4 *out;
  const float *form;
  float *mask_out;
  int iw, ow, oh, xoffs, yoffs;
  int lab;       // the Lab space (develop/blends/blendif_lab.c) instead of scene-referred RGB
  int display;   // the display-referred RGB space (develop/blends/blendif_rgb_hsl.c)
  int raw;       // the raw space (develop/blends/blendif_raw.c): `in` and `out` hold one float per site
  int kind;      // 0: mask = opacity; 1: mask = form * opacity (a raster mask alone); 2: seed, then the parametric stage
  int seed_form; // kind 2: the seed is the form mask, else `fill`
  float fill, opacity;
  int pm;        // parametric stage: 0 = opacity * m (or opacity * (1 - m) inverted); 1 = the constant pm_const; 2 = channels
  int inversed, inclusive;
  float pm_const;
  unsigned blendif;
  float par[BLENDIF_ITEMS * BLENDIF_SIZE]; // per channel (blend.h:132-187): four limits and two slopes
  float c_scale;                           // Lab: 1 / (128 * sqrt(2)), the scale of the chroma channel
  float masking[9];                        // RGB: matrix_out of the masking profile (RGB -> XYZ D65), row by row
  float lum[3];
  int tone;
  float contrast_e, brightness;
  unsigned mode;
  int reverse, keep_alpha;
  float p;
};

__device__ __forceinline__ float bl_factor(float value, unsigned invert, const float *p)
{ // _blendif_compute_factor(), :42-73
  float f;
  if(value <= p[0])
    f = 0.0f;
  else if(value < p[1])
    f = (value - p[0]) * p[4];
  else if(value <= p[2])
    f = 1.0f;
  else if(value < p[3])
    f = 1.0f - (value - p[2]) * p[5];
  else
    f = 0.0f;
  return invert ? 1.0f - f : f;
}
__device__ __forceinline__ float bl_channels(const float px[4], float t, unsigned blendif, const float *par, const float *lum)
{ // _blendif_combine_channels(), :151-185: gray, red, green, blue; bl_jzczhz() is the rest of it
  if(blendif & 1u) t *= bl_factor(lum[0] * px[0] + lum[1] * px[1] + lum[2] * px[2], (blendif >> 16) & 1u, par);
#pragma unroll
  for(int c = 0; c < 3; c++)
    if(blendif & (2u << c)) t *= bl_factor(px[c], (blendif >> 16) & (2u << c), par + BLENDIF_ITEMS * (1 + c));
  return t;
}
__device__ __forceinline__ float bl_divc(float a, float b)
{ // IEEE division by a constant that is not a power of two: opaque to nvcc's x / c -> x * (1 / c) rewrite under -ftz=true
#ifdef B200_KERNELS_ON_CPU
  return a / b;
#else
  float q;
  asm("div.rn.ftz.f32 %0, %1, %2;" : "=f"(q) : "f"(a), "f"(b));
  return q;
#endif
}
// dt_Lab_2_LCH / dt_LCH_2_Lab, common/colorspaces_inline_conversions.h:594-615, on glibc's atan2f / hypotf / cosf / sinf (flt32_math.cuh)
__device__ __forceinline__ void bl_lab_to_lch(const float lab[3], float lch[3])
{
  const float two_pi = 2.0f * 3.14159265358979324f;
  float h = f32m::atan2f_(lab[2], lab[1]);
  if(h > 0.0f)
    h = bl_divc(h, two_pi);
  else
    h = 1.0f - bl_divc(fabsf(h), two_pi);
  lch[0] = lab[0];
  lch[1] = f32m::hypotf_(lab[1], lab[2]);
  lch[2] = h;
}
__device__ __forceinline__ void bl_lch_to_lab(const float lch[3], float lab[3])
{
  const float two_pi = 2.0f * 3.14159265358979324f;
  lab[0] = lch[0];
  lab[1] = f32m::cosf_(two_pi * lch[2]) * lch[1];
  lab[2] = f32m::sinf_(two_pi * lch[2]) * lch[1];
}
__device__ __forceinline__ float bl_channels_lab(const float px[4], float t, unsigned blendif, const float *par, float c_scale)
{ // blendif_lab.c _blendif_combine_channels :139-173: L / 100, a / 256, b / 256, then chroma and hue together
  if(blendif & 1u) t *= bl_factor(bl_divc(px[0], 100.0f), (blendif >> 16) & 1u, par);
  if(blendif & 2u) t *= bl_factor(px[1] / 256.0f, (blendif >> 16) & 2u, par + BLENDIF_ITEMS);
  if(blendif & 4u) t *= bl_factor(px[2] / 256.0f, (blendif >> 16) & 4u, par + BLENDIF_ITEMS * 2);
  if(blendif & 0x300u)
  {
    float lch[3], factor = 1.0f;
    bl_lab_to_lch(px, lch);
    factor *= bl_factor(lch[1] * c_scale, (blendif >> 16) & 0x100u, par + BLENDIF_ITEMS * 8);
    factor *= bl_factor(lch[2], (blendif >> 16) & 0x200u, par + BLENDIF_ITEMS * 9);
    t *= factor;
  }
  return t;
}
// the Jz, Cz, hz channels of the RGB space, blendif_rgb_jzczhz.c:122-149: XYZ D65 through the masking profile's matrix (dt_mat3x4_mul_vec4,
// system/simd.h:189-197), dt_XYZ_2_JzAzBz and dt_JzAzBz_2_JzCzhz (common/colorspaces_inline_conversions.h:672-722, :775-781) on glibc's powf,
// atan2f and hypotf (flt32_math.cuh); the three factors multiplied together, then into the mask
__device__ __forceinline__ float bl_jzczhz(const float px[4], float t, unsigned blendif, const float *par, const float *mo)
{
  if(!(blendif & 0x700u)) return t;
  const f32m::tables_t tb = f32m::global_tables();
  const float b = 1.15f, g = 0.66f, c1 = 0.8359375f, c2 = 18.8515625f, c3 = 18.6875f, n = 0.159301758f, p = 134.034375f, d = -0.56f, d0 = 1.6295499532821566e-11f;
  float d65[3];
#pragma unroll
  for(int c = 0; c < 3; c++) d65[c] = mo[3 * c + 2] * px[2] + (mo[3 * c + 1] * px[1] + mo[3 * c] * px[0]);
  const float xyz[3] = { b * d65[0] - (b - 1.0f) * d65[2], g * d65[1] - (g - 1.0f) * d65[0], d65[2] };
  const float M[3][3] = { { 0.41478972f, 0.579999f, 0.0146480f }, { -0.2015100f, 1.120649f, 0.0531008f }, { -0.0166008f, 0.264800f, 0.6684799f } };
  const float At[3][3] = { { 0.5f, 3.524000f, 0.199076f }, { 0.5f, -4.066708f, 1.096799f }, { 0.0f, 0.542708f, -1.295875f } };
  float lms[3], jab[3];
#pragma unroll
  for(int i = 0; i < 3; i++)
  {
    float v = M[i][0] * xyz[0] + M[i][1] * xyz[1] + M[i][2] * xyz[2];
    v = f32m::powf_(tb, fmaxf(bl_divc(v, 10000.f), 0.0f), n);
    lms[i] = f32m::powf_(tb, (c1 + c2 * v) / (1.0f + c3 * v), p);
  }
#pragma unroll
  for(int c = 0; c < 3; c++) jab[c] = At[0][c] * lms[0] + At[1][c] * lms[1] + At[2][c] * lms[2];
  jab[0] = fmaxf(((1.0f + d) * jab[0]) / (1.0f + d * jab[0]) - d0, 0.f);
  const float h = bl_divc(f32m::atan2f_(jab[2], jab[1]), 2.0f * 3.14159265358979324f);
  const float jch[3] = { jab[0], f32m::hypotf_(jab[1], jab[2]), h >= 0.0f ? h : 1.0f + h };
  float factor = 1.0f;
#pragma unroll
  for(int i = 0; i < 3; i++) factor *= bl_factor(jch[i], (blendif >> 16) & (0x100u << i), par + BLENDIF_ITEMS * (8 + i));
  return t * factor;
}
// ---- display-referred RGB, blendif_rgb_hsl.c.  HSL / HSV: common/colorspaces_inline_conversions.h _dt_RGB_2_Hue :420-435, _dt_Hue_2_RGB
// :438-484, dt_RGB_2_HSL :488-514, dt_HSL_2_RGB :517-528, dt_RGB_2_HSV :532-555, dt_HSV_2_RGB :558-564 ----
__device__ __forceinline__ float bl_clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); } // clamp_simd, math/openmp_maths.h:128-131
__device__ __forceinline__ float bl_rgb_to_hue(const float rgb[3], float max, float delta)
{
  float hue;
  if(rgb[0] == max)
    hue = (rgb[1] - rgb[2]) / delta;
  else if(rgb[1] == max)
    hue = 2.0f + (rgb[2] - rgb[0]) / delta;
  else
    hue = 4.0f + (rgb[0] - rgb[1]) / delta;
  hue = bl_divc(hue, 6.0f);
  if(hue < 0.0f) hue += 1.0f;
  if(hue > 1.0f) hue -= 1.0f;
  return hue;
}
__device__ __forceinline__ void bl_hue_to_rgb(float rgb[3], float H, float C, float min)
{
  const float h = H * 6.0f, i = floorf(h), f = h - i, fc = f * C, top = C + min, inc = fc + min, dec = top - fc;
  // the source switches on (size_t)i: sextants 0..4 by value, anything else (5, negative, huge, NaN) takes the last branch; compared as floats
  // here, because a conversion of a negative or NaN float to an unsigned integer saturates on the device and wraps on x86
  if(i == 0.0f)
  {
    rgb[0] = top;
    rgb[1] = inc;
    rgb[2] = min;
  }
  else if(i == 1.0f)
  {
    rgb[0] = dec;
    rgb[1] = top;
    rgb[2] = min;
  }
  else if(i == 2.0f)
  {
    rgb[0] = min;
    rgb[1] = top;
    rgb[2] = inc;
  }
  else if(i == 3.0f)
  {
    rgb[0] = min;
    rgb[1] = dec;
    rgb[2] = top;
  }
  else if(i == 4.0f)
  {
    rgb[0] = inc;
    rgb[1] = min;
    rgb[2] = top;
  }
  else
  {
    rgb[0] = top;
    rgb[1] = min;
    rgb[2] = dec;
  }
}
__device__ __forceinline__ void bl_rgb_to_hsl(const float rgb[3], float hsl[3])
{
  const float min = fminf(rgb[0], fminf(rgb[1], rgb[2])), max = fmaxf(rgb[0], fmaxf(rgb[1], rgb[2])), delta = max - min;
  const float L = (max + min) / 2.0f;
  float H = 0.0f, S = 0.0f;
  if(fabsf(max) > 1e-6f && fabsf(delta) > 1e-6f)
  {
    S = L < 0.5f ? delta / (max + min) : delta / (2.0f - max - min);
    H = bl_rgb_to_hue(rgb, max, delta);
  }
  hsl[0] = H;
  hsl[1] = S;
  hsl[2] = L;
}
__device__ __forceinline__ void bl_hsl_to_rgb(const float hsl[3], float rgb[3])
{
  const float L = hsl[2];
  const float C = L < 0.5f ? L * hsl[1] : (1.0f - L) * hsl[1];
  bl_hue_to_rgb(rgb, hsl[0], 2.0f * C, L - C);
}
__device__ __forceinline__ void bl_rgb_to_hsv(const float rgb[3], float hsv[3])
{
  const float min = fminf(rgb[0], fminf(rgb[1], rgb[2])), max = fmaxf(rgb[0], fmaxf(rgb[1], rgb[2])), delta = max - min;
  float H = 0.0f, S = 0.0f;
  if(fabsf(max) > 1e-6f && fabsf(delta) > 1e-6f)
  {
    S = delta / max;
    H = bl_rgb_to_hue(rgb, max, delta);
  }
  hsv[0] = H;
  hsv[1] = S;
  hsv[2] = max;
}
__device__ __forceinline__ void bl_hsv_to_rgb(const float hsv[3], float rgb[3])
{
  const float C = hsv[1] * hsv[2];
  bl_hue_to_rgb(rgb, hsv[0], C, hsv[2] - C);
}
// the H, S, L channels :149-163 and their call :206-215
__device__ __forceinline__ float bl_hsl(const float px[4], float t, unsigned blendif, const float *par)
{
  if(!(blendif & 0x700u)) return t;
  float hsl[3], factor = 1.0f;
  bl_rgb_to_hsl(px, hsl);
#pragma unroll
  for(int i = 0; i < 3; i++) factor *= bl_factor(hsl[i], (blendif >> 16) & (0x100u << i), par + BLENDIF_ITEMS * (8 + i));
  return t * factor;
}
// the operators :347-913
__device__ __forceinline__ void bl_operator_display(unsigned mode, const float a[4], const float b[4], float lo, float out[4])
{
  const float lo2 = lo * lo, na = 1.0f - lo, na2 = 1.0f - lo2;
  const unsigned m = mode & 0xFFu;
  switch(m)
  {
    case 0x02:
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = bl_clamp01(a[k] * na + fmaxf(a[k], b[k]) * lo);
      break;
    case 0x03:
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = bl_clamp01(a[k] * na + fminf(a[k], b[k]) * lo);
      break;
    case 0x04:
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = bl_clamp01(a[k] * na + (a[k] * b[k]) * lo);
      break;
    case 0x05:
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = bl_clamp01(a[k] * na + (a[k] + b[k]) / 2.0f * lo);
      break;
    case 0x06:
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = bl_clamp01(a[k] * na + (a[k] + b[k]) * lo);
      break;
    case 0x07:
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = bl_clamp01(a[k] * na + ((b[k] + a[k]) - 1.0f) * lo);
      break;
    case 0x08:
    case 0x17:
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = bl_clamp01(a[k] * na + fabsf(a[k] - b[k]) * lo);
      break;
    case 0x09: // screen
#pragma unroll
      for(int k = 0; k < 3; k++)
      {
        const float la = bl_clamp01(a[k]), lb = bl_clamp01(b[k]);
        out[k] = bl_clamp01(la * na + (1.0f - (1.0f - la) * (1.0f - lb)) * lo);
      }
      break;
    case 0x0A: // overlay
    case 0x0C: // hardlight: the same with the test on the upper layer
#pragma unroll
      for(int k = 0; k < 3; k++)
      {
        const float la = bl_clamp01(a[k]), lb = bl_clamp01(b[k]);
        out[k] = bl_clamp01(la * na2 + ((m == 0x0A ? la : lb) > 0.5f ? 1.0f - (1.0f - 2.0f * (la - 0.5f)) * (1.0f - lb) : 2.0f * la * lb) * lo2);
      }
      break;
    case 0x0B: // softlight
#pragma unroll
      for(int k = 0; k < 3; k++)
      {
        const float la = bl_clamp01(a[k]), lb = bl_clamp01(b[k]);
        out[k] = bl_clamp01(la * na2 + (lb > 0.5f ? 1.0f - (1.0f - la) * (1.0f - (lb - 0.5f)) : la * (lb + 0.5f)) * lo2);
      }
      break;
    case 0x0D: // vividlight
#pragma unroll
      for(int k = 0; k < 3; k++)
      {
        const float la = bl_clamp01(a[k]), lb = bl_clamp01(b[k]);
        out[k] = bl_clamp01(la * na2 + (lb > 0.5f ? (lb >= 1.0f ? 1.0f : la / (2.0f * (1.0f - lb))) : (lb <= 0.0f ? 0.0f : 1.0f - (1.0f - la) / (2.0f * lb))) * lo2);
      }
      break;
    case 0x0E: // linearlight
#pragma unroll
      for(int k = 0; k < 3; k++)
      {
        const float la = bl_clamp01(a[k]), lb = bl_clamp01(b[k]);
        out[k] = bl_clamp01(la * na2 + (la + 2.0f * lb - 1.0f) * lo2);
      }
      break;
    case 0x0F: // pinlight
#pragma unroll
      for(int k = 0; k < 3; k++)
      {
        const float la = bl_clamp01(a[k]), lb = bl_clamp01(b[k]);
        out[k] = bl_clamp01(la * na2 + (lb > 0.5f ? fmaxf(la, 2.0f * (lb - 0.5f)) : fminf(la, 2.0f * lb)) * lo2);
      }
      break;
    case 0x10: // lightness
    case 0x11: // chromaticity
    case 0x12: // hue
    case 0x13: // colour
    case 0x16: // colour adjustment: through HSL :645-808
    {
      float ta[3], tb[3], tta[3], ttb[3];
#pragma unroll
      for(int k = 0; k < 3; k++)
      {
        ta[k] = bl_clamp01(a[k]);
        tb[k] = bl_clamp01(b[k]);
      }
      bl_rgb_to_hsl(ta, tta);
      bl_rgb_to_hsl(tb, ttb);
      if(m == 0x10 || m == 0x11)
        ttb[0] = tta[0];
      else
      { // the hue along the shortest way round the circle
        const float d = fabsf(tta[0] - ttb[0]);
        const float sh = d > 0.5f ? -lo * (1.0f - d) / d : lo;
        ttb[0] = fmodf((tta[0] * (1.0f - sh)) + ttb[0] * sh + 1.0f, 1.0f);
      }
      ttb[1] = (m == 0x10 || m == 0x12) ? tta[1] : (tta[1] * (1.0f - lo)) + ttb[1] * lo;
      if(m == 0x10)
        ttb[2] = (tta[2] * (1.0f - lo)) + ttb[2] * lo;
      else if(m != 0x16)
        ttb[2] = tta[2];
      bl_hsl_to_rgb(ttb, out);
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = bl_clamp01(out[k]);
      break;
    }
    case 0x19: // normal, bounded
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = bl_clamp01(a[k] * na + b[k] * lo);
      break;
    case 0x1C: // HSV value
    {
      float ta[3], tb[3];
      bl_rgb_to_hsv(a, ta);
      bl_rgb_to_hsv(b, tb);
      tb[0] = ta[0];
      tb[1] = ta[1];
      tb[2] = ta[2] * (1.0f - lo) + tb[2] * lo;
      bl_hsv_to_rgb(tb, out);
      break;
    }
    case 0x1D: // HSV colour: hue and saturation blended as a vector
    {
      const float two_pi = 2.0f * 3.14159265358979324f;
      float ta[3], tb[3];
      bl_rgb_to_hsv(a, ta);
      bl_rgb_to_hsv(b, tb);
      const float xa = ta[1] * f32m::cosf_(two_pi * ta[0]), ya = ta[1] * f32m::sinf_(two_pi * ta[0]);
      const float xb = tb[1] * f32m::cosf_(two_pi * tb[0]), yb = tb[1] * f32m::sinf_(two_pi * tb[0]);
      const float xc = xa * (1.0f - lo) + xb * lo, yc = ya * (1.0f - lo) + yb * lo;
      tb[0] = bl_divc(f32m::atan2f_(yc, xc), two_pi);
      if(tb[0] < 0.0f) tb[0] += 1.0f;
      tb[1] = sqrtf(xc * xc + yc * yc);
      tb[2] = ta[2];
      bl_hsv_to_rgb(tb, out);
      break;
    }
    case 0x21:
    case 0x22:
    case 0x23:
    {
      const int c = (int)m - 0x21;
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = (k == c) ? a[k] * (1.0f - lo) + b[k] * lo : a[k];
      break;
    }
    default: // normal
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = a[k] * na + b[k] * lo;
      break;
  }
  out[3] = lo;
}

__device__ __forceinline__ float bl_mask(const blend_plan_t &pl, const float a[4], const float b[4], float form)
{
  if(pl.kind == 0) return pl.opacity;
  if(pl.kind == 1) return form * pl.opacity;
  float m = pl.seed_form ? form : pl.fill;
  const float g = pl.opacity;
  if(pl.pm == 0)
    m = pl.inversed ? g * (1.0f - m) : m * g; // :221-232
  else if(pl.pm == 1)
    m = pl.pm_const; // :233-240
  else
  { // :241-320
    float t;
    if(pl.lab)
    {
      t = bl_channels_lab(a, 1.0f, pl.blendif, pl.par, pl.c_scale);
      t = bl_channels_lab(b, t, pl.blendif >> 4, pl.par + BLENDIF_ITEMS * 4, pl.c_scale);
    }
    else
    {
      t = bl_channels(a, 1.0f, pl.blendif, pl.par, pl.lum);
      t = pl.display ? bl_hsl(a, t, pl.blendif, pl.par) : bl_jzczhz(a, t, pl.blendif, pl.par, pl.masking);
      t = bl_channels(b, t, pl.blendif >> 4, pl.par + BLENDIF_ITEMS * 4, pl.lum);
      t = pl.display ? bl_hsl(b, t, pl.blendif >> 4, pl.par + BLENDIF_ITEMS * 4) : bl_jzczhz(b, t, pl.blendif >> 4, pl.par + BLENDIF_ITEMS * 4, pl.masking);
    }
    if(pl.inclusive)
      m = pl.inversed ? g * (1.0f - m) * t : g * (1.0f - (1.0f - m) * t);
    else
      m = pl.inversed ? g * (1.0f - m * t) : g * m * t;
  }
  if(pl.tone)
  { // _develop_blend_process_mask_tone_curve(), :626-655
    const float mask_epsilon = 16 * FLT_EPSILON, e = pl.contrast_e, brightness = pl.brightness;
    float x = m / g;
    x = 2.f * x - 1.f;
    if(1.f - brightness <= 0.f)
      x = m <= mask_epsilon ? -1.f : 1.f;
    else if(1.f + brightness <= 0.f)
      x = m >= 1.f - mask_epsilon ? 1.f : -1.f;
    else if(brightness > 0.f)
    {
      x = (x + brightness) / (1.f - brightness);
      x = fminf(x, 1.f);
    }
    else
    {
      x = (x + brightness) / (1.f + brightness);
      x = fmaxf(x, -1.f);
    }
    const float v = ((x * e / (1.f + (e - 1.f) * fabsf(x))) / 2.f + 0.5f) * g;
    m = v > 1.f ? 1.f : (v < 0.f ? 0.f : v); // clamp_range_f, math/math.h:98
  }
  return m;
}
__device__ __forceinline__ float bl_sq(float x) { return x * x; }
// the operators, :328-585: a = the lower layer, b = the upper one, lo = the mask
__device__ __forceinline__ void bl_operator(unsigned mode, const float a[4], const float b[4], float p, float lo, float out[4])
{
  const float na = 1.0f - lo;
  switch(mode & 0xFFu)
  {
    case 0x04:
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = a[k] * na + (a[k] * b[k] * p) * lo;
      break;
    case 0x05:
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = a[k] * na + (a[k] + b[k]) / 2.0f * lo;
      break;
    case 0x06:
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = a[k] * na + (a[k] + p * b[k]) * lo;
      break;
    case 0x07:
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = a[k] * na + fmaxf(a[k] - p * b[k], 0.0f) * lo;
      break;
    case 0x25:
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = a[k] * na + fmaxf(b[k] - p * a[k], 0.0f) * lo;
      break;
    case 0x08:
    case 0x17:
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = a[k] * na + fabsf(a[k] - b[k]) * lo;
      break;
    case 0x26:
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = a[k] * na + a[k] / fmaxf(p * b[k], 1e-6f) * lo;
      break;
    case 0x27:
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = a[k] * na + b[k] / fmaxf(p * a[k], 1e-6f) * lo;
      break;
    case 0x28:
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = a[k] * na + sqrtf(fmaxf(a[k] * b[k], 0.0f)) * lo;
      break;
    case 0x29:
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = a[k] * na + 2.0f * a[k] * b[k] / (fmaxf(a[k], 5e-7f) + fmaxf(b[k], 5e-7f)) * lo;
      break;
    case 0x10:
    case 0x11:
    {
      const float norm_a = fmaxf(sqrtf(bl_sq(a[0]) + bl_sq(a[1]) + bl_sq(a[2])), 1e-6f), norm_b = fmaxf(sqrtf(bl_sq(b[0]) + bl_sq(b[1]) + bl_sq(b[2])), 1e-6f);
      if((mode & 0xFFu) == 0x11)
      {
#pragma unroll
        for(int k = 0; k < 3; k++) out[k] = a[k] * na + b[k] * norm_a / norm_b * lo;
      }
      else
      {
#pragma unroll
        for(int k = 0; k < 3; k++) out[k] = a[k] * na + a[k] * norm_b / norm_a * lo;
      }
      break;
    }
    case 0x21:
    case 0x22:
    case 0x23:
    {
      const int c = (int)(mode & 0xFFu) - 0x21;
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = (k == c) ? a[k] * na + p * b[k] * lo : a[k];
      break;
    }
    default:
#pragma unroll
      for(int k = 0; k < 3; k++) out[k] = a[k] * na + b[k] * lo;
      break;
  }
  out[3] = lo;
}

// ---- Lab, blendif_lab.c:302-1068: the pixels scaled to L / 100, a / 128, b / 128, blended between min = { 0, -1, -1 } and max = { 1, 1, 1 },
// scaled back.
__device__ __forceinline__ float bl_clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); } // _CLAMP :45-48
// a and b follow the lightness: what multiply, overlay, softlight, hardlight, vividlight and linearlight do to them
__device__ __forceinline__ void bl_lab_follow(const float ta[3], float tb[3], float o)
{
  const float f = fmaxf(ta[0], 0.01f);
  tb[1] = bl_clamp(ta[1] * (1.0f - o) + (ta[1] + tb[1]) * tb[0] / f * o, -1.0f, 1.0f);
  tb[2] = bl_clamp(ta[2] * (1.0f - o) + (ta[2] + tb[2]) * tb[0] / f * o, -1.0f, 1.0f);
}
__device__ __forceinline__ void bl_operator_lab(unsigned mode, const float a[4], const float b[4], float lo, float out[4])
{
  const float mn[3] = { 0.0f, -1.0f, -1.0f }, mx[3] = { 1.0f, 1.0f, 1.0f };
  float ta[3] = { a[0] * (1 / 100.0f), a[1] * (1 / 128.0f), a[2] * (1 / 128.0f) }, tb[3] = { b[0] * (1 / 100.0f), b[1] * (1 / 128.0f), b[2] * (1 / 128.0f) };
  const float lo2 = lo * lo;
  // the shifted lightness of the "light" family: lmax = max[0] + |min[0]| = 1, la and lb in [0, lmax]
  const float lmax = mx[0] + fabsf(mn[0]), halfmax = lmax / 2.0f, doublemax = lmax * 2.0f;
  const float la = bl_clamp(ta[0] + fabsf(mn[0]), 0.0f, lmax), lb = bl_clamp(tb[0] + fabsf(mn[0]), 0.0f, lmax);
  switch(mode & 0xFFu)
  {
    case 0x02: // lighten
    case 0x03: // darken
    {
      const float pick = (mode & 0xFFu) == 0x02 ? (ta[0] > tb[0] ? ta[0] : tb[0]) : (ta[0] < tb[0] ? ta[0] : tb[0]);
      tb[0] = bl_clamp(ta[0] * (1.0f - lo) + pick * lo, mn[0], mx[0]);
      tb[1] = bl_clamp(ta[1] * (1.0f - fabsf(tb[0] - ta[0])) + 0.5f * (ta[1] + tb[1]) * fabsf(tb[0] - ta[0]), mn[1], mx[1]);
      tb[2] = bl_clamp(ta[2] * (1.0f - fabsf(tb[0] - ta[0])) + 0.5f * (ta[2] + tb[2]) * fabsf(tb[0] - ta[0]), mn[2], mx[2]);
      break;
    }
    case 0x04: // multiply
      tb[0] = bl_clamp(ta[0] * (1.0f - lo) + (ta[0] * tb[0]) * lo, mn[0], mx[0]);
      bl_lab_follow(ta, tb, lo);
      break;
    case 0x05: // average
#pragma unroll
      for(int c = 0; c < 3; c++) tb[c] = bl_clamp(ta[c] * (1.0f - lo) + (ta[c] + tb[c]) / 2.0f * lo, mn[c], mx[c]);
      break;
    case 0x06: // add
#pragma unroll
      for(int c = 0; c < 3; c++) tb[c] = bl_clamp(ta[c] * (1.0f - lo) + (ta[c] + tb[c]) * lo, mn[c], mx[c]);
      break;
    case 0x07: // subtract
#pragma unroll
      for(int c = 0; c < 3; c++) tb[c] = bl_clamp(ta[c] * (1.0f - lo) + ((tb[c] + ta[c]) - (fabsf(mn[c] + mx[c]))) * lo, mn[c], mx[c]);
      break;
    case 0x08: // difference (deprecated)
#pragma unroll
      for(int c = 0; c < 3; c++)
      {
        const float cmax = mx[c] + fabsf(mn[c]);
        const float ca = bl_clamp(ta[c] + fabsf(mn[c]), 0.0f, cmax), cb = bl_clamp(tb[c] + fabsf(mn[c]), 0.0f, cmax);
        tb[c] = bl_clamp(ca * (1.0f - lo) + fabsf(ca - cb) * lo, 0.0f, cmax) - fabsf(mn[c]);
      }
      break;
    case 0x17: // difference
#pragma unroll
      for(int c = 0; c < 3; c++) tb[c] = fabsf(ta[c] - tb[c]) / fabsf(mx[c] - mn[c]);
      tb[0] = fmaxf(tb[0], fmaxf(tb[1], tb[2]));
      tb[0] = bl_clamp(ta[0] * (1.0f - lo) + tb[0] * lo, mn[0], mx[0]);
      tb[1] = 0.0f;
      tb[2] = 0.0f;
      break;
    case 0x09: // screen
    {
      tb[0] = bl_clamp(la * (1.0f - lo) + ((lmax - (lmax - la) * (lmax - lb))) * lo, 0.0f, lmax) - fabsf(mn[0]);
      const float f = fmaxf(ta[0], 0.01f);
      tb[1] = bl_clamp(ta[1] * (1.0f - lo) + 0.5f * (ta[1] + tb[1]) * tb[0] / f * lo, mn[1], mx[1]);
      tb[2] = bl_clamp(ta[2] * (1.0f - lo) + 0.5f * (ta[2] + tb[2]) * tb[0] / f * lo, mn[2], mx[2]);
      break;
    }
    case 0x0A: // overlay
      tb[0] = bl_clamp(la * (1.0f - lo2) + (la > halfmax ? lmax - (lmax - doublemax * (la - halfmax)) * (lmax - lb) : (doublemax * la) * lb) * lo2, 0.0f, lmax)
              - fabsf(mn[0]);
      bl_lab_follow(ta, tb, lo2);
      break;
    case 0x0B: // softlight
      tb[0] = bl_clamp(la * (1.0f - lo2) + (lb > halfmax ? lmax - (lmax - la) * (lmax - (lb - halfmax)) : la * (lb + halfmax)) * lo2, 0.0f, lmax) - fabsf(mn[0]);
      bl_lab_follow(ta, tb, lo2);
      break;
    case 0x0C: // hardlight
      tb[0] = bl_clamp(la * (1.0f - lo2) + (lb > halfmax ? lmax - (lmax - doublemax * (la - halfmax)) * (lmax - lb) : doublemax * la * lb) * lo2, 0.0f, lmax)
              - fabsf(mn[0]);
      bl_lab_follow(ta, tb, lo2);
      break;
    case 0x0D: // vividlight
      tb[0] = bl_clamp(la * (1.0f - lo2)
                           + (lb > halfmax ? (lb >= lmax ? lmax : la / (doublemax * (lmax - lb))) : (lb <= 0.0f ? 0.0f : lmax - (lmax - la) / (doublemax * lb))) * lo2,
                       0.0f, lmax)
              - fabsf(mn[0]);
      bl_lab_follow(ta, tb, lo2);
      break;
    case 0x0E: // linearlight
      tb[0] = bl_clamp(la * (1.0f - lo2) + (la + doublemax * lb - lmax) * lo2, 0.0f, lmax) - fabsf(mn[0]);
      bl_lab_follow(ta, tb, lo2);
      break;
    case 0x0F: // pinlight
      tb[0] = bl_clamp(la * (1.0f - lo2) + (lb > halfmax ? fmaxf(la, doublemax * (lb - halfmax)) : fminf(la, doublemax * lb)) * lo2, 0.0f, lmax) - fabsf(mn[0]);
      tb[1] = bl_clamp(ta[1], mn[1], mx[1]);
      tb[2] = bl_clamp(ta[2], mn[2], mx[2]);
      break;
    case 0x10: // lightness
      tb[0] = bl_clamp(ta[0] * (1.0f - lo) + tb[0] * lo, mn[0], mx[0]);
      tb[1] = bl_clamp(ta[1], mn[1], mx[1]);
      tb[2] = bl_clamp(ta[2], mn[2], mx[2]);
      break;
    case 0x11: // chromaticity
    case 0x12: // hue
    case 0x13: // colour
    case 0x16: // colour adjustment: the four operators through LCh :843-976
    {
      const unsigned m = mode & 0xFFu;
      float tta[3], ttb[3];
#pragma unroll
      for(int c = 0; c < 3; c++)
      {
        ta[c] = bl_clamp(ta[c], mn[c], mx[c]);
        tb[c] = bl_clamp(tb[c], mn[c], mx[c]);
      }
      bl_lab_to_lch(ta, tta);
      bl_lab_to_lch(tb, ttb);
      if(m != 0x16) ttb[0] = tta[0];
      ttb[1] = m == 0x12 ? tta[1] : (tta[1] * (1.0f - lo)) + ttb[1] * lo;
      if(m == 0x11)
        ttb[2] = tta[2];
      else
      { // the hue along the shortest way round the circle :888-891
        const float d = fabsf(tta[2] - ttb[2]);
        const float sh = d > 0.5f ? -lo * (1.0f - d) / d : lo;
        ttb[2] = fmodf((tta[2] * (1.0f - sh)) + ttb[2] * sh + 1.0f, 1.0f);
      }
      bl_lch_to_lab(ttb, tb);
#pragma unroll
      for(int c = 0; c < 3; c++) tb[c] = bl_clamp(tb[c], mn[c], mx[c]);
      break;
    }
    case 0x19: // normal, bounded
#pragma unroll
      for(int c = 0; c < 3; c++) tb[c] = bl_clamp(ta[c] * (1.0f - lo) + tb[c] * lo, mn[c], mx[c]);
      break;
    case 0x1A:
    case 0x1E: // Lab lightness
      tb[0] = ta[0] * (1.0f - lo) + tb[0] * lo;
      tb[1] = ta[1];
      tb[2] = ta[2];
      break;
    case 0x1F: // Lab a
      tb[0] = ta[0];
      tb[1] = ta[1] * (1.0f - lo) + tb[1] * lo;
      tb[2] = ta[2];
      break;
    case 0x20: // Lab b
      tb[0] = ta[0];
      tb[1] = ta[1];
      tb[2] = ta[2] * (1.0f - lo) + tb[2] * lo;
      break;
    case 0x1B: // Lab colour
      tb[0] = ta[0];
      tb[1] = ta[1] * (1.0f - lo) + tb[1] * lo;
      tb[2] = ta[2] * (1.0f - lo) + tb[2] * lo;
      break;
    default: // normal
#pragma unroll
      for(int c = 0; c < 3; c++) tb[c] = ta[c] * (1.0f - lo) + tb[c] * lo;
      break;
  }
  out[0] = tb[0] * 100.0f;
  out[1] = tb[1] * 128.0f;
  out[2] = tb[2] * 128.0f;
  out[3] = lo;
}

// ---- raw, blendif_raw.c:64-352: one sample per site; the operators of the display-referred space that work channel by channel, anything else
// is the unbounded normal blend ----
__device__ __forceinline__ float bl_operator_raw(unsigned mode, float a, float b, float lo)
{
  const float lo2 = lo * lo, na = 1.0f - lo, na2 = 1.0f - lo2;
  const float la = bl_clamp01(a), lb = bl_clamp01(b);
  switch(mode & 0xFFu)
  {
    case 0x02: return bl_clamp01(a * na + fmaxf(a, b) * lo);
    case 0x03: return bl_clamp01(a * na + fminf(a, b) * lo);
    case 0x04: return bl_clamp01(a * na + (a * b) * lo);
    case 0x05: return bl_clamp01(a * na + (a + b) / 2.0f * lo);
    case 0x06: return bl_clamp01(a * na + (a + b) * lo);
    case 0x07: return bl_clamp01(a * na + ((b + a) - 1.0f) * lo);
    case 0x08:
    case 0x17: return bl_clamp01(a * na + fabsf(a - b) * lo);
    case 0x09: return bl_clamp01(la * na + (1.0f - (1.0f - la) * (1.0f - lb)) * lo);
    case 0x0A: return bl_clamp01(la * na2 + (la > 0.5f ? 1.0f - (1.0f - 2.0f * (la - 0.5f)) * (1.0f - lb) : 2.0f * la * lb) * lo2);
    case 0x0B: return bl_clamp01(la * na2 + (lb > 0.5f ? 1.0f - (1.0f - la) * (1.0f - (lb - 0.5f)) : la * (lb + 0.5f)) * lo2);
    case 0x0C: return bl_clamp01(la * na2 + (lb > 0.5f ? 1.0f - (1.0f - 2.0f * (la - 0.5f)) * (1.0f - lb) : 2.0f * la * lb) * lo2);
    case 0x0D:
      return bl_clamp01(la * na2 + (lb > 0.5f ? (lb >= 1.0f ? 1.0f : la / (2.0f * (1.0f - lb))) : (lb <= 0.0f ? 0.0f : 1.0f - (1.0f - la) / (2.0f * lb))) * lo2);
    case 0x0E: return bl_clamp01(la * na2 + (la + 2.0f * lb - 1.0f) * lo2);
    case 0x0F: return bl_clamp01(la * na2 + (lb > 0.5f ? fmaxf(la, 2.0f * (lb - 0.5f)) : fminf(la, 2.0f * lb)) * lo2);
    case 0x19: return bl_clamp01(a * na + b * lo);
    default: return a * na + b * lo;
  }
}
// the raw space: 4 B of input, 4 B of output and 4 B of form mask in, 4 B out (+ 4 B of mask) per site
__global__ void __launch_bounds__(256) blend_raw_kernel(const __grid_constant__ blend_plan_t pl)
{
  const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
  if(x >= pl.ow) return;
  const size_t o = (size_t)y * pl.ow + x;
  const float *in = (const float *)pl.in;
  float *out = (float *)pl.out;
  const float a = __ldg(in + (size_t)(y + pl.yoffs) * pl.iw + pl.xoffs + x), b = out[o];
  const float none[4] = { 0.f, 0.f, 0.f, 0.f };
  const float m = bl_mask(pl, none, none, pl.form ? __ldg(pl.form + o) : 0.0f);
  out[o] = pl.reverse ? bl_operator_raw(pl.mode, b, a, m) : bl_operator_raw(pl.mode, a, b, m);
  if(pl.mask_out) pl.mask_out[o] = m;
}

__global__ void __launch_bounds__(256) blend_kernel(const __grid_constant__ blend_plan_t pl)
{
  const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
  if(x >= pl.ow) return;
  const size_t o = (size_t)y * pl.ow + x;
  const float4 av = __ldg(pl.in + (size_t)(y + pl.yoffs) * pl.iw + pl.xoffs + x), bv = pl.out[o];
  const float a[4] = { av.x, av.y, av.z, av.w }, b[4] = { bv.x, bv.y, bv.z, bv.w };
  const float m = bl_mask(pl, a, b, pl.form ? __ldg(pl.form + o) : 0.0f);
  float res[4];
  if(pl.lab)
    bl_operator_lab(pl.mode, pl.reverse ? b : a, pl.reverse ? a : b, m, res);
  else if(pl.display)
    bl_operator_display(pl.mode, pl.reverse ? b : a, pl.reverse ? a : b, m, res);
  else if(pl.reverse)
    bl_operator(pl.mode, b, a, pl.p, m, res);
  else
    bl_operator(pl.mode, a, b, pl.p, m, res);
  if(pl.keep_alpha) res[3] = a[3]; // :952-961: an earlier module's mask stays in the alpha lane
  pl.out[o] = make_float4(res[0], res[1], res[2], res[3]);
  if(pl.mask_out) pl.mask_out[o] = m;
}

// dt_develop_blendif_process_parameters(), blend.c:214-260; in Lab the limits of the a and b channels are offset by a half
void bl_parameters(float *par, const b200_blend_params_t *d)
{
  for(int i = 0; i < BLENDIF_SIZE; i++)
  {
    float *p = par + BLENDIF_ITEMS * i;
    const float *b = d->blendif_parameters + 4 * i;
    if(d->blendif & (1u << i))
    {
      const float boost = exp2f(d->blendif_boost_factors[i]);
      const float offset = (d->blend_cst == CS_LAB && (i == 1 || i == 2 || i == 5 || i == 6)) ? 0.5f : 0.0f;
      for(int k = 0; k < 4; k++) p[k] = (b[k] - offset) * boost;
      p[4] = 1.0f / fmaxf(0.001f, p[1] - p[0]);
      p[5] = 1.0f / fmaxf(0.001f, p[3] - p[2]);
      if(b[0] <= 0.0f && b[1] <= 0.0f) p[0] = p[1] = -INFINITY;
      if(b[2] >= 1.0f && b[3] >= 1.0f) p[2] = p[3] = INFINITY;
    }
    else
    {
      p[0] = p[1] = -INFINITY;
      p[2] = p[3] = INFINITY;
      p[4] = p[5] = 0.0f;
    }
  }
}

// which of the reference's branches a parameter block takes (blend.c:669-760, blendif_rgb_jzczhz.c:196-240); 1 = blending is off,
// < 0 = an error code
int bl_plan(blend_plan_t &pl, const b200_blend_params_t *d, bool have_form)
{
  memset(&pl, 0, sizeof(pl));
  if(!(d->mask_mode & MASK_ENABLED)) return 1; // :673
  const bool lab = d->blend_cst == CS_LAB, display = d->blend_cst == CS_RGB_DISPLAY, raw = d->blend_cst == CS_RAW;
  if(!lab && !display && !raw && d->blend_cst != CS_RGB_SCENE) return B200_ERR_UNSUPPORTED;
  if(!lab && !raw && d->profile_nonlinear) return B200_ERR_UNSUPPORTED;
  if(d->feathering_radius > 0.1f || d->blur_radius > 0.1f || d->details != 0.0f) return B200_ERR_UNSUPPORTED;
  { // dt_develop_blendif_init_masking_profile(), develop/blend.c:322-353: the profile's matrix_in taken to D65 by Bradford's matrix
    const float M[3][3] = { { 0.9555766f, -0.0230393f, 0.0631636f }, { -0.0282895f, 1.0099416f, 0.0210077f }, { 0.0122982f, -0.0204830f, 1.3299098f } };
    for(int y = 0; y < 3; y++)
      for(int x = 0; x < 3; x++)
      {
        float sum = 0.0f;
        for(int i = 0; i < 3; i++) sum += M[y][i] * d->matrix_in[3 * i + x];
        pl.masking[3 * y + x] = sum;
      }
  }
  pl.c_scale = 1.0f / (128.0f * sqrtf(2.0f)); // blendif_lab.c:125
  const unsigned channel_mask = lab ? (unsigned)BLENDIF_LAB_MASK : (unsigned)BLENDIF_RGB_MASK;
  pl.lab = lab;
  pl.display = display;
  pl.raw = raw;
  bool parametric = false; // dt_develop_blend_get_mask_usage(), :290-312
  if(d->mask_mode & MASK_PARAMETRIC)
    for(unsigned ch = 0; ch < BLENDIF_SIZE; ch++)
    {
      if(!(channel_mask & (1u << ch)) || !(d->blendif & (1u << ch))) continue;
      const float *c = d->blendif_parameters + 4 * ch;
      if(fabsf(c[0]) > 1e-6f || fabsf(c[1]) > 1e-6f || fabsf(c[2] - 1.0f) > 1e-6f || fabsf(c[3] - 1.0f) > 1e-6f) parametric = true;
    }
  const bool raster = d->raster_used && have_form, drawn = d->drawn_used && have_form;
  pl.opacity = fminf(fmaxf(d->opacity / 100.0f, 0.0f), 1.0f);
  if(!raster && !drawn && !parametric)
    pl.kind = 0;
  else if(raster && !drawn && !parametric)
    pl.kind = 1;
  else
  {
    pl.kind = 2;
    pl.seed_form = raster || drawn;
    pl.fill = (d->mask_combine & COMBINE_INCL) ? 0.0f : 1.0f;
    const unsigned any_active = d->blendif & channel_mask;
    pl.inclusive = (d->mask_combine & COMBINE_INCL) != 0;
    pl.inversed = (d->mask_combine & COMBINE_INV) != 0;
    pl.blendif = d->blendif ^ (pl.inclusive ? channel_mask << 16 : 0u);
    const unsigned canceling = (pl.blendif >> 16) & ~pl.blendif & channel_mask;
    if(raw || !(d->mask_mode & MASK_PARAMETRIC) || (!canceling && !any_active))
      pl.pm = 0; // the raw space has no channels: opacity and inversion only, blendif_raw.c:36-61
    else if(canceling || !any_active)
    {
      pl.pm = 1;
      pl.pm_const = ((pl.inversed == 0) ^ (pl.inclusive == 0)) ? pl.opacity : 0.0f;
    }
    else
    {
      pl.pm = 2;
      bl_parameters(pl.par, d);
    }
    pl.tone = (fabsf(d->contrast) >= 0.01f || fabsf(d->brightness) >= 0.01f) && pl.opacity > 1e-4f; // :432, :463
    pl.contrast_e = expf(3.f * d->contrast);
    pl.brightness = d->brightness;
  }
  for(int k = 0; k < 3; k++) pl.lum[k] = d->luminance[k];
  pl.p = exp2f(d->blend_parameter); // :913
  pl.mode = d->blend_mode;
  pl.reverse = (d->blend_mode & BLEND_REVERSE) == BLEND_REVERSE;
  pl.keep_alpha = (d->mask_display & B200_DISPLAY_MASK) != 0;
  return 0;
}
} // namespace

#ifndef B200_KERNELS_ON_CPU
using namespace b200;

static int blend_check(const b200_piece_t *piece, const b200_blend_params_t *bp, const void *in, void *out)
{
  if(!piece || !bp || !in || !out) return fail(B200_ERR_ARG, "blend: NULL argument");
  if(piece->roi_out.width <= 0 || piece->roi_out.height <= 0) return fail(B200_ERR_ARG, "blend: empty roi_out");
  // blend.c:690-708: roi_out has the scale of roi_in and lies inside it
  const int xoffs = piece->roi_out.x - piece->roi_in.x, yoffs = piece->roi_out.y - piece->roi_in.y;
  if(piece->roi_out.scale != piece->roi_in.scale || xoffs < 0 || yoffs < 0
     || ((xoffs > 0 || yoffs > 0) && (piece->roi_out.width + xoffs > piece->roi_in.width || piece->roi_out.height + yoffs > piece->roi_in.height)))
    return fail(B200_ERR_UNSUPPORTED, "blend: roi's do not match (the reference skips the blend here too)");
  return B200_OK;
}

extern "C" int b200_blend_process_dev(const b200_piece_t *piece, const b200_blend_params_t *bp, const void *d_in, void *d_out, const float *d_form_mask,
                                      float *d_mask, void *stream)
{
  int rc = blend_check(piece, bp, d_in, d_out);
  if(rc) return rc;
  if((rc = bind_device(piece->devid))) return rc;
  blend_plan_t pl;
  rc = bl_plan(pl, bp, d_form_mask != nullptr);
  if(rc == 1) return B200_OK; // blending is off: the module's output stays as it is
  if(rc) return fail(rc, "blend: not built for these parameters (colour space %d, feathering %.2f, blur %.2f, details %.2f, blendif 0x%x)", bp->blend_cst,
                     bp->feathering_radius, bp->blur_radius, bp->details, bp->blendif);
  pl.in = (const float4 *)d_in;
  pl.out = (float4 *)d_out;
  pl.form = d_form_mask;
  pl.mask_out = d_mask;
  pl.iw = piece->roi_in.width;
  pl.ow = piece->roi_out.width;
  pl.oh = piece->roi_out.height;
  pl.xoffs = piece->roi_out.x - piece->roi_in.x;
  pl.yoffs = piece->roi_out.y - piece->roi_in.y;
  const dim3 grid((unsigned)((pl.ow + 255) / 256), (unsigned)pl.oh);
  if(pl.raw)
    blend_raw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pl);
  else
    blend_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pl);
  B200_CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_blend_process_host(const b200_piece_t *piece, const b200_blend_params_t *bp, const void *in, void *out, const float *form_mask, float *mask)
{
  int rc = blend_check(piece, bp, in, out);
  if(rc) return rc;
  if((rc = bind_device(piece->devid))) return rc;
  const size_t bpp = bp->blend_cst == CS_RAW ? 4 : 16; // one float per site in the raw space
  const size_t ibytes = (size_t)piece->roi_in.width * piece->roi_in.height * bpp, opx = (size_t)piece->roi_out.width * piece->roi_out.height;
  void *d_in = nullptr, *d_out = nullptr, *d_form = nullptr, *d_mask = nullptr;
  cudaStream_t s;
  if((rc = host_stream(&s))) return rc;
  if((rc = scratch(SLOT_IN, ibytes, &d_in))) return rc;
  if((rc = scratch(SLOT_OUT, opx * bpp, &d_out))) return rc;
  if((rc = copy_h2d(d_in, in, ibytes, s))) return rc;
  if((rc = copy_h2d(d_out, out, opx * bpp, s))) return rc;
  if(form_mask)
  {
    if((rc = scratch(SLOT_TMP0, opx * 4, &d_form))) return rc;
    if((rc = copy_h2d(d_form, form_mask, opx * 4, s))) return rc;
  }
  if(mask && (rc = scratch(SLOT_TMP1, opx * 4, &d_mask))) return rc;
  if((rc = b200_blend_process_dev(piece, bp, d_in, d_out, (const float *)d_form, (float *)d_mask, (void *)s))) return rc;
  if((rc = copy_d2h(out, d_out, opx * bpp, s))) return rc;
  if(mask && (bp->mask_mode & MASK_ENABLED) && (rc = copy_d2h(mask, d_mask, opx * 4, s))) return rc;
  B200_CUDA_TRY(cudaStreamSynchronize(s));
  return B200_OK;
}

// tiling_callback_blendop(), develop/blend.c:1672-1691: the mask and the copy of the output the reference's passes need; the fused
// kernel needs neither, the factors are kept so that the host tiler cuts the same tiles
extern "C" void b200_blend_tiling(const b200_piece_t *piece, b200_tiling_t *tiling)
{
  (void)piece;
  if(!tiling) return;
  tiling->factor = 3.5f; // in + out + (guide, tmp) + two quarter buffers for the mask
  tiling->factor_cl = 3.5f;
  tiling->maxbuf = 1.0f;
  tiling->maxbuf_cl = 1.0f;
  tiling->overhead = 0;
  tiling->overlap = 0;
  tiling->xalign = 1;
  tiling->yalign = 1;
}
#endif
