/* CPU restatement of a module's blending (mask + blend operator) in the scene-referred RGB space and in Lab.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src: develop/blend.c dt_develop_blend_process :657-860 (mask usage :262-320, post operations :427-469,
 * dt_develop_blendif_process_parameters :214-260, _develop_blend_process_mask_tone_curve :626-655) and
 * develop/blends/blendif_rgb_jzczhz.c (_blendif_compute_factor :42-73, the gray / red / green / blue channels :75-121,
 * _blendif_combine_channels :151-194, dt_develop_blendif_rgb_jzczhz_make_mask :196-325, the sixteen operators :328-585,
 * _choose_blend_func :587-649, dt_develop_blendif_rgb_jzczhz_blend :878-961).  Pinned bit-for-bit against those lines cut verbatim
 * (oracle/_ref: ref_blend.c).  For blend_cst == DEVELOP_BLEND_CS_LAB: develop/blends/blendif_lab.c (the L / a / b / C / h channels :90-137,
 * _blendif_combine_channels :139-173, dt_develop_blendif_lab_make_mask :175-298, the 26 operators :302-1068, _choose_blend_func
 * :1070-1162, dt_develop_blendif_lab_blend :1302-1418), pinned against oracle/_ref: ref_blend_lab.c.  For
 * DEVELOP_BLEND_CS_RGB_DISPLAY: develop/blends/blendif_rgb_hsl.c (the gray / R / G / B / H / S / L channels :89-216, make_mask :218-345, the 27
 * operators :347-913, _choose_blend_func :915-1007, dt_develop_blendif_rgb_hsl_blend :1204-1288) with dt_RGB_2_HSL, dt_RGB_2_HSV and back
 * (common/colorspaces_inline_conversions.h:420-565), pinned against oracle/_ref: ref_blend_rgb_hsl.c.  For DEVELOP_BLEND_CS_RAW (buffers of
 * one float per site): develop/blends/blendif_raw.c (make_mask :36-61, the 16 operators :64-287, _choose_blend_func :290-352, the blend
 * :355-409), pinned against oracle/_ref: ref_blend_raw.c.
 *
 * Everything here is a function of one pixel of the module's input, the same pixel of its output and the same pixel of the form
 * mask (the raster / drawn mask the host rasterised): the reference's passes over whole buffers are folded into one evaluation per
 * pixel, which is also how the CUDA kernel does it.  Not restated (the entry point returns -1, the product B200_ERR_UNSUPPORTED):
 * feathering (guided filter), Gaussian blur and detail refinement of the mask, the GUI's channel display, the display-RGB and raw blend
 * colour spaces.
 */
#include "oracle_common.h"
#include "flt32_math.h"
#include <float.h>
#include <stdlib.h>
#include <string.h>

enum
{
  MASK_ENABLED = 1, MASK_SHAPE = 2, MASK_PARAMETRIC = 4, MASK_RASTER = 8, /* dt_develop_mask_mode_t, blend.h:110-118 */
  COMBINE_INV = 1, COMBINE_INCL = 2,                                      /* dt_develop_mask_combine_mode_t :120-131 */
  BLENDIF_SIZE = 16, BLENDIF_ITEMS = 6, BLENDIF_RGB_MASK = 0x77FF,        /* :188-191, :329 */
  BLENDIF_LAB_MASK = 0x3377, CS_RAW = 1, CS_LAB = 2, CS_RGB_DISPLAY = 3, CS_RGB_SCENE = 4, /* :52-59 */
  DISPLAY_MASK = 1                                                         /* develop.h:123 */
};
#define BLEND_REVERSE 0x80000000u /* blend.h:106 */

typedef struct orc_blend_params_t
{ /* the members of dt_develop_blend_params_t (blend.h:197-237) the path reads, then what the host looks up */
  uint32_t mask_mode;
  int32_t blend_cst;
  uint32_t blend_mode;
  float blend_parameter, opacity;
  uint32_t mask_combine, blendif;
  float feathering_radius;
  uint32_t feathering_guide;
  float blur_radius, contrast, brightness, details;
  float blendif_parameters[4 * BLENDIF_SIZE], blendif_boost_factors[BLENDIF_SIZE];
  int32_t raster_used, drawn_used; /* dt_develop_blend_get_mask_usage(): the form mask is a raster mask / a drawn mask (or both, combined) */
  float luminance[3];              /* row Y of the work profile's matrix_in */
  int32_t profile_nonlinear;
  uint32_t mask_display;           /* pipe->mask_display */
  float matrix_in[9];              /* the work profile's matrix_in (RGB -> XYZ D50), row by row: what the Jz / Cz / hz channels start from */
} orc_blend_params_t;

/* :214-260: in Lab the limits of the a and b channels are offset by a half */
static void blendif_parameters(float *par, const orc_blend_params_t *d)
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
/* :42-73 */
static float blendif_factor(float value, unsigned invert, const float *p)
{
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
/* :151-194 for one pixel: gray, red, green, blue of the channel set starting at bit 0 of `blendif` / at `par` */
static float blendif_channels(const float *px, float t, unsigned blendif, const float *par, const float *lum)
{
  if(blendif & 1u) t *= blendif_factor(lum[0] * px[0] + lum[1] * px[1] + lum[2] * px[2], (blendif >> 16) & 1u, par);
  for(int c = 0; c < 3; c++)
    if(blendif & (2u << c)) t *= blendif_factor(px[c], (blendif >> 16) & (2u << c), par + BLENDIF_ITEMS * (1 + c));
  return t;
}

#define PI_F 3.14159265358979324f /* DT_M_PI_F, math/math.h */
/* dt_Lab_2_LCH / dt_LCH_2_Lab, common/colorspaces_inline_conversions.h:594-615 */
static void lab_to_lch(const float *lab, float *lch)
{
  float h = atan2f(lab[2], lab[1]);
  if(h > 0.0f)
    h = h / (2.0f * PI_F);
  else
    h = 1.0f - fabsf(h) / (2.0f * PI_F);
  lch[0] = lab[0];
  lch[1] = hypotf(lab[1], lab[2]);
  lch[2] = h;
}
static void lch_to_lab(const float *lch, float *lab)
{
  lab[0] = lch[0];
  lab[1] = cosf(2.0f * PI_F * lch[2]) * lch[1];
  lab[2] = sinf(2.0f * PI_F * lch[2]) * lch[1];
}
/* blendif_lab.c:139-173 for one pixel: L, a, b, then chroma and hue together, of the channel set starting at bit 0 of `blendif` / at `par` */
static float blendif_channels_lab(const float *px, float t, unsigned blendif, const float *par)
{
  if(blendif & 1u) t *= blendif_factor(px[0] / 100.0f, (blendif >> 16) & 1u, par);
  if(blendif & 2u) t *= blendif_factor(px[1] / 256.0f, (blendif >> 16) & 2u, par + BLENDIF_ITEMS);
  if(blendif & 4u) t *= blendif_factor(px[2] / 256.0f, (blendif >> 16) & 4u, par + BLENDIF_ITEMS * 2);
  if(blendif & 0x300u)
  {
    const float c_scale = 1.0f / (128.0f * sqrtf(2.0f));
    float lch[3], factor = 1.0f;
    lab_to_lch(px, lch);
    factor *= blendif_factor(lch[1] * c_scale, (blendif >> 16) & 0x100u, par + BLENDIF_ITEMS * 8);
    factor *= blendif_factor(lch[2], (blendif >> 16) & 0x200u, par + BLENDIF_ITEMS * 9);
    t *= factor;
  }
  return t;
}

/* ---- the Jz, Cz, hz channels of the RGB space, blendif_rgb_jzczhz.c:122-149 ---- */
/* dt_develop_blendif_init_masking_profile(), develop/blend.c:322-353: the profile's matrix_in taken to D65 by Bradford's matrix */
static void masking_matrix(const float *matrix_in, float out[3][3])
{
  static const float M[3][3] = { { 0.9555766f, -0.0230393f, 0.0631636f }, { -0.0282895f, 1.0099416f, 0.0210077f }, { 0.0122982f, -0.0204830f, 1.3299098f } };
  for(int y = 0; y < 3; y++)
    for(int x = 0; x < 3; x++)
    {
      float sum = 0.0f;
      for(int i = 0; i < 3; i++) sum += M[y][i] * matrix_in[3 * i + x];
      out[y][x] = sum;
    }
}
/* dt_XYZ_2_JzAzBz + dt_JzAzBz_2_JzCzhz, common/colorspaces_inline_conversions.h:672-722, :775-781 */
static void xyz_to_jzczhz(const float *xyz_d65, float *jch)
{
  const float b = 1.15f, g = 0.66f, c1 = 0.8359375f, c2 = 18.8515625f, c3 = 18.6875f, n = 0.159301758f, p = 134.034375f, d = -0.56f, d0 = 1.6295499532821566e-11f;
  static const float M[3][3] = { { 0.41478972f, 0.579999f, 0.0146480f }, { -0.2015100f, 1.120649f, 0.0531008f }, { -0.0166008f, 0.264800f, 0.6684799f } };
  static const float At[3][3] = { { 0.5f, 3.524000f, 0.199076f }, { 0.5f, -4.066708f, 1.096799f }, { 0.0f, 0.542708f, -1.295875f } };
  float xyz[3], lms[3], jab[3];
  xyz[0] = b * xyz_d65[0] - (b - 1.0f) * xyz_d65[2];
  xyz[1] = g * xyz_d65[1] - (g - 1.0f) * xyz_d65[0];
  xyz[2] = xyz_d65[2];
  for(int i = 0; i < 3; i++)
  {
    lms[i] = M[i][0] * xyz[0] + M[i][1] * xyz[1] + M[i][2] * xyz[2];
    lms[i] = f32m_powf(fmaxf(lms[i] / 10000.f, 0.0f), n);
    lms[i] = f32m_powf((c1 + c2 * lms[i]) / (1.0f + c3 * lms[i]), p);
  }
  for(int c = 0; c < 3; c++) jab[c] = At[0][c] * lms[0] + At[1][c] * lms[1] + At[2][c] * lms[2];
  jab[0] = fmaxf(((1.0f + d) * jab[0]) / (1.0f + d * jab[0]) - d0, 0.f);
  const float h = f32m_atan2f(jab[2], jab[1]) / (2.0f * PI_F);
  jch[0] = jab[0];
  jch[1] = f32m_hypotf(jab[1], jab[2]);
  jch[2] = h >= 0.0f ? h : 1.0f + h;
}
static float blendif_jzczhz(const float *px, float t, unsigned blendif, const float *par, const float (*mo)[3])
{ /* :122-149 and its call :186-193: the three factors multiplied together, then into the mask */
  if(!(blendif & 0x700u)) return t;
  float xyz[3], jch[3];
  /* dt_mat3x4_mul_vec4, system/simd.h:189-197: row0 * in[0], then row1 * in[1] + that, then row2 * in[2] + that */
  for(int c = 0; c < 3; c++) xyz[c] = mo[c][2] * px[2] + (mo[c][1] * px[1] + mo[c][0] * px[0]);
  xyz_to_jzczhz(xyz, jch);
  float factor = 1.0f;
  for(int i = 0; i < 3; i++) factor *= blendif_factor(jch[i], (blendif >> 16) & (0x100u << i), par + BLENDIF_ITEMS * (8 + i));
  return t * factor;
}

typedef struct
{
  int kind;        /* 0: mask = opacity; 1: mask = form * opacity (raster only); 2: seed, then the parametric stage */
  int seed_form;   /* kind 2: the seed is the form mask, else `fill` */
  float fill, opacity;
  int pm;          /* parametric stage: 0 = opacity * m (or opacity * (1 - m) inverted); 1 = the constant `pm_const`; 2 = channels */
  int inversed, inclusive;
  float pm_const;
  unsigned blendif;
  float par[BLENDIF_ITEMS * BLENDIF_SIZE];
  float masking[3][3]; /* RGB space: matrix_out of the masking profile */
  int tone;        /* mask tone curve :626-655 */
  float contrast_e, brightness;
} blend_plan_t;

static float blendif_hsl(const float *px, float t, unsigned blendif, const float *par); /* the display-referred space, below */
static float plan_mask(const blend_plan_t *pl, const orc_blend_params_t *d, const float *a, const float *b, float form)
{
  if(pl->kind == 0) return pl->opacity;
  if(pl->kind == 1) return form * pl->opacity;
  float m = pl->seed_form ? form : pl->fill;
  const float g = pl->opacity;
  if(pl->pm == 0)
    m = pl->inversed ? g * (1.0f - m) : m * g; /* :221-232 */
  else if(pl->pm == 1)
    m = pl->pm_const; /* :233-240 */
  else
  { /* :241-320 */
    float t;
    if(d->blend_cst == CS_LAB)
    {
      t = blendif_channels_lab(a, 1.0f, pl->blendif, pl->par);
      t = blendif_channels_lab(b, t, pl->blendif >> 4, pl->par + BLENDIF_ITEMS * 4);
    }
    else
    {
      const int display = d->blend_cst == CS_RGB_DISPLAY;
      t = blendif_channels(a, 1.0f, pl->blendif, pl->par, d->luminance);
      t = display ? blendif_hsl(a, t, pl->blendif, pl->par) : blendif_jzczhz(a, t, pl->blendif, pl->par, pl->masking);
      t = blendif_channels(b, t, pl->blendif >> 4, pl->par + BLENDIF_ITEMS * 4, d->luminance);
      t = display ? blendif_hsl(b, t, pl->blendif >> 4, pl->par + BLENDIF_ITEMS * 4)
                  : blendif_jzczhz(b, t, pl->blendif >> 4, pl->par + BLENDIF_ITEMS * 4, pl->masking);
    }
    if(pl->inclusive)
      m = pl->inversed ? g * (1.0f - m) * t : g * (1.0f - (1.0f - m) * t);
    else
      m = pl->inversed ? g * (1.0f - m * t) : g * m * t;
  }
  if(pl->tone)
  { /* :626-655 */
    const float mask_epsilon = 16 * FLT_EPSILON, e = pl->contrast_e, brightness = pl->brightness;
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
    m = v < 0.f ? 0.f : (v > 1.f ? 1.f : v); /* clamp_range_f, math/math.h:98 (a NaN passes through both tests) */
  }
  return m;
}

static float sq(float x) { return x * x; }
/* :328-585: a = the lower layer, b = the upper one (swapped by the caller for DEVELOP_BLEND_REVERSE), lo = the mask */
static void blend_pixel(unsigned mode, const float *a, const float *b, float p, float lo, float *out)
{
  const float na = 1.0f - lo;
  switch(mode & 0xFFu)
  {
    case 0x04: for(int k = 0; k < 3; k++) out[k] = a[k] * na + (a[k] * b[k] * p) * lo; break;                      /* multiply */
    case 0x05: for(int k = 0; k < 3; k++) out[k] = a[k] * na + (a[k] + b[k]) / 2.0f * lo; break;                    /* average */
    case 0x06: for(int k = 0; k < 3; k++) out[k] = a[k] * na + (a[k] + p * b[k]) * lo; break;                       /* add */
    case 0x07: for(int k = 0; k < 3; k++) out[k] = a[k] * na + fmaxf(a[k] - p * b[k], 0.0f) * lo; break;            /* subtract */
    case 0x25: for(int k = 0; k < 3; k++) out[k] = a[k] * na + fmaxf(b[k] - p * a[k], 0.0f) * lo; break;            /* subtract inverse */
    case 0x08:
    case 0x17: for(int k = 0; k < 3; k++) out[k] = a[k] * na + fabsf(a[k] - b[k]) * lo; break;                      /* difference */
    case 0x26: for(int k = 0; k < 3; k++) out[k] = a[k] * na + a[k] / fmaxf(p * b[k], 1e-6f) * lo; break;           /* divide */
    case 0x27: for(int k = 0; k < 3; k++) out[k] = a[k] * na + b[k] / fmaxf(p * a[k], 1e-6f) * lo; break;           /* divide inverse */
    case 0x28: for(int k = 0; k < 3; k++) out[k] = a[k] * na + sqrtf(fmaxf(a[k] * b[k], 0.0f)) * lo; break;         /* geometric mean */
    case 0x29:                                                                                                        /* harmonic mean */
      for(int k = 0; k < 3; k++) out[k] = a[k] * na + 2.0f * a[k] * b[k] / (fmaxf(a[k], 5e-7f) + fmaxf(b[k], 5e-7f)) * lo;
      break;
    case 0x10:
    case 0x11:
    { /* luminance (0x10, DEVELOP_BLEND_LIGHTNESS) and chromaticity (0x11) */
      const float norm_a = fmaxf(sqrtf(sq(a[0]) + sq(a[1]) + sq(a[2])), 1e-6f), norm_b = fmaxf(sqrtf(sq(b[0]) + sq(b[1]) + sq(b[2])), 1e-6f);
      for(int k = 0; k < 3; k++)
        out[k] = (mode & 0xFFu) == 0x11 ? a[k] * na + b[k] * norm_a / norm_b * lo : a[k] * na + a[k] * norm_b / norm_a * lo;
      break;
    }
    case 0x21:
    case 0x22:
    case 0x23:
    { /* one channel of RGB */
      const int c = (int)(mode & 0xFFu) - 0x21;
      for(int k = 0; k < 3; k++) out[k] = a[k];
      out[c] = a[c] * na + p * b[c] * lo;
      break;
    }
    default: for(int k = 0; k < 3; k++) out[k] = a[k] * na + b[k] * lo; break;                                     /* normal */
  }
  out[3] = lo;
}

/* ---- Lab, blendif_lab.c:302-1068: a = the lower layer, b = the upper one, lo = the mask; the pixels are scaled to L / 100, a / 128, b / 128,
 * blended between min = { 0, -1, -1 } and max = { 1, 1, 1 } and scaled back ---- */
static float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); } /* _CLAMP :45-48 */
static const float LAB_MIN[3] = { 0.0f, -1.0f, -1.0f }, LAB_MAX[3] = { 1.0f, 1.0f, 1.0f };
/* the shifted lightness of the operators of the "light" family: la, lb in [0, lmax] */
static void light_pair(const float *ta, const float *tb, float *la, float *lb, float *lmax)
{
  *lmax = LAB_MAX[0] + fabsf(LAB_MIN[0]);
  *la = clampf(ta[0] + fabsf(LAB_MIN[0]), 0.0f, *lmax);
  *lb = clampf(tb[0] + fabsf(LAB_MIN[0]), 0.0f, *lmax);
}
/* a and b follow the lightness: what multiply, overlay, softlight, hardlight, vividlight and linearlight do to them */
static void follow(const float *ta, float *tb, float o)
{
  const float f = fmaxf(ta[0], 0.01f);
  tb[1] = clampf(ta[1] * (1.0f - o) + (ta[1] + tb[1]) * tb[0] / f * o, LAB_MIN[1], LAB_MAX[1]);
  tb[2] = clampf(ta[2] * (1.0f - o) + (ta[2] + tb[2]) * tb[0] / f * o, LAB_MIN[2], LAB_MAX[2]);
}
static void hue_towards(const float *tta, float *ttb, float lo)
{ /* blend hue along the shortest distance on the colour circle :888-891 */
  const float d = fabsf(tta[2] - ttb[2]);
  const float s = d > 0.5f ? -lo * (1.0f - d) / d : lo;
  ttb[2] = fmodf((tta[2] * (1.0f - s)) + ttb[2] * s + 1.0f, 1.0f);
}
static void lab_blend_pixel(unsigned mode, const float *a, const float *b, float lo, float *out)
{
  static const float scale[3] = { 1 / 100.0f, 1 / 128.0f, 1 / 128.0f }, rescale[3] = { 100.0f, 128.0f, 128.0f };
  const float *mn = LAB_MIN, *mx = LAB_MAX;
  float ta[3], tb[3];
  for(int c = 0; c < 3; c++)
  {
    ta[c] = a[c] * scale[c];
    tb[c] = b[c] * scale[c];
  }
  const float lo2 = lo * lo;
  float la, lb, lmax;
  switch(mode & 0xFFu)
  {
    case 0x02: /* lighten */
    case 0x03: /* darken */
    {
      const float pick = (mode & 0xFFu) == 0x02 ? (ta[0] > tb[0] ? ta[0] : tb[0]) : (ta[0] < tb[0] ? ta[0] : tb[0]);
      tb[0] = clampf(ta[0] * (1.0f - lo) + pick * lo, mn[0], mx[0]);
      tb[1] = clampf(ta[1] * (1.0f - fabsf(tb[0] - ta[0])) + 0.5f * (ta[1] + tb[1]) * fabsf(tb[0] - ta[0]), mn[1], mx[1]);
      tb[2] = clampf(ta[2] * (1.0f - fabsf(tb[0] - ta[0])) + 0.5f * (ta[2] + tb[2]) * fabsf(tb[0] - ta[0]), mn[2], mx[2]);
      break;
    }
    case 0x04: /* multiply */
      tb[0] = clampf(ta[0] * (1.0f - lo) + (ta[0] * tb[0]) * lo, mn[0], mx[0]);
      follow(ta, tb, lo);
      break;
    case 0x05: /* average */
      for(int c = 0; c < 3; c++) tb[c] = clampf(ta[c] * (1.0f - lo) + (ta[c] + tb[c]) / 2.0f * lo, mn[c], mx[c]);
      break;
    case 0x06: /* add */
      for(int c = 0; c < 3; c++) tb[c] = clampf(ta[c] * (1.0f - lo) + (ta[c] + tb[c]) * lo, mn[c], mx[c]);
      break;
    case 0x07: /* subtract */
      for(int c = 0; c < 3; c++) tb[c] = clampf(ta[c] * (1.0f - lo) + ((tb[c] + ta[c]) - (fabsf(mn[c] + mx[c]))) * lo, mn[c], mx[c]);
      break;
    case 0x08: /* difference (deprecated) */
      for(int c = 0; c < 3; c++)
      {
        const float cmax = mx[c] + fabsf(mn[c]);
        const float ca = clampf(ta[c] + fabsf(mn[c]), 0.0f, cmax), cb = clampf(tb[c] + fabsf(mn[c]), 0.0f, cmax);
        tb[c] = clampf(ca * (1.0f - lo) + fabsf(ca - cb) * lo, 0.0f, cmax) - fabsf(mn[c]);
      }
      break;
    case 0x17: /* difference */
      for(int c = 0; c < 3; c++) tb[c] = fabsf(ta[c] - tb[c]) / fabsf(mx[c] - mn[c]);
      tb[0] = fmaxf(tb[0], fmaxf(tb[1], tb[2]));
      tb[0] = clampf(ta[0] * (1.0f - lo) + tb[0] * lo, mn[0], mx[0]);
      tb[1] = 0.0f;
      tb[2] = 0.0f;
      break;
    case 0x09: /* screen */
    {
      light_pair(ta, tb, &la, &lb, &lmax);
      tb[0] = clampf(la * (1.0f - lo) + ((lmax - (lmax - la) * (lmax - lb))) * lo, 0.0f, lmax) - fabsf(mn[0]);
      const float f = fmaxf(ta[0], 0.01f);
      tb[1] = clampf(ta[1] * (1.0f - lo) + 0.5f * (ta[1] + tb[1]) * tb[0] / f * lo, mn[1], mx[1]);
      tb[2] = clampf(ta[2] * (1.0f - lo) + 0.5f * (ta[2] + tb[2]) * tb[0] / f * lo, mn[2], mx[2]);
      break;
    }
    case 0x0A: /* overlay */
    {
      light_pair(ta, tb, &la, &lb, &lmax);
      const float halfmax = lmax / 2.0f, doublemax = lmax * 2.0f;
      tb[0] = clampf(la * (1.0f - lo2) + (la > halfmax ? lmax - (lmax - doublemax * (la - halfmax)) * (lmax - lb) : (doublemax * la) * lb) * lo2, 0.0f, lmax)
              - fabsf(mn[0]);
      follow(ta, tb, lo2);
      break;
    }
    case 0x0B: /* softlight */
    {
      light_pair(ta, tb, &la, &lb, &lmax);
      const float halfmax = lmax / 2.0f;
      tb[0] = clampf(la * (1.0f - lo2) + (lb > halfmax ? lmax - (lmax - la) * (lmax - (lb - halfmax)) : la * (lb + halfmax)) * lo2, 0.0f, lmax) - fabsf(mn[0]);
      follow(ta, tb, lo2);
      break;
    }
    case 0x0C: /* hardlight */
    {
      light_pair(ta, tb, &la, &lb, &lmax);
      const float halfmax = lmax / 2.0f, doublemax = lmax * 2.0f;
      tb[0] = clampf(la * (1.0f - lo2) + (lb > halfmax ? lmax - (lmax - doublemax * (la - halfmax)) * (lmax - lb) : doublemax * la * lb) * lo2, 0.0f, lmax)
              - fabsf(mn[0]);
      follow(ta, tb, lo2);
      break;
    }
    case 0x0D: /* vividlight */
    {
      light_pair(ta, tb, &la, &lb, &lmax);
      const float halfmax = lmax / 2.0f, doublemax = lmax * 2.0f;
      tb[0] = clampf(la * (1.0f - lo2)
                         + (lb > halfmax ? (lb >= lmax ? lmax : la / (doublemax * (lmax - lb))) : (lb <= 0.0f ? 0.0f : lmax - (lmax - la) / (doublemax * lb))) * lo2,
                     0.0f, lmax)
              - fabsf(mn[0]);
      follow(ta, tb, lo2);
      break;
    }
    case 0x0E: /* linearlight */
    {
      light_pair(ta, tb, &la, &lb, &lmax);
      const float doublemax = lmax * 2.0f;
      tb[0] = clampf(la * (1.0f - lo2) + (la + doublemax * lb - lmax) * lo2, 0.0f, lmax) - fabsf(mn[0]);
      follow(ta, tb, lo2);
      break;
    }
    case 0x0F: /* pinlight */
    {
      light_pair(ta, tb, &la, &lb, &lmax);
      const float halfmax = lmax / 2.0f, doublemax = lmax * 2.0f;
      tb[0] = clampf(la * (1.0f - lo2) + (lb > halfmax ? fmaxf(la, doublemax * (lb - halfmax)) : fminf(la, doublemax * lb)) * lo2, 0.0f, lmax) - fabsf(mn[0]);
      tb[1] = clampf(ta[1], mn[1], mx[1]);
      tb[2] = clampf(ta[2], mn[2], mx[2]);
      break;
    }
    case 0x10: /* lightness */
      tb[0] = clampf(ta[0] * (1.0f - lo) + tb[0] * lo, mn[0], mx[0]);
      tb[1] = clampf(ta[1], mn[1], mx[1]);
      tb[2] = clampf(ta[2], mn[2], mx[2]);
      break;
    case 0x11: /* chromaticity */
    case 0x12: /* hue */
    case 0x13: /* colour */
    case 0x16: /* colour adjustment */
    {
      const unsigned m = mode & 0xFFu;
      float tta[3], ttb[3];
      for(int c = 0; c < 3; c++)
      {
        ta[c] = clampf(ta[c], mn[c], mx[c]);
        tb[c] = clampf(tb[c], mn[c], mx[c]);
      }
      lab_to_lch(ta, tta);
      lab_to_lch(tb, ttb);
      if(m != 0x16) ttb[0] = tta[0];
      if(m == 0x12)
        ttb[1] = tta[1];
      else
        ttb[1] = (tta[1] * (1.0f - lo)) + ttb[1] * lo;
      if(m == 0x11)
        ttb[2] = tta[2];
      else
        hue_towards(tta, ttb, lo);
      lch_to_lab(ttb, tb);
      for(int c = 0; c < 3; c++) tb[c] = clampf(tb[c], mn[c], mx[c]);
      break;
    }
    case 0x19: /* normal, bounded */
      for(int c = 0; c < 3; c++) tb[c] = clampf(ta[c] * (1.0f - lo) + tb[c] * lo, mn[c], mx[c]);
      break;
    case 0x1A:
    case 0x1E: /* Lab lightness */
      tb[0] = ta[0] * (1.0f - lo) + tb[0] * lo;
      tb[1] = ta[1];
      tb[2] = ta[2];
      break;
    case 0x1F: /* Lab a */
      tb[0] = ta[0];
      tb[1] = ta[1] * (1.0f - lo) + tb[1] * lo;
      tb[2] = ta[2];
      break;
    case 0x20: /* Lab b */
      tb[0] = ta[0];
      tb[1] = ta[1];
      tb[2] = ta[2] * (1.0f - lo) + tb[2] * lo;
      break;
    case 0x1B: /* Lab colour */
      tb[0] = ta[0];
      tb[1] = ta[1] * (1.0f - lo) + tb[1] * lo;
      tb[2] = ta[2] * (1.0f - lo) + tb[2] * lo;
      break;
    default: /* normal */
      for(int c = 0; c < 3; c++) tb[c] = ta[c] * (1.0f - lo) + tb[c] * lo;
      break;
  }
  for(int c = 0; c < 3; c++) out[c] = tb[c] * rescale[c];
  out[3] = lo;
}

/* ---- display-referred RGB, blendif_rgb_hsl.c ---- */
static float clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); } /* clamp_simd, math/openmp_maths.h:128-131 */
/* common/colorspaces_inline_conversions.h: _dt_RGB_2_Hue :420-435, _dt_Hue_2_RGB :438-484, dt_RGB_2_HSL :488-514, dt_HSL_2_RGB :517-528,
 * dt_RGB_2_HSV :532-555, dt_HSV_2_RGB :558-564 */
static float rgb_to_hue(const float *rgb, float max, float delta)
{
  float hue;
  if(rgb[0] == max)
    hue = (rgb[1] - rgb[2]) / delta;
  else if(rgb[1] == max)
    hue = 2.0f + (rgb[2] - rgb[0]) / delta;
  else
    hue = 4.0f + (rgb[0] - rgb[1]) / delta;
  hue /= 6.0f;
  if(hue < 0.0f) hue += 1.0f;
  if(hue > 1.0f) hue -= 1.0f;
  return hue;
}
static void hue_to_rgb(float *rgb, float H, float C, float min)
{
  const float h = H * 6.0f, i = floorf(h), f = h - i, fc = f * C, top = C + min, inc = fc + min, dec = top - fc;
  const size_t i_idx = (size_t)i;
  const float r[6][3] = { { top, inc, min }, { dec, top, min }, { min, top, inc }, { min, dec, top }, { inc, min, top }, { top, min, dec } };
  const int k = i_idx < 5 ? (int)i_idx : 5;
  for(int c = 0; c < 3; c++) rgb[c] = r[k][c];
}
static void rgb_to_hsl(const float *rgb, float *hsl)
{
  const float min = fminf(rgb[0], fminf(rgb[1], rgb[2])), max = fmaxf(rgb[0], fmaxf(rgb[1], rgb[2])), delta = max - min;
  const float L = (max + min) / 2.0f;
  float H = 0.0f, S = 0.0f;
  if(fabsf(max) > 1e-6f && fabsf(delta) > 1e-6f)
  {
    S = L < 0.5f ? delta / (max + min) : delta / (2.0f - max - min);
    H = rgb_to_hue(rgb, max, delta);
  }
  hsl[0] = H;
  hsl[1] = S;
  hsl[2] = L;
}
static void hsl_to_rgb(const float *hsl, float *rgb)
{
  const float L = hsl[2];
  const float C = L < 0.5f ? L * hsl[1] : (1.0f - L) * hsl[1];
  hue_to_rgb(rgb, hsl[0], 2.0f * C, L - C);
}
static void rgb_to_hsv(const float *rgb, float *hsv)
{
  const float min = fminf(rgb[0], fminf(rgb[1], rgb[2])), max = fmaxf(rgb[0], fmaxf(rgb[1], rgb[2])), delta = max - min;
  float H = 0.0f, S = 0.0f;
  if(fabsf(max) > 1e-6f && fabsf(delta) > 1e-6f)
  {
    S = delta / max;
    H = rgb_to_hue(rgb, max, delta);
  }
  hsv[0] = H;
  hsv[1] = S;
  hsv[2] = max;
}
static void hsv_to_rgb(const float *hsv, float *rgb)
{
  const float C = hsv[1] * hsv[2];
  hue_to_rgb(rgb, hsv[0], C, hsv[2] - C);
}
/* the H, S, L channels :149-163 and their call :206-215 */
static float blendif_hsl(const float *px, float t, unsigned blendif, const float *par)
{
  if(!(blendif & 0x700u)) return t;
  float hsl[3], factor = 1.0f;
  rgb_to_hsl(px, hsl);
  for(int i = 0; i < 3; i++) factor *= blendif_factor(hsl[i], (blendif >> 16) & (0x100u << i), par + BLENDIF_ITEMS * (8 + i));
  return t * factor;
}
/* the operators :347-913: a = the lower layer, b = the upper one, lo = the mask */
static void hsl_blend_pixel(unsigned mode, const float *a, const float *b, float lo, float *out)
{
  const float lo2 = lo * lo, na = 1.0f - lo, na2 = 1.0f - lo2;
  const unsigned m = mode & 0xFFu;
  switch(m)
  {
    case 0x02: for(int k = 0; k < 3; k++) out[k] = clamp01(a[k] * na + fmaxf(a[k], b[k]) * lo); break;        /* lighten */
    case 0x03: for(int k = 0; k < 3; k++) out[k] = clamp01(a[k] * na + fminf(a[k], b[k]) * lo); break;        /* darken */
    case 0x04: for(int k = 0; k < 3; k++) out[k] = clamp01(a[k] * na + (a[k] * b[k]) * lo); break;            /* multiply */
    case 0x05: for(int k = 0; k < 3; k++) out[k] = clamp01(a[k] * na + (a[k] + b[k]) / 2.0f * lo); break;     /* average */
    case 0x06: for(int k = 0; k < 3; k++) out[k] = clamp01(a[k] * na + (a[k] + b[k]) * lo); break;            /* add */
    case 0x07: for(int k = 0; k < 3; k++) out[k] = clamp01(a[k] * na + ((b[k] + a[k]) - 1.0f) * lo); break;   /* subtract */
    case 0x08:
    case 0x17: for(int k = 0; k < 3; k++) out[k] = clamp01(a[k] * na + fabsf(a[k] - b[k]) * lo); break;       /* difference */
    case 0x09: /* screen */
      for(int k = 0; k < 3; k++)
      {
        const float la = clamp01(a[k]), lb = clamp01(b[k]);
        out[k] = clamp01(la * na + (1.0f - (1.0f - la) * (1.0f - lb)) * lo);
      }
      break;
    case 0x0A: /* overlay */
    case 0x0C: /* hardlight: the same with the test on the upper layer */
      for(int k = 0; k < 3; k++)
      {
        const float la = clamp01(a[k]), lb = clamp01(b[k]);
        out[k] = clamp01(la * na2 + ((m == 0x0A ? la : lb) > 0.5f ? 1.0f - (1.0f - 2.0f * (la - 0.5f)) * (1.0f - lb) : 2.0f * la * lb) * lo2);
      }
      break;
    case 0x0B: /* softlight */
      for(int k = 0; k < 3; k++)
      {
        const float la = clamp01(a[k]), lb = clamp01(b[k]);
        out[k] = clamp01(la * na2 + (lb > 0.5f ? 1.0f - (1.0f - la) * (1.0f - (lb - 0.5f)) : la * (lb + 0.5f)) * lo2);
      }
      break;
    case 0x0D: /* vividlight */
      for(int k = 0; k < 3; k++)
      {
        const float la = clamp01(a[k]), lb = clamp01(b[k]);
        out[k] = clamp01(la * na2 + (lb > 0.5f ? (lb >= 1.0f ? 1.0f : la / (2.0f * (1.0f - lb))) : (lb <= 0.0f ? 0.0f : 1.0f - (1.0f - la) / (2.0f * lb))) * lo2);
      }
      break;
    case 0x0E: /* linearlight */
      for(int k = 0; k < 3; k++)
      {
        const float la = clamp01(a[k]), lb = clamp01(b[k]);
        out[k] = clamp01(la * na2 + (la + 2.0f * lb - 1.0f) * lo2);
      }
      break;
    case 0x0F: /* pinlight */
      for(int k = 0; k < 3; k++)
      {
        const float la = clamp01(a[k]), lb = clamp01(b[k]);
        out[k] = clamp01(la * na2 + (lb > 0.5f ? fmaxf(la, 2.0f * (lb - 0.5f)) : fminf(la, 2.0f * lb)) * lo2);
      }
      break;
    case 0x10: /* lightness */
    case 0x11: /* chromaticity */
    case 0x12: /* hue */
    case 0x13: /* colour */
    case 0x16: /* colour adjustment: through HSL :645-808 */
    {
      float ta[3], tb[3], tta[3], ttb[3];
      for(int k = 0; k < 3; k++)
      {
        ta[k] = clamp01(a[k]);
        tb[k] = clamp01(b[k]);
      }
      rgb_to_hsl(ta, tta);
      rgb_to_hsl(tb, ttb);
      if(m == 0x10 || m == 0x11)
        ttb[0] = tta[0];
      else
      { /* the hue along the shortest way round the circle */
        const float d = fabsf(tta[0] - ttb[0]);
        const float s = d > 0.5f ? -lo * (1.0f - d) / d : lo;
        ttb[0] = fmodf((tta[0] * (1.0f - s)) + ttb[0] * s + 1.0f, 1.0f);
      }
      if(m == 0x10 || m == 0x12)
        ttb[1] = tta[1];
      else
        ttb[1] = (tta[1] * (1.0f - lo)) + ttb[1] * lo;
      if(m == 0x10)
        ttb[2] = (tta[2] * (1.0f - lo)) + ttb[2] * lo;
      else if(m != 0x16)
        ttb[2] = tta[2];
      hsl_to_rgb(ttb, out);
      for(int k = 0; k < 3; k++) out[k] = clamp01(out[k]);
      break;
    }
    case 0x19: for(int k = 0; k < 3; k++) out[k] = clamp01(a[k] * na + b[k] * lo); break; /* normal, bounded */
    case 0x1C: /* HSV value */
    {
      float ta[3], tb[3];
      rgb_to_hsv(a, ta);
      rgb_to_hsv(b, tb);
      tb[0] = ta[0];
      tb[1] = ta[1];
      tb[2] = ta[2] * (1.0f - lo) + tb[2] * lo;
      hsv_to_rgb(tb, out);
      break;
    }
    case 0x1D: /* HSV colour */
    {
      float ta[3], tb[3];
      rgb_to_hsv(a, ta);
      rgb_to_hsv(b, tb);
      const float xa = ta[1] * f32m_cosf(2.0f * PI_F * ta[0]), ya = ta[1] * f32m_sinf(2.0f * PI_F * ta[0]);
      const float xb = tb[1] * f32m_cosf(2.0f * PI_F * tb[0]), yb = tb[1] * f32m_sinf(2.0f * PI_F * tb[0]);
      const float xc = xa * (1.0f - lo) + xb * lo, yc = ya * (1.0f - lo) + yb * lo;
      tb[0] = f32m_atan2f(yc, xc) / (2.0f * PI_F);
      if(tb[0] < 0.0f) tb[0] += 1.0f;
      tb[1] = sqrtf(xc * xc + yc * yc);
      tb[2] = ta[2];
      hsv_to_rgb(tb, out);
      break;
    }
    case 0x21:
    case 0x22:
    case 0x23:
      for(int k = 0; k < 3; k++) out[k] = a[k];
      out[m - 0x21] = a[m - 0x21] * (1.0f - lo) + b[m - 0x21] * lo;
      break;
    default: for(int k = 0; k < 3; k++) out[k] = a[k] * na + b[k] * lo; break; /* normal */
  }
  out[3] = lo;
}

/* ---- raw, blendif_raw.c:64-352: one sample per site; the operators of the display-referred space that work channel by channel, anything
 * else is the unbounded normal blend ---- */
static float raw_blend_value(unsigned mode, float a, float b, float lo)
{
  const float lo2 = lo * lo, na = 1.0f - lo, na2 = 1.0f - lo2;
  const float la = clamp01(a), lb = clamp01(b);
  switch(mode & 0xFFu)
  {
    case 0x02: return clamp01(a * na + fmaxf(a, b) * lo);
    case 0x03: return clamp01(a * na + fminf(a, b) * lo);
    case 0x04: return clamp01(a * na + (a * b) * lo);
    case 0x05: return clamp01(a * na + (a + b) / 2.0f * lo);
    case 0x06: return clamp01(a * na + (a + b) * lo);
    case 0x07: return clamp01(a * na + ((b + a) - 1.0f) * lo);
    case 0x08:
    case 0x17: return clamp01(a * na + fabsf(a - b) * lo);
    case 0x09: return clamp01(la * na + (1.0f - (1.0f - la) * (1.0f - lb)) * lo);
    case 0x0A: return clamp01(la * na2 + (la > 0.5f ? 1.0f - (1.0f - 2.0f * (la - 0.5f)) * (1.0f - lb) : 2.0f * la * lb) * lo2);
    case 0x0B: return clamp01(la * na2 + (lb > 0.5f ? 1.0f - (1.0f - la) * (1.0f - (lb - 0.5f)) : la * (lb + 0.5f)) * lo2);
    case 0x0C: return clamp01(la * na2 + (lb > 0.5f ? 1.0f - (1.0f - 2.0f * (la - 0.5f)) * (1.0f - lb) : 2.0f * la * lb) * lo2);
    case 0x0D: return clamp01(la * na2 + (lb > 0.5f ? (lb >= 1.0f ? 1.0f : la / (2.0f * (1.0f - lb))) : (lb <= 0.0f ? 0.0f : 1.0f - (1.0f - la) / (2.0f * lb))) * lo2);
    case 0x0E: return clamp01(la * na2 + (la + 2.0f * lb - 1.0f) * lo2);
    case 0x0F: return clamp01(la * na2 + (lb > 0.5f ? fmaxf(la, 2.0f * (lb - 0.5f)) : fminf(la, 2.0f * lb)) * lo2);
    case 0x19: return clamp01(a * na + b * lo);
    default: return a * na + b * lo;
  }
}

/* dt_develop_blend_process() for the four colour spaces (in the raw one the buffers hold one float per site instead of four).  in: the module's input (iw x ih RGBA), out: its output
 * (ow x oh RGBA, roi_out at (xoffs, yoffs) inside roi_in), blended in place; form: the form mask of roi_out or NULL; mask_out: the final
 * mask or NULL.  0 = done (also when blending is off), -1 = not restated. */
int orc_blend_process(const float *in, float *out, int iw, int ih, int ow, int oh, int xoffs, int yoffs, const orc_blend_params_t *d,
                      const float *form, float *mask_out)
{
  (void)ih;
  if(!(d->mask_mode & MASK_ENABLED)) return 0; /* :673 */
  const int lab = d->blend_cst == CS_LAB, display = d->blend_cst == CS_RGB_DISPLAY, raw = d->blend_cst == CS_RAW;
  if(!lab && !raw && ((d->blend_cst != CS_RGB_SCENE && !display) || d->profile_nonlinear)) return -1;
  if(d->feathering_radius > 0.1f || d->blur_radius > 0.1f || d->details != 0.0f) return -1;
  const unsigned channel_mask = lab ? (unsigned)BLENDIF_LAB_MASK : (unsigned)BLENDIF_RGB_MASK;
  orc_fp_fast_mode(); /* the pipe's threads run with FTZ|DAZ (darktable.c:877, common/dtpthread.c:54) */
  blend_plan_t pl;
  memset(&pl, 0, sizeof(pl));
  int parametric = 0; /* :290-312 */
  if(d->mask_mode & MASK_PARAMETRIC)
    for(unsigned ch = 0; ch < BLENDIF_SIZE; ch++)
    {
      if(!(channel_mask & (1u << ch)) || !(d->blendif & (1u << ch))) continue;
      const float *c = d->blendif_parameters + 4 * ch;
      if(fabsf(c[0]) > 1e-6f || fabsf(c[1]) > 1e-6f || fabsf(c[2] - 1.0f) > 1e-6f || fabsf(c[3] - 1.0f) > 1e-6f) parametric = 1;
    }
  const int raster = d->raster_used && form, drawn = d->drawn_used && form;
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
    /* make_mask, :196-325 */
    const unsigned any_active = d->blendif & channel_mask;
    pl.inclusive = (d->mask_combine & COMBINE_INCL) != 0;
    pl.inversed = (d->mask_combine & COMBINE_INV) != 0;
    pl.blendif = d->blendif ^ (pl.inclusive ? channel_mask << 16 : 0u);
    const unsigned canceling = (pl.blendif >> 16) & ~pl.blendif & channel_mask;
    if(raw || !(d->mask_mode & MASK_PARAMETRIC) || (!canceling && !any_active))
      pl.pm = 0; /* the raw space has no channels: opacity and inversion only, blendif_raw.c:36-61 */
    else if(canceling || !any_active)
    {
      pl.pm = 1;
      pl.pm_const = ((pl.inversed == 0) ^ (pl.inclusive == 0)) ? pl.opacity : 0.0f;
    }
    else
    {
      pl.pm = 2;
      blendif_parameters(pl.par, d);
    }
    masking_matrix(d->matrix_in, pl.masking);
    pl.tone = (fabsf(d->contrast) >= 0.01f || fabsf(d->brightness) >= 0.01f) && pl.opacity > 1e-4f; /* :432, :463 */
    pl.contrast_e = expf(3.f * d->contrast);
    pl.brightness = d->brightness;
  }
  const float p = exp2f(d->blend_parameter);
  const int reverse = (d->blend_mode & BLEND_REVERSE) == BLEND_REVERSE;
  if(raw)
  {
    for(int y = 0; y < oh; y++)
      for(int x = 0; x < ow; x++)
      {
        const float av = in[(size_t)(y + yoffs) * iw + xoffs + x];
        float *bv = out + (size_t)y * ow + x;
        const float none[4] = { 0.f, 0.f, 0.f, 0.f };
        const float m = plan_mask(&pl, d, none, none, form ? form[(size_t)y * ow + x] : 0.0f);
        *bv = reverse ? raw_blend_value(d->blend_mode, *bv, av, m) : raw_blend_value(d->blend_mode, av, *bv, m);
        if(mask_out) mask_out[(size_t)y * ow + x] = m;
      }
    return 0;
  }
  for(int y = 0; y < oh; y++)
    for(int x = 0; x < ow; x++)
    {
      const float *a = in + 4 * ((size_t)(y + yoffs) * iw + xoffs + x);
      float *b = out + 4 * ((size_t)y * ow + x);
      const float m = plan_mask(&pl, d, a, b, form ? form[(size_t)y * ow + x] : 0.0f);
      float res[4];
      if(lab)
        lab_blend_pixel(d->blend_mode, reverse ? b : a, reverse ? a : b, m, res);
      else if(display)
        hsl_blend_pixel(d->blend_mode, reverse ? b : a, reverse ? a : b, m, res);
      else if(reverse)
        blend_pixel(d->blend_mode, b, a, p, m, res);
      else
        blend_pixel(d->blend_mode, a, b, p, m, res);
      if(d->mask_display & DISPLAY_MASK) res[3] = a[3]; /* :952-961: an earlier module's mask stays in the alpha lane */
      memcpy(b, res, sizeof(res));
      if(mask_out) mask_out[(size_t)y * ow + x] = m;
    }
  return 0;
}
