/* CPU restatement of a module's blending (mask + blend operator) in the scene-referred RGB space.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src: develop/blend.c dt_develop_blend_process :657-860 (mask usage :262-320, post operations :427-469,
 * dt_develop_blendif_process_parameters :214-260, _develop_blend_process_mask_tone_curve :626-655) and
 * develop/blends/blendif_rgb_jzczhz.c (_blendif_compute_factor :42-73, the gray / red / green / blue channels :75-121,
 * _blendif_combine_channels :151-194, dt_develop_blendif_rgb_jzczhz_make_mask :196-325, the sixteen operators :328-585,
 * _choose_blend_func :587-649, dt_develop_blendif_rgb_jzczhz_blend :878-961).  Pinned bit-for-bit against those lines cut verbatim
 * (oracle/_ref: ref_blend.c).
 *
 * Everything here is a function of one pixel of the module's input, the same pixel of its output and the same pixel of the form
 * mask (the raster / drawn mask the host rasterised): the reference's passes over whole buffers are folded into one evaluation per
 * pixel, which is also how the CUDA kernel does it.  Not restated (the entry point returns -1, the product B200_ERR_UNSUPPORTED):
 * feathering (guided filter), Gaussian blur and detail refinement of the mask, the JzCzhz channels of the parametric mask, the GUI's
 * channel display, the other blend colour spaces.
 */
#include "oracle_common.h"
#include <float.h>
#include <stdlib.h>
#include <string.h>

enum
{
  MASK_ENABLED = 1, MASK_SHAPE = 2, MASK_PARAMETRIC = 4, MASK_RASTER = 8, /* dt_develop_mask_mode_t, blend.h:110-118 */
  COMBINE_INV = 1, COMBINE_INCL = 2,                                      /* dt_develop_mask_combine_mode_t :120-131 */
  BLENDIF_SIZE = 16, BLENDIF_ITEMS = 6, BLENDIF_RGB_MASK = 0x77FF,        /* :188-191, :329 */
  CS_RGB_SCENE = 4,                                                        /* :52-59 */
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
} orc_blend_params_t;

/* :214-260 for the RGB spaces (no Lab offset) */
static void blendif_parameters(float *par, const orc_blend_params_t *d)
{
  for(int i = 0; i < BLENDIF_SIZE; i++)
  {
    float *p = par + BLENDIF_ITEMS * i;
    const float *b = d->blendif_parameters + 4 * i;
    if(d->blendif & (1u << i))
    {
      const float boost = exp2f(d->blendif_boost_factors[i]);
      for(int k = 0; k < 4; k++) p[k] = (b[k] - 0.0f) * boost;
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
  int tone;        /* mask tone curve :626-655 */
  float contrast_e, brightness;
} blend_plan_t;

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
    float t = blendif_channels(a, 1.0f, pl->blendif, pl->par, d->luminance);
    t = blendif_channels(b, t, pl->blendif >> 4, pl->par + BLENDIF_ITEMS * 4, d->luminance);
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

/* dt_develop_blend_process() for blend_cst == DEVELOP_BLEND_CS_RGB_SCENE.  in: the module's input (iw x ih RGBA), out: its output
 * (ow x oh RGBA, roi_out at (xoffs, yoffs) inside roi_in), blended in place; form: the form mask of roi_out or NULL; mask_out: the final
 * mask or NULL.  0 = done (also when blending is off), -1 = not restated. */
int orc_blend_process(const float *in, float *out, int iw, int ih, int ow, int oh, int xoffs, int yoffs, const orc_blend_params_t *d,
                      const float *form, float *mask_out)
{
  (void)ih;
  if(!(d->mask_mode & MASK_ENABLED)) return 0; /* :673 */
  if(d->blend_cst != CS_RGB_SCENE || d->profile_nonlinear) return -1;
  if(d->feathering_radius > 0.1f || d->blur_radius > 0.1f || d->details != 0.0f) return -1;
  if((d->mask_mode & MASK_PARAMETRIC) && (d->blendif & 0x7700u)) return -1;
  orc_fp_fast_mode(); /* the pipe's threads run with FTZ|DAZ (darktable.c:877, common/dtpthread.c:54) */
  blend_plan_t pl;
  memset(&pl, 0, sizeof(pl));
  int parametric = 0; /* :290-312 */
  if(d->mask_mode & MASK_PARAMETRIC)
    for(unsigned ch = 0; ch < BLENDIF_SIZE; ch++)
    {
      if(!(BLENDIF_RGB_MASK & (1u << ch)) || !(d->blendif & (1u << ch))) continue;
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
    const unsigned any_active = d->blendif & BLENDIF_RGB_MASK;
    pl.inclusive = (d->mask_combine & COMBINE_INCL) != 0;
    pl.inversed = (d->mask_combine & COMBINE_INV) != 0;
    pl.blendif = d->blendif ^ (pl.inclusive ? (unsigned)BLENDIF_RGB_MASK << 16 : 0u);
    const unsigned canceling = (pl.blendif >> 16) & ~pl.blendif & BLENDIF_RGB_MASK;
    if(!(d->mask_mode & MASK_PARAMETRIC) || (!canceling && !any_active))
      pl.pm = 0;
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
    pl.tone = (fabsf(d->contrast) >= 0.01f || fabsf(d->brightness) >= 0.01f) && pl.opacity > 1e-4f; /* :432, :463 */
    pl.contrast_e = expf(3.f * d->contrast);
    pl.brightness = d->brightness;
  }
  const float p = exp2f(d->blend_parameter);
  const int reverse = (d->blend_mode & BLEND_REVERSE) == BLEND_REVERSE;
  for(int y = 0; y < oh; y++)
    for(int x = 0; x < ow; x++)
    {
      const float *a = in + 4 * ((size_t)(y + yoffs) * iw + xoffs + x);
      float *b = out + 4 * ((size_t)y * ow + x);
      const float m = plan_mask(&pl, d, a, b, form ? form[(size_t)y * ow + x] : 0.0f);
      float res[4];
      if(reverse)
        blend_pixel(d->blend_mode, b, a, p, m, res);
      else
        blend_pixel(d->blend_mode, a, b, p, m, res);
      if(d->mask_display & DISPLAY_MASK) res[3] = a[3]; /* :952-961: an earlier module's mask stays in the alpha lane */
      memcpy(b, res, sizeof(res));
      if(mask_out) mask_out[(size_t)y * ow + x] = m;
    }
  return 0;
}
