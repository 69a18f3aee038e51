// Blending of a module's output over its input (mask + blend operator), scene-referred RGB space, for B200 / sm_100a.
//
// What the reference computes: develop/blend.c dt_develop_blend_process :657-860 with blend_cst == DEVELOP_BLEND_CS_RGB_SCENE:
// the mask (uniform opacity | the raster / drawn mask the host rasterised | the parametric mask of
// develop/blends/blendif_rgb_jzczhz.c :42-325 on the gray, red, green and blue channels of the module's input and output, combined
// exclusively or inclusively, inverted or not | the mask tone curve :626-655), then one of the sixteen blend operators :328-649,
// the result in place of the module's output with the mask in its alpha lane (:878-961).  Parity contract: bit-identical to those
// lines under C float semantics (oracle/restate/blend_oracle.c, pinned against them compiled in place).
//
// The reference makes up to seven passes over full buffers (seed, one per parametric channel set, opacity, tone curve, a copy of
// the output, the operator, the alpha copy).  Every one of them is pointwise, so the kernel is ONE pass: 16 B of input, 16 B of
// output and 4 B of form mask in, 16 B out (+ 4 B when the caller wants the mask, e.g. to publish it as a raster mask) --
// 52 B/px algorithmic.  What the host decides once per call (which of the reference's branches a parameter block takes, the
// slopes of the parametric channels, exp2f / expf of the parameters) arrives in the plan; what depends on the pixel is evaluated
// here.  Not built (B200_ERR_UNSUPPORTED, the caller falls back to the reference's own path): feathering (guided filter), Gaussian
// blur and detail refinement of the mask, the JzCzhz channels of the parametric mask, the GUI's channel display, the Lab, display
// RGB and raw colour spaces.
#ifndef B200_KERNELS_ON_CPU // tests/emul compiles the kernel of this file with g++ to check it against the oracle without a GPU
#include "runtime.h"
#endif
#include <float.h>
#include <math.h>
#include <string.h>

namespace
{
enum
{
  MASK_ENABLED = 1, MASK_SHAPE = 2, MASK_PARAMETRIC = 4, MASK_RASTER = 8, // dt_develop_mask_mode_t, blend.h:110-118
  COMBINE_INV = 1, COMBINE_INCL = 2,                                      // dt_develop_mask_combine_mode_t :120-131
  BLENDIF_SIZE = 16, BLENDIF_ITEMS = 6, BLENDIF_RGB_MASK = 0x77FF,        // :188-191, :329
  CS_RGB_SCENE = 4                                                         // :52-59
};
constexpr unsigned BLEND_REVERSE = 0x80000000u; // blend.h:106

struct blend_plan_t
{
  const float4 *in;
  float4 *out;
  const float *form;
  float *mask_out;
  int iw, ow, oh, xoffs, yoffs;
  int kind;      // 0: mask = opacity; 1: mask = form * opacity (a raster mask alone); 2: seed, then the parametric stage
  int seed_form; // kind 2: the seed is the form mask, else `fill`
  float fill, opacity;
  int pm;        // parametric stage: 0 = opacity * m (or opacity * (1 - m) inverted); 1 = the constant pm_const; 2 = channels
  int inversed, inclusive;
  float pm_const;
  unsigned blendif;
  float par[BLENDIF_ITEMS * 8]; // gray, red, green, blue of the input, then of the output: four limits and two slopes each
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
{ // _blendif_combine_channels(), :151-194, without the JzCzhz set
  if(blendif & 1u) t *= bl_factor(lum[0] * px[0] + lum[1] * px[1] + lum[2] * px[2], (blendif >> 16) & 1u, par);
#pragma unroll
  for(int c = 0; c < 3; c++)
    if(blendif & (2u << c)) t *= bl_factor(px[c], (blendif >> 16) & (2u << c), par + BLENDIF_ITEMS * (1 + c));
  return t;
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
    float t = bl_channels(a, 1.0f, pl.blendif, pl.par, pl.lum);
    t = bl_channels(b, t, pl.blendif >> 4, pl.par + BLENDIF_ITEMS * 4, pl.lum);
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

__global__ void __launch_bounds__(256) blend_kernel(const __grid_constant__ blend_plan_t pl)
{
  const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
  if(x >= pl.ow) return;
  const size_t o = (size_t)y * pl.ow + x;
  const float4 av = __ldg(pl.in + (size_t)(y + pl.yoffs) * pl.iw + pl.xoffs + x), bv = pl.out[o];
  const float a[4] = { av.x, av.y, av.z, av.w }, b[4] = { bv.x, bv.y, bv.z, bv.w };
  const float m = bl_mask(pl, a, b, pl.form ? __ldg(pl.form + o) : 0.0f);
  float res[4];
  if(pl.reverse)
    bl_operator(pl.mode, b, a, pl.p, m, res);
  else
    bl_operator(pl.mode, a, b, pl.p, m, res);
  if(pl.keep_alpha) res[3] = a[3]; // :952-961: an earlier module's mask stays in the alpha lane
  pl.out[o] = make_float4(res[0], res[1], res[2], res[3]);
  if(pl.mask_out) pl.mask_out[o] = m;
}

// dt_develop_blendif_process_parameters(), blend.c:214-260, for the eight RGB channels (no Lab offset)
void bl_parameters(float *par, const b200_blend_params_t *d)
{
  for(int i = 0; i < 8; i++)
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

// which of the reference's branches a parameter block takes (blend.c:669-760, blendif_rgb_jzczhz.c:196-240); 1 = blending is off,
// < 0 = an error code
int bl_plan(blend_plan_t &pl, const b200_blend_params_t *d, bool have_form)
{
  memset(&pl, 0, sizeof(pl));
  if(!(d->mask_mode & MASK_ENABLED)) return 1; // :673
  if(d->blend_cst != CS_RGB_SCENE) return B200_ERR_UNSUPPORTED;
  if(d->profile_nonlinear) return B200_ERR_UNSUPPORTED;
  if(d->feathering_radius > 0.1f || d->blur_radius > 0.1f || d->details != 0.0f) return B200_ERR_UNSUPPORTED;
  if((d->mask_mode & MASK_PARAMETRIC) && (d->blendif & 0x7700u)) return B200_ERR_UNSUPPORTED;
  bool parametric = false; // dt_develop_blend_get_mask_usage(), :290-312
  if(d->mask_mode & MASK_PARAMETRIC)
    for(unsigned ch = 0; ch < BLENDIF_SIZE; ch++)
    {
      if(!(BLENDIF_RGB_MASK & (1u << ch)) || !(d->blendif & (1u << ch))) continue;
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
  blend_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pl);
  B200_CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_blend_process_host(const b200_piece_t *piece, const b200_blend_params_t *bp, const void *in, void *out, const float *form_mask, float *mask)
{
  int rc = blend_check(piece, bp, in, out);
  if(rc) return rc;
  if((rc = bind_device(piece->devid))) return rc;
  const size_t ibytes = (size_t)piece->roi_in.width * piece->roi_in.height * 16, opx = (size_t)piece->roi_out.width * piece->roi_out.height;
  void *d_in = nullptr, *d_out = nullptr, *d_form = nullptr, *d_mask = nullptr;
  cudaStream_t s;
  if((rc = host_stream(&s))) return rc;
  if((rc = scratch(SLOT_IN, ibytes, &d_in))) return rc;
  if((rc = scratch(SLOT_OUT, opx * 16, &d_out))) return rc;
  if((rc = copy_h2d(d_in, in, ibytes, s))) return rc;
  if((rc = copy_h2d(d_out, out, opx * 16, s))) return rc;
  if(form_mask)
  {
    if((rc = scratch(SLOT_TMP0, opx * 4, &d_form))) return rc;
    if((rc = copy_h2d(d_form, form_mask, opx * 4, s))) return rc;
  }
  if(mask && (rc = scratch(SLOT_TMP1, opx * 4, &d_mask))) return rc;
  if((rc = b200_blend_process_dev(piece, bp, d_in, d_out, (const float *)d_form, (float *)d_mask, (void *)s))) return rc;
  if((rc = copy_d2h(out, d_out, opx * 16, s))) return rc;
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
