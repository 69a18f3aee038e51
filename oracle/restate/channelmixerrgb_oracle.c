/* CPU restatement of colour calibration's pixel loop.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src/iop/channelmixerrgb.c loop_switch :765-959, gamut_mapping :641-706, luma_chroma :707-763;
 * pixel/chromatic_adaptation.h (the Bradford and CAT16 matrices :49-108, bradford_adapt_D50 :178-187, CAT16_adapt_D50
 * :199-207, XYZ_adapt_D50 :217-223, _downscale/_upscale_vector_simd :277-290); system/simd.h dt_mat3x4_mul_vec4
 * :188-197, dt_simd_max_zero :108-114; math/math.h scalar_product :185-195 (the `omp simd reduction` the compiler
 * evaluates as (p0 + p2) + p1), euclidean_norm :206-209, DT_FMA :61-65 (a * b + c without FP_FAST_FMAF).
 * Pinned bit-for-bit against those lines cut verbatim (oracle/_ref: ref_channelmixerrgb.c).  glibc's powf: flt32_math.h.
 */
#include "oracle_common.h"
#include "b200iop.h"
#include "flt32_math.h"

#define NORM_MIN 1.52587890625e-05f
#define INVERSE_SQRT_3 0.5773502691896258f

typedef struct { float v[3]; } v3;

/* columns of the reference's *_transposed tables = rows of the matrices */
static const float XYZ_to_Bradford[3][3] = { { 0.8951f, 0.2664f, -0.1614f }, { -0.7502f, 1.7135f, 0.0367f }, { 0.0389f, -0.0685f, 1.0296f } };
static const float Bradford_to_XYZ[3][3] = { { 0.9870f, -0.1471f, 0.1600f }, { 0.4323f, 0.5184f, 0.0493f }, { -0.0085f, 0.0400f, 0.9685f } };
static const float XYZ_to_CAT16[3][3] = { { 0.401288f, 0.650173f, -0.051461f }, { -0.250268f, 1.204414f, 0.045854f }, { -0.002079f, 0.048952f, 0.953127f } };
static const float CAT16_to_XYZ[3][3] = { { 1.862068f, -1.011255f, 0.149187f }, { 0.38752f, 0.621447f, -0.008974f }, { -0.015841f, -0.034123f, 1.049964f } };

/* dt_mat3x4_mul_vec4 with the rows of the transposed matrix: out[r] = (M[r][0] * in0 + M[r][1] * in1) + M[r][2] * in2, as
 * out = row0 * in0; out = row1 * in1 + out; out = row2 * in2 + out */
static v3 mul(const float M[3][3], v3 in)
{
  v3 o;
  for(int r = 0; r < 3; r++)
  {
    float acc = M[r][0] * in.v[0];
    acc = M[r][1] * in.v[1] + acc;
    acc = M[r][2] * in.v[2] + acc;
    o.v[r] = acc;
  }
  return o;
}
static v3 mul4(const float M[3][4], v3 in)
{
  const float m[3][3] = { { M[0][0], M[0][1], M[0][2] }, { M[1][0], M[1][1], M[1][2] }, { M[2][0], M[2][1], M[2][2] } };
  return mul(m, in);
}
static float max_zero(float v) { return isfinite(v) ? (v > 0.0f ? v : 0.0f) : 0.f; }
static v3 max_zero3(v3 a)
{
  for(int c = 0; c < 3; c++) a.v[c] = max_zero(a.v[c]);
  return a;
}
static float scale_of(float Y) { return ((Y > NORM_MIN) && !isnan(Y)) ? (Y + NORM_MIN) : NORM_MIN; }
static v3 downscale(v3 a, float Y)
{
  const float s = scale_of(Y);
  for(int c = 0; c < 3; c++) a.v[c] = a.v[c] / s;
  return a;
}
static v3 upscale(v3 a, float Y)
{
  const float s = scale_of(Y);
  for(int c = 0; c < 3; c++) a.v[c] = a.v[c] * s;
  return a;
}
static float sqf(float x) { return x * x; }
static float dot3(const float *a, const float *b) { return 0.f + ((a[0] * b[0] + a[2] * b[2]) + a[1] * b[1]); }
static float norm3(const float *a) { return fmaxf(sqrtf(sqf(a[0]) + sqf(a[1]) + sqf(a[2])), NORM_MIN); }

static v3 gamut_mapping(v3 input, float compression, int clip)
{
  const float sum = input.v[0] + input.v[1] + input.v[2];
  const float Y = input.v[1];
  v3 o = { { 0.f, 0.f, 0.f } };
  if(sum > 0.f && Y > 0.f)
  {
    float x = input.v[0] / sum, y = input.v[1] / sum;
    const float uv_denominator = -2.f * x + 12.f * y + 3.f;
    float u = 4.f * x / uv_denominator, v = 9.f * y / uv_denominator;
    const float D50[2] = { 0.20915914598542354f, 0.488075320769787f };
    const float delta[2] = { D50[0] - u, D50[1] - v };
    const float Delta = Y * (sqf(delta[0]) + sqf(delta[1]));
    const float correction = (compression == 0.0f) ? 0.f : f32m_powf(Delta, compression);
    const float tmp_u = correction * delta[0] + u, tmp_v = correction * delta[1] + v;
    u = (u > D50[0]) ? fmaxf(tmp_u, D50[0]) : fminf(tmp_u, D50[0]);
    v = (v > D50[1]) ? fmaxf(tmp_v, D50[1]) : fminf(tmp_v, D50[1]);
    const float xy_denominator = 6.f * u - 16.f * v + 12.f;
    x = 9.f * u / xy_denominator;
    y = 4.f * v / xy_denominator;
    if(clip)
    {
      x = fmaxf(x, 0.0f);
      y = fmaxf(y, 0.0f);
    }
    y = fmaxf(y, NORM_MIN);
    const float scale = x + y;
    if(scale >= 1.f)
    {
      x /= scale;
      y /= scale;
    }
    o.v[0] = Y * x / y;
    o.v[1] = Y;
    o.v[2] = Y * (1.f - x - y) / y;
  }
  return o;
}

static v3 luma_chroma(v3 in, const float *saturation, const float *lightness, int version)
{
  const float *input = in.v;
  v3 out;
  float *output = out.v;
  float norm = norm3(input);
  const float avg = fmaxf((input[0] + input[1] + input[2]) / 3.0f, NORM_MIN);
  if(norm > 0.f && avg > 0.f)
  {
    const float mix = dot3(input, lightness);
    if(version == 2) norm *= INVERSE_SQRT_3;
    for(int c = 0; c < 3; c++) output[c] = input[c] / norm;
    float coeff_ratio = 0.f;
    if(version == 0)
      for(int c = 0; c < 3; c++) coeff_ratio += sqf(1.0f - output[c]) * saturation[c];
    else
      coeff_ratio = dot3(output, saturation) / 3.f;
    for(int c = 0; c < 3; c++)
    {
      const float min_ratio = (output[c] < 0.0f) ? output[c] : 0.0f;
      const float output_inverse = 1.0f - output[c];
      output[c] = fmaxf(output_inverse * coeff_ratio + output[c], min_ratio);
    }
    if(version == 2) norm /= norm3(output) * INVERSE_SQRT_3;
    norm *= fmaxf(1.f + mix / avg, 0.f);
    for(int c = 0; c < 3; c++) output[c] *= norm;
  }
  else
    for(int c = 0; c < 3; c++) output[c] = input[c];
  return out;
}

int orc_channelmixerrgb(const float *in, float *out, int width, int height, const b200_channelmixerrgb_piece_t *pc)
{
  const b200_channelmixerrgb_data_t *d = &pc->data;
  const int kind = d->adaptation, clip = d->clip;
  if(kind < B200_ADAPTATION_LINEAR_BRADFORD || kind > B200_ADAPTATION_RGB) return 0; /* DT_ADAPTATION_LAST / default: nothing is written */
  const float(*R2X)[4] = pc->work_profile.matrix_in, (*X2R)[4] = pc->work_profile.matrix_out;
  const int bradford = kind == B200_ADAPTATION_LINEAR_BRADFORD || kind == B200_ADAPTATION_FULL_BRADFORD;
  const float(*to_lms)[3] = bradford ? XYZ_to_Bradford : XYZ_to_CAT16, (*to_xyz)[3] = bradford ? Bradford_to_XYZ : CAT16_to_XYZ;
  const int lms = bradford || kind == B200_ADAPTATION_CAT16;
  static const float D50_bradford[3] = { 0.996078f, 1.020646f, 0.818155f }, D50_cat16[3] = { 0.994535f, 1.000997f, 0.833036f },
                     D50_xyz[3] = { 0.9642119944211994f, 1.0f, 0.8251882845188288f };
  for(size_t k = 0; k < (size_t)width * height; k++)
  {
    const float *px = in + 4 * k;
    v3 two = { { px[0], px[1], px[2] } }, one;
    if(clip) two = max_zero3(two);
    if(kind == B200_ADAPTATION_RGB)
    {
      one = mul4(d->MIX, two);
      one = mul4(R2X, one);
    }
    else
    {
      one = mul4(R2X, two);
      const float Y = one.v[1];
      if(kind == B200_ADAPTATION_XYZ)
      {
        two = downscale(one, Y);
        for(int c = 0; c < 3; c++) two.v[c] = two.v[c] * D50_xyz[c] / d->illuminant[c];
        two = upscale(two, Y);
        one = mul4(d->MIX, two);
      }
      else
      {
        two = downscale(mul(to_lms, one), Y);
        if(bradford)
        {
          for(int c = 0; c < 3; c++) two.v[c] = two.v[c] / d->illuminant[c];
          if(kind == B200_ADAPTATION_FULL_BRADFORD) two.v[2] = (two.v[2] > 0.f) ? f32m_powf(two.v[2], d->p) : two.v[2];
          for(int c = 0; c < 3; c++) two.v[c] = D50_bradford[c] * two.v[c];
        }
        else
          for(int c = 0; c < 3; c++) two.v[c] = two.v[c] * D50_cat16[c] / d->illuminant[c]; /* CAT16_adapt_D50(.., 1.0f, TRUE) */
        one = upscale(two, Y);
        two = mul4(d->MIX, one);
        one = mul(to_xyz, two);
      }
    }
    two = gamut_mapping(one, d->gamut, clip);
    one = lms ? mul(to_lms, two) : (kind == B200_ADAPTATION_XYZ ? two : mul4(X2R, two));
    if(clip) one = max_zero3(one);
    two = luma_chroma(one, d->saturation, d->lightness, d->version);
    if(clip) two = max_zero3(two);
    float *o = out + 4 * k;
    if(d->apply_grey)
    {
      const float grey_mix = fmaxf(two.v[0] * d->grey[0] + two.v[1] * d->grey[1] + two.v[2] * d->grey[2], 0.0f);
      o[0] = o[1] = o[2] = grey_mix;
    }
    else
    {
      one = lms ? mul(to_xyz, two) : (kind == B200_ADAPTATION_XYZ ? two : mul4(R2X, two));
      if(clip) one = max_zero3(one);
      two = mul4(X2R, one);
      if(clip) two = max_zero3(two);
      o[0] = two.v[0], o[1] = two.v[1], o[2] = two.v[2];
    }
    o[3] = px[3];
  }
  return 0;
}
