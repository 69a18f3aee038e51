/* CPU restatement of filmic rgb, AgX colour science (v8 family).  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src/iop/filmicrgb.c
 *   process() AgX branch :2843-2856, filmic_agx :2495-2587, filmic_agx_compress_negatives :2461-2492,
 *   filmic_agx_prepare_bracket :2390-2459, _filmic_agx_build_displaced :2344-2388,
 *   filmic_v4_prepare_matrices :2033-2063, pipe_RGB_to_Ych_simd :1740-1760, Ych_to_pipe_RGB_simd :1763-1777,
 *   RGB_tone_mapping_v4_simd :2133-2149, log_tonemapping :1046-1051, filmic_spline :1062-1160,
 *   gamut_mapping_simd :1986-2030, gamut_check_RGB_simd :1949-1983, gamut_check_Yrg_filmic_simd :1928-1946,
 *   clip_chroma{,_white,_white_raw,_black} :1826-1925, filmic_desaturate_v4 :1779-1816
 * and common/colorspaces_inline_conversions.h (LMS/Yrg/gradingRGB :902-1075), pixel/chromatic_adaptation.h
 * (CAT16 matrices :248-261), math/matrices.h (dt_colormatrix_mul :167-179, mat3SSEinv :37-65,
 * dot_product :201-206), system/simd.h (dt_mat3x4_mul_vec4 :188-197).
 *
 * Pinned bit-for-bit (tests/test_cpu_oracle_pin.py) against those very functions cut verbatim out of
 * filmicrgb.c and compiled with C-standard semantics (oracle/_ref/libref_strict.so, ref_filmic.c).
 * powf/log2f are glibc's (flt32_math.h).  The colour sciences before AgX follow further down (orc_filmic_legacy);
 * the highlight reconstruction in front of either is filmic_reconstruct_oracle.c.
 */
#include "oracle_common.h"
#include "flt32_math.h"
#include "b200iop.h"
#include <float.h>
#include <string.h>

typedef float m34[3][4]; /* three rows of a dt_colormatrix_t */
#define CLAMPF(a, mn, mx) ((a) >= (mn) ? ((a) <= (mx) ? (a) : (mx)) : (mn)) /* math/math.h:91 */
#define CLAMP(x, lo, hi) (((x) > (hi)) ? (hi) : (((x) < (lo)) ? (lo) : (x))) /* glib */
#define MINF(a, b) (((a) < (b)) ? (a) : (b))
#define MAXF(a, b) (((a) > (b)) ? (a) : (b))
#define Y31_TO_Y06(x) (1.05785528f * (x)) /* filmicrgb.c:1823 */

static const m34 XYZ_D50_to_D65_CAT16 = { { 9.89466254e-01f, -4.00304626e-02f, 4.40530317e-02f, 0.f },
                                          { -5.40518733e-03f, 1.00666069e+00f, -1.75551955e-03f, 0.f },
                                          { -4.03920992e-04f, 1.50768030e-02f, 1.30210211e+00f, 0.f } };
static const m34 XYZ_D65_to_D50_CAT16 = { { 1.01085433e+00f, 4.07086103e-02f, -3.41445825e-02f, 0.f },
                                          { 5.42814201e-03f, 9.93581926e-01f, 1.15592039e-03f, 0.f },
                                          { 2.50722468e-04f, -1.14918759e-02f, 7.67964947e-01f, 0.f } };
static const m34 XYZ_D65_to_LMS = { { 0.257085f, 0.859943f, -0.031061f, 0.f },
                                    { -0.394427f, 1.175800f, 0.106423f, 0.f },
                                    { 0.064856f, -0.076250f, 0.559067f, 0.f } };
static const m34 LMS_to_XYZ_D65 = { { 1.80794659f, -1.29971660f, 0.34785879f, 0.f },
                                    { 0.61783960f, 0.39595453f, -0.04104687f, 0.f },
                                    { -0.12546960f, 0.20478038f, 1.74274183f, 0.f } };
static const m34 filmlightRGB_to_LMS = { { 0.95f, 0.38f, 0.00f, 0.f }, { 0.05f, 0.62f, 0.03f, 0.f }, { 0.00f, 0.00f, 0.97f, 0.f } };
static const m34 LMS_to_filmlightRGB = { { 1.0877193f, -0.66666667f, 0.02061856f, 0.f },
                                         { -0.0877193f, 1.66666667f, -0.05154639f, 0.f },
                                         { 0.f, 0.f, 1.03092784f, 0.f } };

/* dt_mat3x4_mul_vec4 on the transposed rows of M: lane i = (M[i][0]*x + M[i][1]*y) + M[i][2]*z, lane 3 the
 * same expression over zeros */
static inline void mat4(const m34 M, const float in[4], float out[4])
{
  float r[4];
  for(int i = 0; i < 4; i++)
  {
    const float a = i < 3 ? M[i][0] : 0.f, b = i < 3 ? M[i][1] : 0.f, c = i < 3 ? M[i][2] : 0.f;
    float acc = a * in[0];
    acc = b * in[1] + acc;
    r[i] = c * in[2] + acc;
  }
  memcpy(out, r, sizeof(r));
}
/* dot_product / scalar_product, matrices.h:201-206, math.h:185-195.  scalar_product() is an
 * `omp simd reduction(+:acc)` loop: gcc evaluates it as a 4-lane vector (p0, p1, p2, 0) reduced
 * pairwise, (p0 + p2) + (p1 + 0), also in the strict build -- verified against oracle/_ref.  Host-side
 * set-up only (the displaced primaries); no per-pixel code goes through it. */
static inline void dot3(const float v[4], const m34 M, float out[4])
{
  for(int i = 0; i < 3; i++)
  {
    const float p0 = v[0] * M[i][0], p1 = v[1] * M[i][1], p2 = v[2] * M[i][2];
    out[i] = (p0 + p2) + p1;
  }
}
/* dt_colormatrix_mul, matrices.h:167-179 (all four columns, three rows) */
static void mmul(m34 dst, const m34 m1, const m34 m2)
{
  m34 r;
  for(int k = 0; k < 3; k++)
    for(int i = 0; i < 4; i++)
    {
      float sum = 0.0f;
      for(int j = 0; j < 3; j++) sum += m1[k][j] * m2[j][i];
      r[k][i] = sum;
    }
  memcpy(dst, r, sizeof(r));
}
/* mat3SSEinv, matrices.h:37-65 */
static int minv(m34 dst, const m34 src)
{
#define A(y, x) src[(y - 1)][(x - 1)]
#define B(y, x) dst[(y - 1)][(x - 1)]
  const float det = A(1, 1) * (A(3, 3) * A(2, 2) - A(3, 2) * A(2, 3)) - A(2, 1) * (A(3, 3) * A(1, 2) - A(3, 2) * A(1, 3))
                    + A(3, 1) * (A(2, 3) * A(1, 2) - A(2, 2) * A(1, 3));
  if(fabsf(det) < 1e-7f) return 1;
  const float invDet = 1.f / det;
  B(1, 1) = invDet * (A(3, 3) * A(2, 2) - A(3, 2) * A(2, 3));
  B(1, 2) = -invDet * (A(3, 3) * A(1, 2) - A(3, 2) * A(1, 3));
  B(1, 3) = invDet * (A(2, 3) * A(1, 2) - A(2, 2) * A(1, 3));
  B(2, 1) = -invDet * (A(3, 3) * A(2, 1) - A(3, 1) * A(2, 3));
  B(2, 2) = invDet * (A(3, 3) * A(1, 1) - A(3, 1) * A(1, 3));
  B(2, 3) = -invDet * (A(2, 3) * A(1, 1) - A(2, 1) * A(1, 3));
  B(3, 1) = invDet * (A(3, 2) * A(2, 1) - A(3, 1) * A(2, 2));
  B(3, 2) = -invDet * (A(3, 2) * A(1, 1) - A(3, 1) * A(1, 2));
  B(3, 3) = invDet * (A(2, 2) * A(1, 1) - A(2, 1) * A(1, 2));
#undef A
#undef B
  return 0;
}
static void ident(m34 M)
{
  memset(M, 0, sizeof(m34));
  M[0][0] = M[1][1] = M[2][2] = 1.f;
}

/* ---- scalar (host set-up) colour helpers: colorspaces_inline_conversions.h:1014-1063 ------------- */
static void lms_to_yrg(const float LMS[4], float Yrg[4])
{
  const float Y = 0.68990272f * LMS[0] + 0.34832189f * LMS[1];
  const float a = LMS[0] + LMS[1] + LMS[2];
  float lms[4] = { 0 }, rgb[4] = { 0 };
  for(int c = 0; c < 4; c++) lms[c] = (a == 0.f) ? 0.f : LMS[c] / a;
  dot3(lms, LMS_to_filmlightRGB, rgb);
  Yrg[0] = Y;
  Yrg[1] = rgb[0];
  Yrg[2] = rgb[1];
}
static void yrg_to_lms(const float Yrg[4], float LMS[4])
{
  const float Y = Yrg[0], r = Yrg[1], g = Yrg[2], b = 1.f - r - g;
  const float rgb[4] = { r, g, b, 0.f };
  float lms[4] = { 0 };
  dot3(rgb, filmlightRGB_to_LMS, lms);
  const float denom = (0.68990272f * lms[0] + 0.34832189f * lms[1]);
  const float a = (denom == 0.f) ? 0.f : Y / denom;
  for(int c = 0; c < 4; c++) LMS[c] = lms[c] * a;
}
static void xyz50_to_yrg(const float xyz[4], float Yrg[4])
{ /* filmicrgb.c:2314-2321 */
  float d65[4] = { 0 }, lms[4] = { 0 };
  dot3(xyz, XYZ_D50_to_D65_CAT16, d65);
  dot3(d65, XYZ_D65_to_LMS, lms);
  lms_to_yrg(lms, Yrg);
}
static void yrg_to_xyz50(const float Yrg[4], float xyz[4])
{ /* filmicrgb.c:2323-2330 */
  float lms[4] = { 0 }, d65[4] = { 0 };
  yrg_to_lms(Yrg, lms);
  dot3(lms, LMS_to_XYZ_D65, d65);
  dot3(d65, XYZ_D65_to_D50_CAT16, xyz);
}

/* filmicrgb.c:2344-2388 */
static int build_displaced(const m34 work_in, const m34 work_out, const float inset[3], const float rotation[3], m34 M)
{
  float white_xyz[4] = { 0 }, white_Yrg[4] = { 0 };
  for(int r = 0; r < 3; r++)
    for(int c = 0; c < 3; c++) white_xyz[r] += work_in[r][c];
  xyz50_to_yrg(white_xyz, white_Yrg);
  m34 P = { { 0 } };
  for(int i = 0; i < 3; i++)
  {
    const float pxyz[4] = { work_in[0][i], work_in[1][i], work_in[2][i], 0.f };
    float pY[4] = { 0 };
    xyz50_to_yrg(pxyz, pY);
    const float dr = pY[1] - white_Yrg[1], dg = pY[2] - white_Yrg[2];
    const float scale = 1.f - CLAMPF(inset[i], 0.f, 0.9f);
    const float cos_a = cosf(rotation[i]), sin_a = sinf(rotation[i]);
    const float dY[4] = { pY[0], white_Yrg[1] + scale * (cos_a * dr - sin_a * dg), white_Yrg[2] + scale * (sin_a * dr + cos_a * dg), 0.f };
    float dxyz[4] = { 0 };
    yrg_to_xyz50(dY, dxyz);
    for(int r = 0; r < 3; r++) P[r][i] = dxyz[r];
  }
  m34 Pinv = { { 0 } };
  if(minv(Pinv, P)) return 0;
  float s[4] = { 0 };
  dot3(white_xyz, Pinv, s);
  for(int r = 0; r < 3; r++)
    for(int c = 0; c < 3; c++) P[r][c] *= s[c];
  mmul(M, work_out, P);
  return 1;
}

typedef struct
{
  m34 input, output, export_input, export_output, inset, outset;
  float luma[4];
  int use_output_profile;
} filmic_mats_t;

/* filmic_v4_prepare_matrices :2033-2063 + filmic_agx_prepare_bracket :2390-2459 */
static void prepare(filmic_mats_t *m, int version, const m34 work_in, const m34 work_out, const m34 *exp_in, const m34 *exp_out)
{
  memset(m, 0, sizeof(*m));
  m34 tmp;
  mmul(tmp, XYZ_D50_to_D65_CAT16, work_in);
  mmul(m->input, XYZ_D65_to_LMS, tmp);
  mmul(tmp, XYZ_D65_to_D50_CAT16, LMS_to_XYZ_D65);
  mmul(m->output, work_out, tmp);
  m->use_output_profile = exp_in != NULL;
  if(exp_in)
  {
    mmul(tmp, XYZ_D65_to_D50_CAT16, LMS_to_XYZ_D65);
    mmul(m->export_output, *exp_out, tmp);
    mmul(tmp, XYZ_D50_to_D65_CAT16, *exp_in);
    mmul(m->export_input, XYZ_D65_to_LMS, tmp);
  }
  static const float K[5][12] = {
    /* V6 no bleach */ { 0.5991055f, 0.6000000f, 0.3300009f, 0.0571015f, 0.1999891f, 0.0886110f, 0.761433f, 0.752267f, 0.465293f, -0.0034297f, 0.1952448f, -0.0480109f },
    /* V7 low      */ { 0.6410825f, 0.6898110f, 0.3194529f, 0.0405734f, 0.1631286f, 0.0350584f, 0.784757f, 0.789387f, 0.445403f, -0.0057845f, 0.1593207f, -0.0592955f },
    /* V8 medium   */ { 0.6509540f, 0.7488775f, 0.3517703f, 0.0278602f, 0.1214671f, -0.0228829f, 0.793082f, 0.815169f, 0.460318f, -0.0053781f, 0.1187604f, -0.0794801f },
    /* V9 high     */ { 0.6379749f, 0.7878689f, 0.3753822f, 0.0106096f, 0.0582598f, -0.0696729f, 0.790237f, 0.831376f, 0.465406f, -0.0080070f, 0.0571100f, -0.0912220f },
    /* V10 extra   */ { 0.5770235f, 0.8102094f, 0.4000390f, -0.0081060f, -0.0034008f, -0.1035236f, 0.766420f, 0.838020f, 0.465130f, -0.0122011f, -0.0021732f, -0.0971215f },
  };
  const int row = (version >= 5 && version <= 9) ? version - 5 : 0;
  m34 rec = { { 0 } };
  if(!build_displaced(work_in, work_out, K[row], K[row] + 3, m->inset) || !build_displaced(work_in, work_out, K[row] + 6, K[row] + 9, rec)
     || minv(m->outset, rec))
  {
    ident(m->inset);
    ident(m->outset);
  }
  m->luma[0] = work_in[1][0];
  m->luma[1] = work_in[1][1];
  m->luma[2] = work_in[1][2];
  m->luma[3] = 0.f;
}

/* ---- per-pixel helpers ------------------------------------------------------------------------ */
static inline void lms_to_yrg_simd(const float LMS[4], float Yrg[4])
{ /* colorspaces_inline_conversions.h:1033-1042 */
  const float Y = 0.68990272f * LMS[0] + 0.34832189f * LMS[1];
  const float a = LMS[0] + LMS[1] + LMS[2];
  const float inv_a = (a == 0.f) ? 0.f : 1.f / a;
  const float lms[4] = { LMS[0] * inv_a, LMS[1] * inv_a, LMS[2] * inv_a, 0.f };
  float rgb[4];
  mat4(LMS_to_filmlightRGB, lms, rgb);
  Yrg[0] = Y;
  Yrg[1] = rgb[0];
  Yrg[2] = rgb[1];
  Yrg[3] = 0.f;
}
static inline void yrg_to_lms_simd(const float Yrg[4], float LMS[4])
{ /* :1065-1073 */
  const float rgb[4] = { Yrg[1], Yrg[2], 1.f - Yrg[1] - Yrg[2], 0.f };
  float lms[4];
  mat4(filmlightRGB_to_LMS, rgb, lms);
  const float denom = 0.68990272f * lms[0] + 0.34832189f * lms[1];
  const float a = (denom == 0.f) ? 0.f : Yrg[0] / denom;
  LMS[0] = lms[0] * a;
  LMS[1] = lms[1] * a;
  LMS[2] = lms[2] * a;
  LMS[3] = 0.f;
}
static inline void rgb_to_ych(const float in[4], const m34 M, float Ych[4])
{ /* filmicrgb.c:1740-1760 */
  float lms[4], Yrg[4];
  mat4(M, in, lms);
  lms_to_yrg_simd(lms, Yrg);
  const float r = Yrg[1] - 0.21902143f, g = Yrg[2] - 0.54371398f;
  const float c = sqrtf(g * g + r * r); /* dt_fast_hypotf(g, r) */
  Ych[0] = Yrg[0];
  Ych[1] = c;
  Ych[2] = c != 0.f ? r / c : 1.f;
  Ych[3] = c != 0.f ? g / c : 0.f;
}
static inline void ych_to_rgb(const float in[4], const m34 M, float out[4])
{ /* :1763-1777 */
  const float Yrg[4] = { in[0], in[1] * in[2] + 0.21902143f, in[1] * in[3] + 0.54371398f, 0.f };
  float lms[4];
  yrg_to_lms_simd(Yrg, lms);
  mat4(M, lms, out);
}

static inline float spline_eval(float x, const b200_filmic_spline_t *s)
{ /* filmic_spline :1062-1160 */
  const float *M1 = s->M1, *M2 = s->M2, *M3 = s->M3, *M4 = s->M4, *M5 = s->M5;
  float result;
  if(x < s->latitude_min)
  {
    if(s->type[0] == 3)
    {
      if(M5[0] != 0.f)
        result = M3[2] + fmaxf(0.f, M3[0] * f32m_powf(fmaxf(x, 0.f), M4[0]));
      else
      {
        const float ty = s->latitude_min * M2[2] + M1[2];
        const float u = M2[2] * (x - s->latitude_min) / M1[0];
        result = M1[0] * (u / f32m_powf(1.f + f32m_powf(u, M2[0]), 1.f / M2[0])) + ty;
      }
    }
    else if(s->type[0] == 0)
      result = M1[0] + x * (M2[0] + x * (M3[0] + x * (M4[0] + x * M5[0])));
    else if(s->type[0] == 1)
      result = M1[0] + x * (M2[0] + x * (M3[0] + x * M4[0]));
    else
    {
      const float xi = s->latitude_min - x;
      const float rat = xi * (xi * M2[0] + 1.f);
      result = M4[0] - M1[0] * rat / (rat + M3[0]);
    }
  }
  else if(x > s->latitude_max)
  {
    if(s->type[1] == 3)
    {
      if(M5[1] != 0.f)
        result = M4[2] - fmaxf(0.f, M3[1] * f32m_powf(fmaxf(1.f - x, 0.f), M4[1]));
      else
      {
        const float ty = s->latitude_max * M2[2] + M1[2];
        const float u = M2[2] * (x - s->latitude_max) / M1[1];
        result = M1[1] * (u / f32m_powf(1.f + f32m_powf(u, M2[1]), 1.f / M2[1])) + ty;
      }
    }
    else if(s->type[1] == 0)
      result = M1[1] + x * (M2[1] + x * (M3[1] + x * (M4[1] + x * M5[1])));
    else if(s->type[1] == 1)
      result = M1[1] + x * (M2[1] + x * (M3[1] + x * M4[1]));
    else
    {
      const float xi = x - s->latitude_max;
      const float rat = xi * (xi * M2[1] + 1.f);
      result = M4[1] + M1[1] * rat / (rat + M3[1]);
    }
  }
  else
    result = M1[2] + x * M2[2];
  return result;
}

static inline float clip_white_raw(const float co[4], float tw, float Y, float ch, float sh)
{ /* :1826-1854 */
  const float dY = co[0] * (0.979381443298969f * ch + 0.391752577319588f * sh) + co[1] * (0.0206185567010309f * ch + 0.608247422680412f * sh)
                   - co[2] * (ch + sh);
  const float dt = tw * (0.68285981628866f * ch + 0.482137060515464f * sh);
  if(dY == 0.f) return FLT_MAX;
  const float Ya = dt / dY;
  if(Y <= Ya) return FLT_MAX;
  const float den = Y * dY - dt;
  const float num = -0.427506877216495f * (Y * (co[0] + 0.856492345150334f * co[1] + 0.554995960637719f * co[2]) - 0.988237752433297f * tw);
  return num / den;
}
static inline float clip_white(const float co[4], float tw, float Y, float ch, float sh)
{ /* :1857-1878 */
  const float eps = 1e-3f;
  const float max_Y = Y31_TO_Y06(tw);
  const float delta_Y = MAXF(max_Y - Y, 0.f);
  float mc;
  if(delta_Y < eps)
    mc = delta_Y / (eps * max_Y) * clip_white_raw(co, tw, (1.f - eps) * max_Y, ch, sh);
  else
    mc = clip_white_raw(co, tw, Y, ch, sh);
  return mc >= 0.f ? mc : FLT_MAX;
}
static inline float clip_black(const float co[4], float ch, float sh)
{ /* :1881-1901 */
  const float den = co[0] * (0.979381443298969f * ch + 0.391752577319588f * sh) + co[1] * (0.0206185567010309f * ch + 0.608247422680412f * sh)
                    - co[2] * (ch + sh);
  if(den == 0.f) return FLT_MAX;
  const float num = -0.427506877216495f * (co[0] + 0.856492345150334f * co[1] + 0.554995960637719f * co[2]);
  const float mc = num / den;
  return mc >= 0.f ? mc : FLT_MAX;
}
static inline float clip_chroma(const m34 out_m, float tw, float Y, float ch, float sh, float chroma)
{ /* :1904-1925 */
  const float w = MINF(MINF(clip_white(out_m[0], tw, Y, ch, sh), clip_white(out_m[1], tw, Y, ch, sh)), clip_white(out_m[2], tw, Y, ch, sh));
  const float b = MINF(MINF(clip_black(out_m[0], ch, sh), clip_black(out_m[1], ch, sh)), clip_black(out_m[2], ch, sh));
  return MINF(MINF(chroma, b), w);
}

static inline void gamut_check_rgb(const m34 m_out, const m34 m_in, float black, float white, const float Ych[4], float out[4])
{ /* :1949-1983 */
  float bright[4];
  ych_to_rgb(Ych, m_out, bright);
  const float min_pix = MINF(MINF(bright[0], bright[1]), bright[2]);
  const float off = MAXF(-min_pix, 0.f);
  for(int c = 0; c < 4; c++) bright[c] += off;
  float Yb[4];
  rgb_to_ych(bright, m_in, Yb);
  const float Ym = (Ych[0] + Yb[0]) / 2.f;
  const float Y = CLAMP(Ym, Y31_TO_Y06(black), Y31_TO_Y06(white));
  const float nc = clip_chroma(m_out, white, Y, Ych[2], Ych[3], Ych[1]);
  const float t[4] = { Y, nc, Ych[2], Ych[3] };
  ych_to_rgb(t, m_out, out);
  for(int c = 0; c < 4; c++) out[c] = CLAMP(out[c], 0.f, white);
}

/* gamut_mapping_simd(), :1986-2030 (filmic_desaturate_v4 :1779-1816, gamut_check_Yrg_filmic_simd :1928-1946) */
static inline void gamut_map(float Yf[4], const float Yr[4], float saturation, const filmic_mats_t *m, float black, float white, float *out)
{
  Yf[2] = Yr[2];
  Yf[3] = Yr[3];
  Yf[0] = CLAMP(Yf[0], Y31_TO_Y06(black), Y31_TO_Y06(white));
  {
    const float c_o = Yr[1] * Yr[0];
    float c_f = Yf[1] * Yf[0];
    const float delta = saturation * (c_o - c_f);
    const int brightens = (Yf[0] > Yr[0]), resat = (c_o < c_f), desat = (c_o > c_f);
    const int u_resat = (saturation > 0.f), u_desat = (saturation < 0.f);
    c_f = (brightens && resat) ? (c_o + c_f) / 2.f : (((u_resat && desat) || u_desat) ? c_f + delta : c_f);
    Yf[1] = fmaxf(c_f / Yf[0], 0.f);
  }
  {
    const float y1 = Yf[1] * Yf[2] + 0.21902143f, y2 = Yf[1] * Yf[3] + 0.54371398f;
    float max_c = Yf[1];
    if(y1 < 0.f) max_c = fminf(-0.21902143f / Yf[2], max_c);
    if(y2 < 0.f) max_c = fminf(-0.54371398f / Yf[3], max_c);
    if(y1 + y2 > 1.f) max_c = fminf((1.f - 0.21902143f - 0.54371398f) / (Yf[2] + Yf[3]), max_c);
    Yf[1] = max_c;
  }
  if(!m->use_output_profile)
  {
    gamut_check_rgb(m->output, m->input, black, white, Yf, out);
    return;
  }
  float px[4], lms[4];
  gamut_check_rgb(m->export_output, m->export_input, black, white, Yf, px);
  mat4(m->export_input, px, lms);
  mat4(m->output, lms, out);
}

static inline void agx_pixel(const float *in, float *out, const b200_filmicrgb_data_t *d, const filmic_mats_t *m, float black, float white)
{
  float pix[4] = { in[0], in[1], in[2], in[3] };
  for(int c = 0; c < 3; c++) pix[c] = (pix[c] != pix[c]) ? 0.f : CLAMPF(pix[c], -1e6f, 1e6f); /* :2536 */

  /* filmic_agx_compress_negatives :2461-2492 */
  const float *lc = m->luma;
  float cmp[4];
  {
    const float input_y = pix[0] * lc[0] + pix[1] * lc[1] + pix[2] * lc[2];
    const float max_rgb = fmaxf(fmaxf(pix[0], pix[1]), pix[2]);
    const float min_rgb = fminf(fminf(pix[0], pix[1]), pix[2]);
    float opp[4];
    for(int c = 0; c < 4; c++) opp[c] = max_rgb - pix[c];
    const float opp_y = opp[0] * lc[0] + opp[1] * lc[1] + opp[2] * lc[2];
    const float max_opp = fmaxf(fmaxf(opp[0], opp[1]), opp[2]);
    const float y_comp = max_opp - opp_y + input_y;
    const float offset = fmaxf(-min_rgb, 0.f);
    float sh[4];
    for(int c = 0; c < 4; c++) sh[c] = pix[c] + offset;
    const float max_sh = fmaxf(fmaxf(sh[0], sh[1]), sh[2]);
    float os[4];
    for(int c = 0; c < 4; c++) os[c] = max_sh - sh[c];
    const float max_os = fmaxf(fmaxf(os[0], os[1]), os[2]);
    const float y_os = os[0] * lc[0] + os[1] * lc[1] + os[2] * lc[2];
    float y_new = sh[0] * lc[0] + sh[1] * lc[1] + sh[2] * lc[2];
    y_new += max_os - y_os;
    const float ratio = (y_new > y_comp && y_new > 1e-6f) ? y_comp / y_new : 1.f;
    for(int c = 0; c < 4; c++) cmp[c] = sh[c] * ratio;
  }

  float Yo[4];
  rgb_to_ych(cmp, m->input, Yo);
  float ren[4];
  mat4(m->inset, cmp, ren);
  /* RGB_tone_mapping_v4_simd :2133-2149 */
  for(int c = 0; c < 3; c++)
  {
    const float lg = fminf(fmaxf((f32m_log2f(ren[c] / d->grey_source) - d->black_source) / d->dynamic_range, 0.0f), 1.0f);
    const float sp = spline_eval(lg, &d->spline);
    ren[c] = f32m_powf(CLAMPF(sp, 0.f, d->spline.y[4]), d->output_power);
  }
  float po[4];
  mat4(m->outset, ren, po);
  float Yf[4];
  rgb_to_ych(po, m->input, Yf);
  const float chroma_final = fminf(Yo[1], Yf[1]);
  const float beta = d->agx_beta_hue;
  const float r_mix = beta * Yo[1] * Yo[2] + (1.f - beta) * chroma_final * Yf[2];
  const float g_mix = beta * Yo[1] * Yo[3] + (1.f - beta) * chroma_final * Yf[3];
  const float norm_mix = sqrtf(g_mix * g_mix + r_mix * r_mix);
  float Yr[4] = { Yo[0], Yo[1], (norm_mix > 1e-9f) ? r_mix / norm_mix : Yo[2], (norm_mix > 1e-9f) ? g_mix / norm_mix : Yo[3] };
  Yf[1] = chroma_final;

  gamut_map(Yf, Yr, 0.f, m, black, white, out);
}

static void to_m34(m34 dst, const float src9[9])
{
  memset(dst, 0, sizeof(m34));
  for(int i = 0; i < 3; i++)
    for(int j = 0; j < 3; j++) dst[i][j] = src9[3 * i + j];
}

/* out[0..71]: input, output, export_input, export_output, inset, outset (12 floats each) */
void orc_filmic_prepare(int version, const float work_in[9], const float work_out[9], const float *export_in,
                        const float *export_out, float out[72])
{
  m34 wi, wo, ei, eo;
  to_m34(wi, work_in);
  to_m34(wo, work_out);
  if(export_in)
  {
    to_m34(ei, export_in);
    to_m34(eo, export_out);
  }
  filmic_mats_t m;
  prepare(&m, version, wi, wo, export_in ? &ei : NULL, export_in ? &eo : NULL);
  memcpy(out, m.input, 48);
  memcpy(out + 12, m.output, 48);
  memcpy(out + 24, m.export_input, 48);
  memcpy(out + 36, m.export_output, 48);
  memcpy(out + 48, m.inset, 48);
  memcpy(out + 60, m.outset, 48);
}

/* the AgX branch of process(): the tone mapping itself; with hl_deprecated == 0 the caller feeds it the frame
 * orc_filmic_reconstruct() (filmic_reconstruct_oracle.c) returns.  Returns 0, or 3 for another colour science. */
int orc_filmic_agx(const float *in, float *out, size_t width, size_t height, const b200_filmicrgb_data_t *d,
                   const float work_in[9], const float work_out[9], const float *export_in, const float *export_out)
{
  if(d->version < 5 || d->version > 9) return 3;
  m34 wi, wo, ei, eo;
  to_m34(wi, work_in);
  to_m34(wo, work_out);
  if(export_in)
  {
    to_m34(ei, export_in);
    to_m34(eo, export_out);
  }
  filmic_mats_t m;
  prepare(&m, d->version, wi, wo, export_in ? &ei : NULL, export_in ? &eo : NULL);
  const float white = f32m_powf(d->spline.y[4], d->output_power), black = f32m_powf(d->spline.y[0], d->output_power);
  const size_t n = width * height;
#pragma omp parallel for schedule(static)
  for(size_t k = 0; k < n; k++) agx_pixel(in + 4 * k, out + 4 * k, d, &m, black, white);
  return 0;
}


/* ---- the colour sciences before AgX (filmicrgb.c:2857-2887): v3 (2019) .. v7 (2023) -------------------------- */
#define NORM_MIN 1.52587890625e-05f /* math/math.h:37 */
static inline float clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); } /* clamp_simd, math/openmp_maths.h:128-131 */
static inline float log_tm(float x, const b200_filmicrgb_data_t *d)
{ /* log_tonemapping :1047-1051 */
  return clamp01((f32m_log2f(x / d->grey_source) - d->black_source) / d->dynamic_range);
}
static inline float exp_tm_v2(float x, const b200_filmicrgb_data_t *d)
{ /* :1054-1059 */
  return d->grey_source * f32m_exp2f(d->dynamic_range * x + d->black_source);
}
static inline float lum_work(const float p[4], const m34 work_in)
{ /* dt_ioppr_get_rgb_matrix_luminance, linear profile: iop_profile.h:640-655 */
  return work_in[1][0] * p[0] + work_in[1][1] * p[1] + work_in[1][2] * p[2];
}
static inline float pixel_norm(const float p[4], int variant, const m34 work_in)
{ /* get_pixel_norm_simd :976-1038 */
  switch(variant)
  {
    case 1: return fmaxf(fmaxf(p[0], p[1]), p[2]);
    case 3:
    { /* pixel_rgb_norm_power_simd :949-967 */
      float num = 0.0f, den = 0.0f;
      for(int c = 0; c < 3; c++)
      {
        const float v = fabsf(p[c]);
        const float sq = v * v;
        const float cu = sq * v;
        num += cu;
        den += sq;
      }
      return num / fmaxf(den, 1e-12f);
    }
    case 4: return sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    case 5: return sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]) * 0.5773502691896258f;
    default: return lum_work(p, work_in);
  }
}
static inline float desat_v1(float x, const b200_filmicrgb_data_t *d)
{ /* filmic_desaturate_v1 :1164-1175 */
  const float rt = x, rs = 1.0f - x;
  const float kt = f32m_expf(-0.5f * rt * rt / d->sigma_toe), ks = f32m_expf(-0.5f * rs * rs / d->sigma_shoulder);
  return 1.0f - clamp01((kt + ks) / d->saturation);
}
static inline float desat_v2(float x, const b200_filmicrgb_data_t *d)
{ /* filmic_desaturate_v2 :1178-1189 */
  const float rt = x, rs = 1.0f - x;
  const float sat2 = 0.5f / sqrtf(d->saturation);
  const float kt = f32m_expf(-rt * rt / d->sigma_toe * sat2), ks = f32m_expf(-rs * rs / d->sigma_shoulder * sat2);
  return (d->saturation - (kt + ks) * (d->saturation));
}
static inline float curve_out(float x, float lo, const b200_filmicrgb_data_t *d)
{ /* spline, clamp to [lo, y4], display transfer function */
  return f32m_powf(CLAMPF(spline_eval(x, &d->spline), lo, d->spline.y[4]), d->output_power);
}
/* filmic_split_v1 :1534-1571 and filmic_split_v2_v3 :1575-1612 differ by the desaturation only; lane 3 is not written */
static inline void split_v123(const float *in, float *out, const b200_filmicrgb_data_t *d, const m34 work_in, int v1)
{
  float temp[4];
  for(int c = 0; c < 3; c++) temp[c] = log_tm(fmaxf(in[c], NORM_MIN), d);
  const float lum = lum_work(temp, work_in);
  const float desat = v1 ? desat_v1(lum, d) : desat_v2(lum, d);
  for(int c = 0; c < 3; c++) out[c] = curve_out(lum + desat * (temp[c] - lum), d->spline.y[0], d); /* linear_saturation :1193-1196 */
}
static inline void chroma_v1(const float *in, float *out, const b200_filmicrgb_data_t *d, const m34 work_in)
{ /* :1616-1666 */
  float ratios[4];
  float norm = fmaxf(pixel_norm(in, d->preserve_color, work_in), NORM_MIN);
  for(int c = 0; c < 4; c++) ratios[c] = in[c] / norm;
  const float min_ratios = fminf(fminf(ratios[0], ratios[1]), ratios[2]);
  if(min_ratios < 0.0f)
    for(int c = 0; c < 4; c++) ratios[c] -= min_ratios;
  norm = log_tm(norm, d);
  const float desat = desat_v1(norm, d);
  for(int c = 0; c < 4; c++) ratios[c] *= norm;
  const float lum = lum_work(ratios, work_in);
  for(int c = 0; c < 3; c++) ratios[c] = (lum + desat * (ratios[c] - lum)) / norm;
  norm = curve_out(norm, d->spline.y[0], d);
  for(int c = 0; c < 4; c++) out[c] = ratios[c] * norm;
}
static inline void chroma_v2_v3(const float *in, float *out, const b200_filmicrgb_data_t *d, const m34 work_in)
{ /* :1670-1737 */
  float norm = fmaxf(pixel_norm(in, d->preserve_color, work_in), NORM_MIN);
  float ratios[4];
  for(int c = 0; c < 4; c++) ratios[c] = in[c] / norm;
  const float min_ratios = fminf(fminf(ratios[0], ratios[1]), ratios[2]);
  if(min_ratios < 0.0f)
    for(int c = 0; c < 4; c++) ratios[c] -= min_ratios;
  norm = log_tm(norm, d);
  const float desat = desat_v2(norm, d);
  norm = curve_out(norm, d->spline.y[0], d);
  for(int c = 0; c < 3; c++) ratios[c] = fmaxf(ratios[c] + (1.0f - ratios[c]) * (1.0f - desat), 0.0f);
  if(d->version == 2) norm /= fmaxf(pixel_norm(ratios, d->preserve_color, work_in), NORM_MIN);
  for(int c = 0; c < 4; c++) out[c] = ratios[c] * norm;
  const float max_pix = fmaxf(fmaxf(out[0], out[1]), out[2]);
  if(max_pix > 1.0f)
    for(int c = 0; c < 4; c++)
    {
      ratios[c] = fmaxf(ratios[c] + (1.0f - max_pix), 0.0f);
      out[c] = ratios[c] * norm;
    }
}
static inline void norm_tm_v4(const float *in, int variant, const b200_filmicrgb_data_t *d, const m34 work_in, float nmin, float nmax, float *o)
{ /* norm_tone_mapping_v4_simd :2106-2131 */
  float norm = CLAMPF(pixel_norm(in, variant, work_in), nmin, nmax);
  float ratios[4];
  for(int c = 0; c < 4; c++) ratios[c] = in[c] / norm;
  norm = log_tm(norm, d);
  norm = curve_out(norm, d->spline.y[0], d);
  for(int c = 0; c < 4; c++) o[c] = ratios[c] * norm;
}
static inline void rgb_tm_v4(const float *in, const b200_filmicrgb_data_t *d, float *o)
{ /* RGB_tone_mapping_v4_simd :2133-2149 */
  for(int c = 0; c < 3; c++) o[c] = curve_out(log_tm(in[c], d), 0.f, d);
  o[3] = in[3];
}
static inline void v4_v5_pixel(const float *in, float *out, const b200_filmicrgb_data_t *d, const filmic_mats_t *m, const m34 work_in, float nmin,
                               float nmax, float black, float white)
{ /* filmic_chroma_v4 :2153-2198, filmic_split_v4 :2201-2243, filmic_v5 :2247-2299 */
  float po[4], Yo[4], Yf[4];
  float saturation = d->saturation;
  int clamp_chroma = 0;
  if(d->version == 4)
  {
    float naive[4], maxrgb[4];
    rgb_tm_v4(in, d, naive);
    norm_tm_v4(in, 1, d, work_in, nmin, nmax, maxrgb);
    for(int c = 0; c < 4; c++) po[c] = (0.5f + d->saturation) * maxrgb[c];
    for(int c = 0; c < 4; c++) po[c] = (0.5f - d->saturation) * naive[c] + po[c];
    saturation = 0.f;
    clamp_chroma = 1;
  }
  else if(d->preserve_color == 0)
  {
    rgb_tm_v4(in, d, po);
    clamp_chroma = 1;
  }
  else
    norm_tm_v4(in, d->preserve_color, d, work_in, nmin, nmax, po);
  rgb_to_ych(in, m->input, Yo);
  rgb_to_ych(po, m->input, Yf);
  if(clamp_chroma) Yf[1] = fminf(Yo[1], Yf[1]);
  gamut_map(Yf, Yo, saturation, m, black, white, out);
}

/* process() for version < 5 (reconstruction at its deprecation sentinel).  Lane 3 of split v1..v3 is left as found.
 * Returns 0, or 3 for what is not restated. */
int orc_filmic_legacy(const float *in, float *out, size_t width, size_t height, const b200_filmicrgb_data_t *d, const float work_in[9],
                      const float work_out[9], const float *export_in, const float *export_out)
{
  if(d->version < 0 || d->version > 4) return 3;
  m34 wi, wo, ei, eo;
  to_m34(wi, work_in);
  to_m34(wo, work_out);
  if(export_in)
  {
    to_m34(ei, export_in);
    to_m34(eo, export_out);
  }
  filmic_mats_t m;
  prepare(&m, d->version, wi, wo, export_in ? &ei : NULL, export_in ? &eo : NULL);
  const float white = f32m_powf(d->spline.y[4], d->output_power), black = f32m_powf(d->spline.y[0], d->output_power);
  const float nmin = exp_tm_v2(0.f, d), nmax = exp_tm_v2(1.f, d);
  const size_t n = width * height;
#pragma omp parallel for schedule(static)
  for(size_t k = 0; k < n; k++)
  {
    const float *pi = in + 4 * k;
    float *po = out + 4 * k;
    if(d->version >= 3)
      v4_v5_pixel(pi, po, d, &m, wi, nmin, nmax, black, white);
    else if(d->preserve_color == 0)
      split_v123(pi, po, d, wi, d->version == 0);
    else if(d->version == 0)
      chroma_v1(pi, po, d, wi);
    else
      chroma_v2_v3(pi, po, d, wi);
  }
  return 0;
}
