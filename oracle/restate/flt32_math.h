/* Single-precision libm as the reference's CPU path gets it: glibc 2.39 (Ubuntu 24.04),
 * sysdeps/ieee754/flt-32/{e_powf,e_log2f,e_logf,e_expf,e_exp2f}.c -- the "optimized routines"
 * algorithms (double-precision core, 16-entry log tables, 32-entry exp2 table).
 * TEST INFRASTRUCTURE ONLY (the product has its own device copy in ansel_b200/csrc/flt32_math.cuh).
 *
 * glibc is a third-party dependency that is NOT under /root/reference; the reference reaches it
 * at colorprofiles/iop_profile.h:561 (powf), iop/denoiseprofile.c:938,1020,1041,1086 (powf),
 * iop/filmicrgb.c:1050 (log2f) :1089,1096,1125,1132,2124,2143 (powf), pixel/locallaplacian.c:323
 * (expf).  The published algorithm is restated here; the table constants are the published ones
 * and were cross-checked against the .rodata of this image's /lib/x86_64-linux-gnu/libm.so.6.
 * tests/test_flt32_math.py pins every function bit-for-bit against the system libm.
 *
 * On x86-64 glibc selects its FMA build of these files by ifunc whenever the CPU has FMA
 * (sysdeps/x86_64/fpu/multiarch), in which GCC contracts every `a*b + c` below into one fused
 * operation.  The expressions are single-multiply-feeds-one-add, so the contraction is
 * unambiguous; it is written out with fma() here.  Define FLT32_NO_FMA for the SSE2 build.
 */
#ifndef B200_ORACLE_FLT32_MATH_H
#define B200_ORACLE_FLT32_MATH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef FLT32_NO_FMA
#define F32M_FMA(a, b, c) ((a) * (b) + (c))
#else
#define F32M_FMA(a, b, c) fma((a), (b), (c))
#endif

static inline uint32_t f32m_asuint(float f)
{
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static inline float f32m_asfloat(uint32_t u)
{
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline uint64_t f32m_asuint64(double d)
{
  uint64_t u;
  memcpy(&u, &d, 8);
  return u;
}
static inline double f32m_asdouble(uint64_t u)
{
  double d;
  memcpy(&d, &u, 8);
  return d;
}

/* __exp2f_data (e_exp2f_data.c): tab[i] = bits(2^(i/32)) - (i << 47) */
static const uint64_t f32m_exp2_tab[32] = {
  0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL,
  0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL,
  0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL,
  0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL,
  0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL,
  0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL,
  0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL,
  0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL
};
#define F32M_EXP2_C0 0x1.c6af84b912394p-5
#define F32M_EXP2_C1 0x1.ebfce50fac4f3p-3
#define F32M_EXP2_C2 0x1.62e42ff0c52d6p-1
#define F32M_EXP2_SHIFT_SCALED 0x1.8p+47 /* 0x1.8p52 / 32 */
#define F32M_EXP_SHIFT 0x1.8p+52
#define F32M_INVLN2_SCALED 0x1.71547652b82fep+5 /* 32 / ln 2 */
#define F32M_EXP_C0S 0x1.c6af84b912394p-20 /* C0 / 32^3 */
#define F32M_EXP_C1S 0x1.ebfce50fac4f3p-13 /* C1 / 32^2 */
#define F32M_EXP_C2S 0x1.62e42ff0c52d6p-6  /* C2 / 32 */

/* 1/c of the 16 sub-intervals of [0x3f330000, 2*0x3f330000): shared by logf, log2f and powf */
static const double f32m_invc[16] = {
  0x1.661ec79f8f3bep+0, 0x1.571ed4aaf883dp+0, 0x1.49539f0f010b0p+0, 0x1.3c995b0b80385p+0,
  0x1.30d190c8864a5p+0, 0x1.25e227b0b8ea0p+0, 0x1.1bb4a4a1a343fp+0, 0x1.12358f08ae5bap+0,
  0x1.0953f419900a7p+0, 0x1.0000000000000p+0, 0x1.e608cfd9a47acp-1, 0x1.ca4b31f026aa0p-1,
  0x1.b2036576afce6p-1, 0x1.9c2d163a1aa2dp-1, 0x1.886e6037841edp-1, 0x1.767dcf5534862p-1
};
/* ln(c): __logf_data (e_logf_data.c) */
static const double f32m_lnc[16] = {
  -0x1.57bf7808caadep-2, -0x1.2bef0a7c06ddbp-2, -0x1.01eae7f513a67p-2, -0x1.b31d8a68224e9p-3,
  -0x1.6574f0ac07758p-3, -0x1.1aa2bc79c8100p-3, -0x1.a4e76ce8c0e5ep-4, -0x1.1973c5a611cccp-4,
  -0x1.252f438e10c1ep-5, 0x0.0p+0,              0x1.aa5aa5df25984p-5,  0x1.c5e53aa362eb4p-4,
  0x1.526e57720db08p-3,  0x1.bc2860d224770p-3,  0x1.1058bc8a07ee1p-2,  0x1.4043057b6ee09p-2
};
/* log2(c): __log2f_data and __powf_log2_data (e_log2f_data.c, e_powf_log2_data.c; POWF_SCALE = 1) */
static const double f32m_log2c[16] = {
  -0x1.efec65b963019p-2, -0x1.b0b6832d4fca4p-2, -0x1.7418b0a1fb77bp-2, -0x1.39de91a6dcf7bp-2,
  -0x1.01d9bf3f2b631p-2, -0x1.97c1d1b3b7af0p-3, -0x1.2f9e393af3c9fp-3, -0x1.960cbbf788d5cp-4,
  -0x1.a6f9db6475fcep-5, 0x0.0p+0,              0x1.338ca9f24f53dp-4,  0x1.476a9543891bap-3,
  0x1.e840b4ac4e4d2p-3,  0x1.40645f0c6651cp-2,  0x1.88e9c2c1b9ff8p-2,  0x1.ce0a44eb17bccp-2
};
#define F32M_LN2 0x1.62e42fefa39efp-1
#define F32M_LOGF_A0 -0x1.00ea348b88334p-2
#define F32M_LOGF_A1 0x1.5575b0be00b6ap-2
#define F32M_LOGF_A2 -0x1.ffffef20a4123p-2
#define F32M_LOG2F_A0 -0x1.712b6f70a7e4dp-2
#define F32M_LOG2F_A1 0x1.ecabf496832e0p-2
#define F32M_LOG2F_A2 -0x1.715479ffae3dep-1
#define F32M_LOG2F_A3 0x1.715475f35c8b8p+0
#define F32M_POWF_A0 0x1.27616c9496e0bp-2
#define F32M_POWF_A1 -0x1.71969a075c67ap-2
#define F32M_POWF_A2 0x1.ec70a6ca7baddp-2
#define F32M_POWF_A3 -0x1.7154748bef6c8p-1
#define F32M_POWF_A4 0x1.71547652ab82bp+0
#define F32M_OFF 0x3f330000u

static inline uint32_t f32m_top12(float x) { return f32m_asuint(x) >> 20; }

/* e_expf.c */
static inline float f32m_expf(float x)
{
  const double xd = (double)x;
  const uint32_t abstop = f32m_top12(x) & 0x7ff;
  if(abstop >= f32m_top12(88.0f))
  {
    if(f32m_asuint(x) == f32m_asuint(-INFINITY)) return 0.0f;
    if(abstop >= f32m_top12(INFINITY)) return x + x;
    if(x > 0x1.62e42ep6f) return INFINITY;
    if(x < -0x1.9fe368p6f) return 0.0f;
  }
  double z = F32M_INVLN2_SCALED * xd;
  double kd = z + F32M_EXP_SHIFT;
  const uint64_t ki = f32m_asuint64(kd);
  kd -= F32M_EXP_SHIFT;
  const double r = z - kd;
  uint64_t t = f32m_exp2_tab[ki % 32];
  t += ki << (52 - 5);
  const double s = f32m_asdouble(t);
  z = F32M_FMA(F32M_EXP_C0S, r, F32M_EXP_C1S);
  const double r2 = r * r;
  double y = F32M_FMA(F32M_EXP_C2S, r, 1.0);
  y = F32M_FMA(z, r2, y);
  y = y * s;
  return (float)y;
}

/* e_exp2f.c */
static inline float f32m_exp2f(float x)
{
  const double xd = (double)x;
  const uint32_t abstop = f32m_top12(x) & 0x7ff;
  if(abstop >= f32m_top12(128.0f))
  {
    if(f32m_asuint(x) == f32m_asuint(-INFINITY)) return 0.0f;
    if(abstop >= f32m_top12(INFINITY)) return x + x;
    if(x > 0.0f) return INFINITY;
    if(x <= -150.0f) return 0.0f;
  }
  double kd = xd + F32M_EXP2_SHIFT_SCALED;
  const uint64_t ki = f32m_asuint64(kd);
  kd -= F32M_EXP2_SHIFT_SCALED;
  const double r = xd - kd;
  uint64_t t = f32m_exp2_tab[ki % 32];
  t += ki << (52 - 5);
  const double s = f32m_asdouble(t);
  const double z = F32M_FMA(F32M_EXP2_C0, r, F32M_EXP2_C1);
  const double r2 = r * r;
  double y = F32M_FMA(F32M_EXP2_C2, r, 1.0);
  y = F32M_FMA(z, r2, y);
  y = y * s;
  return (float)y;
}

/* e_logf.c */
static inline float f32m_logf(float x)
{
  uint32_t ix = f32m_asuint(x);
  if(ix == 0x3f800000u) return 0.0f;
  if(ix - 0x00800000u >= 0x7f800000u - 0x00800000u)
  {
    if(ix * 2 == 0) return -INFINITY;
    if(ix == 0x7f800000u) return x;
    if((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return (x - x) / 0.0f;
    ix = f32m_asuint(x * 0x1p23f);
    ix -= 23u << 23;
  }
  const uint32_t tmp = ix - F32M_OFF;
  const int i = (tmp >> (23 - 4)) % 16;
  const int k = (int32_t)tmp >> 23;
  const uint32_t iz = ix - (tmp & (0x1ffu << 23));
  const double z = (double)f32m_asfloat(iz);
  const double r = F32M_FMA(z, f32m_invc[i], -1.0);
  const double y0 = F32M_FMA((double)k, F32M_LN2, f32m_lnc[i]);
  const double r2 = r * r;
  double y = F32M_FMA(F32M_LOGF_A1, r, F32M_LOGF_A2);
  y = F32M_FMA(F32M_LOGF_A0, r2, y);
  y = F32M_FMA(y, r2, y0 + r);
  return (float)y;
}

/* e_log2f.c */
static inline float f32m_log2f(float x)
{
  uint32_t ix = f32m_asuint(x);
  if(ix == 0x3f800000u) return 0.0f;
  if(ix - 0x00800000u >= 0x7f800000u - 0x00800000u)
  {
    if(ix * 2 == 0) return -INFINITY;
    if(ix == 0x7f800000u) return x;
    if((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return (x - x) / 0.0f;
    ix = f32m_asuint(x * 0x1p23f);
    ix -= 23u << 23;
  }
  const uint32_t tmp = ix - F32M_OFF;
  const int i = (tmp >> (23 - 4)) % 16;
  const uint32_t top = tmp & 0xff800000u;
  const uint32_t iz = ix - top;
  const int k = (int32_t)tmp >> 23;
  const double z = (double)f32m_asfloat(iz);
  const double r = F32M_FMA(z, f32m_invc[i], -1.0);
  const double y0 = f32m_log2c[i] + (double)k;
  const double r2 = r * r;
  double y = F32M_FMA(F32M_LOG2F_A1, r, F32M_LOG2F_A2);
  y = F32M_FMA(F32M_LOG2F_A0, r2, y);
  const double p = F32M_FMA(F32M_LOG2F_A3, r, y0);
  y = F32M_FMA(y, r2, p);
  return (float)y;
}

/* e_powf.c: log2_inline + exp2_inline.  checkint(): 0 = not an integer, 1 = odd, 2 = even */
static inline int f32m_checkint(uint32_t iy)
{
  const int e = iy >> 23 & 0xff;
  if(e < 0x7f) return 0;
  if(e > 0x7f + 23) return 2;
  if(iy & ((1u << (0x7f + 23 - e)) - 1)) return 0;
  if(iy & (1u << (0x7f + 23 - e))) return 1;
  return 2;
}
static inline int f32m_zeroinfnan(uint32_t ix) { return 2 * ix - 1 >= 2u * 0x7f800000u - 1; }

static inline float f32m_powf(float x, float y)
{
  uint32_t sign_bias = 0;
  uint32_t ix = f32m_asuint(x);
  const uint32_t iy = f32m_asuint(y);
  if(ix - 0x00800000u >= 0x7f800000u - 0x00800000u || f32m_zeroinfnan(iy))
  {
    if(f32m_zeroinfnan(iy))
    {
      if(2 * iy == 0) return 1.0f;
      if(ix == 0x3f800000u) return 1.0f;
      if(2 * ix > 2u * 0x7f800000u || 2 * iy > 2u * 0x7f800000u) return x + y;
      if(2 * ix == 2 * 0x3f800000u) return 1.0f;
      if((2 * ix < 2 * 0x3f800000u) == !(iy & 0x80000000u)) return 0.0f;
      return y * y;
    }
    if(f32m_zeroinfnan(ix))
    {
      float x2 = x * x;
      if((ix & 0x80000000u) && f32m_checkint(iy) == 1)
      {
        x2 = -x2;
        sign_bias = 1;
      }
      if(2 * ix == 0 && (iy & 0x80000000u)) return sign_bias ? -INFINITY : INFINITY;
      return (iy & 0x80000000u) ? 1 / x2 : x2;
    }
    if(ix & 0x80000000u)
    {
      const int yint = f32m_checkint(iy);
      if(yint == 0) return (x - x) / 0.0f;
      if(yint == 1) sign_bias = 1u << (5 + 11);
      ix &= 0x7fffffffu;
    }
    if(ix < 0x00800000u)
    {
      ix = f32m_asuint(x * 0x1p23f);
      ix &= 0x7fffffffu;
      ix -= 23u << 23;
    }
  }
  /* log2_inline */
  const uint32_t tmp = ix - F32M_OFF;
  const int i = (tmp >> (23 - 4)) % 16;
  const uint32_t top = tmp & 0xff800000u;
  const uint32_t iz = ix - top;
  const int k = (int32_t)top >> 23;
  const double z = (double)f32m_asfloat(iz);
  const double r = F32M_FMA(z, f32m_invc[i], -1.0);
  const double y0 = f32m_log2c[i] + (double)k;
  const double r2 = r * r;
  double yy = F32M_FMA(F32M_POWF_A0, r, F32M_POWF_A1);
  const double p = F32M_FMA(F32M_POWF_A2, r, F32M_POWF_A3);
  const double r4 = r2 * r2;
  double q = F32M_FMA(F32M_POWF_A4, r, y0);
  q = F32M_FMA(p, r2, q);
  yy = F32M_FMA(yy, r4, q);
  const double ylogx = (double)y * yy;
  if((f32m_asuint64(ylogx) >> 47 & 0xffff) >= f32m_asuint64(126.0) >> 47)
  {
    if(ylogx > 0x1.fffffffd1d571p+6) return sign_bias ? -INFINITY : INFINITY;
    if(ylogx <= -150.0) return sign_bias ? -0.0f : 0.0f;
  }
  /* exp2_inline */
  double kd = ylogx + F32M_EXP2_SHIFT_SCALED;
  const uint64_t ki = f32m_asuint64(kd);
  kd -= F32M_EXP2_SHIFT_SCALED;
  const double rr = ylogx - kd;
  uint64_t t = f32m_exp2_tab[ki % 32];
  const uint64_t ski = ki + sign_bias;
  t += ski << (52 - 5);
  const double s = f32m_asdouble(t);
  const double zz = F32M_FMA(F32M_EXP2_C0, rr, F32M_EXP2_C1);
  const double rr2 = rr * rr;
  double e = F32M_FMA(F32M_EXP2_C2, rr, 1.0);
  e = F32M_FMA(zz, rr2, e);
  e = e * s;
  return (float)e;
}

/* ---- sinf / cosf: sysdeps/ieee754/flt-32/{s_sinf.c,s_cosf.c,sincosf.h,s_sincosf_data.c} ---------------------
 * Reached by the reference through iop/noise_generator.h:93-96 (Box-Muller), with arguments 2*pi*u, u in [0,1).
 * Restated for |x| < 120 (the pi/4 polynomial and the single multiply-subtract reduction); the table-driven
 * reduction of larger arguments is not needed on this path and returns NaN here so a misuse cannot go unnoticed.
 * __sincosf_table cross-checked against the .rodata of this image's libm.so.6 (x86-64: TOINT_INTRINSICS 0). */
#define F32M_SC_HPI_INV 0x1.45F306DC9C883p+23
#define F32M_SC_HPI 0x1.921FB54442D18p0
#define F32M_SC_C1 -0x1.ffffffd0c621cp-2
#define F32M_SC_C2 0x1.55553e1068f19p-5
#define F32M_SC_C3 -0x1.6c087e89a359dp-10
#define F32M_SC_C4 0x1.99343027bf8c3p-16
#define F32M_SC_S1 -0x1.555545995a603p-3
#define F32M_SC_S2 0x1.1107605230bc4p-7
#define F32M_SC_S3 -0x1.994eb3774cf24p-13

static inline uint32_t f32m_abstop12(float x) { return (f32m_asuint(x) >> 20) & 0x7ff; }

/* sinf_poly(): sine polynomial for even n, cosine for odd; `neg` selects __sincosf_table[1] (cosine coefficients negated) */
static inline float f32m_sinf_poly(double x, double x2, int neg, int n)
{
  if((n & 1) == 0)
  {
    const double x3 = x * x2;
    const double s1 = F32M_FMA(x2, F32M_SC_S3, F32M_SC_S2);
    const double x7 = x3 * x2;
    const double s = F32M_FMA(x3, F32M_SC_S1, x);
    return (float)F32M_FMA(x7, s1, s);
  }
  const double sg = neg ? -1.0 : 1.0;
  const double x4 = x2 * x2;
  const double c2 = F32M_FMA(x2, sg * F32M_SC_C4, sg * F32M_SC_C3);
  const double c1 = F32M_FMA(x2, sg * F32M_SC_C1, sg * 1.0);
  const double x6 = x4 * x2;
  const double c = F32M_FMA(x4, sg * F32M_SC_C2, c1);
  return (float)F32M_FMA(x6, c2, c);
}
/* reduce_fast(): x - n*pi/2 with n = round(x * 2/pi), the quadrant */
static inline double f32m_reduce_fast(double x, int *np)
{
  const double r = x * F32M_SC_HPI_INV;
  const int n = ((int32_t)r + 0x800000) >> 24;
  *np = n;
  return F32M_FMA(-(double)n, F32M_SC_HPI, x);
}
static const double f32m_sc_sign[4] = { 1.0, -1.0, -1.0, 1.0 };

static inline float f32m_sinf(float y)
{
  double x = y;
  if(f32m_abstop12(y) < f32m_abstop12(0x1.921FB6p-1f))
  {
    if(f32m_abstop12(y) < f32m_abstop12(0x1p-12f)) return y;
    return f32m_sinf_poly(x, x * x, 0, 0);
  }
  if(f32m_abstop12(y) < f32m_abstop12(120.0f))
  {
    int n;
    x = f32m_reduce_fast(x, &n);
    const double s = f32m_sc_sign[n & 3];
    return f32m_sinf_poly(x * s, x * x, (n & 2) != 0, n);
  }
  return NAN;
}
static inline float f32m_cosf(float y)
{
  double x = y;
  if(f32m_abstop12(y) < f32m_abstop12(0x1.921FB6p-1f))
  {
    if(f32m_abstop12(y) < f32m_abstop12(0x1p-12f)) return 1.0f;
    return f32m_sinf_poly(x, x * x, 0, 1);
  }
  if(f32m_abstop12(y) < f32m_abstop12(120.0f))
  {
    int n;
    x = f32m_reduce_fast(x, &n);
    const double s = f32m_sc_sign[n & 3];
    return f32m_sinf_poly(x * s, x * x, (n & 2) != 0, n ^ 1);
  }
  return NAN;
}
/* ---- atanf, atan2f, hypotf: glibc 2.39 sysdeps/ieee754/flt-32/{s_atanf,e_atan2f,e_hypotf}.c ------------------------------------
 * atanf and atan2f are the fdlibm float routines (argument reduction to four breakpoints, an 11-term odd polynomial split in two,
 * head / tail constants of the breakpoints); they have no FMA build, every operation is a float operation in source order.
 * hypotf is one square root of the exact double sum of the squares.  The reference reaches them at
 * common/colorspaces_inline_conversions.h:594-606 (dt_Lab_2_LCH) and :775-781 (dt_JzAzBz_2_JzCzhz).
 * Pinned by tests/test_cpu_flt32_math.py; a standalone sweep of all 2^32 arguments of atanf and of 1.6e9 argument pairs of the other
 * two found no difference from this image's libm. */
static const float f32m_atanhi[4] = { 4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f };
static const float f32m_atanlo[4] = { 5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f };
static const float f32m_aT[11] = { 3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f, 9.0908870101e-02f, -7.6918758452e-02f,
                                   6.6610731184e-02f, -5.8335702866e-02f, 4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f };
static inline float f32m_atanf(float x)
{
  const int32_t hx = (int32_t)f32m_asuint(x), ix = hx & 0x7fffffff;
  int id;
  if(ix >= 0x4c000000)
  { /* |x| >= 2^25 */
    if(ix > 0x7f800000) return x + x;
    return hx > 0 ? f32m_atanhi[3] + f32m_atanlo[3] : -f32m_atanhi[3] - f32m_atanlo[3];
  }
  if(ix < 0x3ee00000)
  { /* |x| < 0.4375 */
    if(ix < 0x31000000) return x; /* |x| < 2^-29 */
    id = -1;
  }
  else
  {
    x = fabsf(x);
    if(ix < 0x3f980000)
    {
      if(ix < 0x3f300000)
      {
        id = 0;
        x = (2.0f * x - 1.0f) / (2.0f + x);
      }
      else
      {
        id = 1;
        x = (x - 1.0f) / (x + 1.0f);
      }
    }
    else if(ix < 0x401c0000)
    {
      id = 2;
      x = (x - 1.5f) / (1.0f + 1.5f * x);
    }
    else
    {
      id = 3;
      x = -1.0f / x;
    }
  }
  const float z = x * x, w = z * z;
  const float *aT = f32m_aT;
  const float s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
  const float s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
  if(id < 0) return x - x * (s1 + s2);
  const float r = f32m_atanhi[id] - ((x * (s1 + s2) - f32m_atanlo[id]) - x);
  return hx < 0 ? -r : r;
}
static inline float f32m_atan2f(float y, float x)
{
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
  const int32_t hx = (int32_t)f32m_asuint(x), ix = hx & 0x7fffffff, hy = (int32_t)f32m_asuint(y), iy = hy & 0x7fffffff;
  if(ix > 0x7f800000 || iy > 0x7f800000) return x + y;
  if(hx == 0x3f800000) return f32m_atanf(y);
  const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2); /* 2 * sign(x) + sign(y) */
  if(iy == 0) return m < 2 ? y : (m == 2 ? pi + tiny : -pi - tiny);
  if(ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
  if(ix == 0x7f800000)
  {
    if(iy == 0x7f800000) return m == 0 ? pi_o_4 + tiny : (m == 1 ? -pi_o_4 - tiny : (m == 2 ? 3.0f * pi_o_4 + tiny : -3.0f * pi_o_4 - tiny));
    return m == 0 ? 0.0f : (m == 1 ? -0.0f : (m == 2 ? pi + tiny : -pi - tiny));
  }
  if(iy == 0x7f800000) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
  const int k = (iy - ix) >> 23;
  float z;
  if(k > 60)
    z = pi_o_2 + 0.5f * pi_lo;
  else if(hx < 0 && k < -60)
    z = 0.0f;
  else
    z = f32m_atanf(fabsf(y / x));
  switch(m)
  {
    case 0: return z;
    case 1: return f32m_asfloat(f32m_asuint(z) ^ 0x80000000u);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
  }
}
static inline float f32m_hypotf(float x, float y)
{
  if(!isfinite(x) || !isfinite(y)) return (isinf(x) || isinf(y)) ? INFINITY : x + y;
  return (float)sqrt((double)x * (double)x + (double)y * (double)y);
}
#endif
