// The handful of 80-bit operations the reference's LCh highlight reconstruction performs (iop/highlights/lch.c:367-401): its
// SQRT3 and SQRT12 are `long double` literals (iop/highlights/common.h:618-619), so on x86-64 `SQRT3 * (R - G)` and
// `L - H / 6.0f + C / SQRT12` are x87 operations -- a product, a quotient and a sum rounded to a 64-bit significand, then a
// second rounding to float on assignment.  The device has no such type: these are the operations in integer arithmetic,
// bit-exact (round to nearest even at 64 bits, then at 24; gradual underflow and overflow of the final float as the x87 store
// does them).  tests/test_cpu_pipe_ends_emulation.py compares them with the host's own long double on millions of operands.
#pragma once
#include <stdint.h>

namespace x87
{
struct ext
{ // (-1)^sign * mant * 2^(exp - 63), mant in [2^63, 2^64), or mant == 0 for zero
  uint64_t mant;
  int exp;
  int sign;
};
constexpr uint64_t SQRT3_MANT = 0xddb3d742c265539eULL; // 1.7320508075688772935274463415058723669L = mant * 2^(0 - 63)
constexpr int SQRT3_EXP = 0, SQRT12_EXP = 1;           // SQRT12 has the same significand, one binade up

__device__ __forceinline__ int clz64(uint64_t v)
{
#ifdef B200_KERNELS_ON_CPU
  return __builtin_clzll(v);
#else
  return __clzll((long long)v);
#endif
}
__device__ __forceinline__ bool finite_f32(float f) { return (__float_as_uint(f) & 0x7f800000u) != 0x7f800000u; }

// exact
__device__ __forceinline__ ext from_float(float f)
{
  const uint32_t u = __float_as_uint(f);
  ext r;
  r.sign = (int)(u >> 31);
  const uint32_t E = (u >> 23) & 0xffu, F = u & 0x7fffffu;
  if(E == 0)
  {
    if(F == 0)
    {
      r.mant = 0;
      r.exp = 0;
      return r;
    }
    const int lz = clz64((uint64_t)F);
    r.mant = (uint64_t)F << lz;
    r.exp = -86 - lz;
    return r;
  }
  r.mant = (uint64_t)(0x800000u | F) << 40;
  r.exp = (int)E - 127;
  return r;
}
// round-to-nearest-even increment of a 64-bit significand; a carry out renormalises
__device__ __forceinline__ void round64(ext &r, bool guard, bool sticky)
{
  if(guard && (sticky || (r.mant & 1)))
  {
    r.mant++;
    if(r.mant == 0)
    {
      r.mant = 0x8000000000000000ULL;
      r.exp++;
    }
  }
}
// a * (cm * 2^(ce - 63)), rounded to 64 bits
__device__ __forceinline__ ext mul_const(ext a, uint64_t cm, int ce)
{
  if(a.mant == 0) return a;
  const unsigned __int128 P = (unsigned __int128)a.mant * cm;
  ext r;
  r.sign = a.sign;
  bool guard, sticky;
  if(P >> 127)
  {
    r.mant = (uint64_t)(P >> 64);
    const uint64_t rem = (uint64_t)P;
    guard = rem >> 63;
    sticky = (rem << 1) != 0;
    r.exp = a.exp + ce + 1;
  }
  else
  {
    r.mant = (uint64_t)(P >> 63);
    const uint64_t rem = (uint64_t)P & 0x7fffffffffffffffULL;
    guard = (rem >> 62) & 1;
    sticky = (rem & 0x3fffffffffffffffULL) != 0;
    r.exp = a.exp + ce;
  }
  round64(r, guard, sticky);
  return r;
}
// a / (cm * 2^(ce - 63)), rounded to 64 bits
__device__ __forceinline__ ext div_const(ext a, uint64_t cm, int ce)
{
  if(a.mant == 0) return a;
  ext r;
  r.sign = a.sign;
  unsigned __int128 N;
  if(a.mant >= cm)
  {
    N = (unsigned __int128)a.mant << 63;
    r.exp = a.exp - ce;
  }
  else
  {
    N = (unsigned __int128)a.mant << 64;
    r.exp = a.exp - ce - 1;
  }
  r.mant = (uint64_t)(N / cm);
  const unsigned __int128 rem2 = (N % cm) << 1; // the next quotient bit and what is left after it
  const bool guard = rem2 >= cm;
  const bool sticky = (guard ? rem2 - cm : rem2) != 0;
  round64(r, guard, sticky);
  return r;
}
// x + y, rounded to 64 bits
__device__ __forceinline__ ext add(ext x, ext y)
{
  if(x.mant == 0)
  {
    if(y.mant == 0) y.sign = x.sign & y.sign; // (+0) + (-0) = +0, (-0) + (-0) = -0
    return y;
  }
  if(y.mant == 0) return x;
  if(y.exp > x.exp || (y.exp == x.exp && y.mant > x.mant))
  {
    const ext t = x;
    x = y;
    y = t;
  }
  // significands at bit 125 of a 128-bit word: two bits of headroom, 62 below for the aligned operand
  unsigned __int128 X = (unsigned __int128)x.mant << 62, Y = (unsigned __int128)y.mant << 62;
  const int d = x.exp - y.exp;
  if(d >= 127)
    Y = 1; // far below: only its presence matters
  else if(d > 0)
  {
    const bool lost = (Y & ((((unsigned __int128)1) << d) - 1)) != 0;
    Y = (Y >> d) | (lost ? 1 : 0);
  }
  const unsigned __int128 S = x.sign == y.sign ? X + Y : X - Y;
  ext r;
  r.sign = x.sign;
  if(S == 0)
  {
    r.mant = 0;
    r.exp = 0;
    r.sign = 0; // exact cancellation gives +0 under round-to-nearest
    return r;
  }
  const uint64_t hi = (uint64_t)(S >> 64), lo = (uint64_t)S;
  const int p = hi ? 127 - clz64(hi) : 63 - clz64(lo); // position of the leading one
  r.exp = x.exp + p - 125;
  if(p <= 63)
  {
    r.mant = lo << (63 - p); // no bits below: exact
    return r;
  }
  const int sh = p - 63;
  r.mant = (uint64_t)(S >> sh);
  const bool guard = (S >> (sh - 1)) & 1;
  const bool sticky = sh > 1 ? (S & ((((unsigned __int128)1) << (sh - 1)) - 1)) != 0 : false;
  round64(r, guard, sticky);
  return r;
}
// the x87 store to a float: round to nearest even at 24 bits, gradual underflow, overflow to infinity
__device__ __forceinline__ float to_float(ext a)
{
  const uint32_t s = (uint32_t)a.sign << 31;
  if(a.mant == 0) return __uint_as_float(s);
  int e = a.exp;
  if(e >= -126)
  {
    uint32_t m = (uint32_t)(a.mant >> 40);
    const bool guard = (a.mant >> 39) & 1, sticky = (a.mant & 0x7fffffffffULL) != 0;
    if(guard && (sticky || (m & 1)))
    {
      m++;
      if(m == 0x1000000u)
      {
        m = 0x800000u;
        e++;
      }
    }
    if(e > 127) return __uint_as_float(s | 0x7f800000u);
    return __uint_as_float(s | ((uint32_t)(e + 127) << 23) | (m & 0x7fffffu));
  }
  const int sh = 40 + (-126 - e); // > 40
  uint32_t m;
  bool guard, sticky;
  if(sh > 64)
  {
    m = 0;
    guard = false;
    sticky = true;
  }
  else if(sh == 64)
  {
    m = 0;
    guard = true; // the leading one
    sticky = (a.mant << 1) != 0;
  }
  else
  {
    m = (uint32_t)(a.mant >> sh);
    guard = (a.mant >> (sh - 1)) & 1;
    sticky = (a.mant & ((1ULL << (sh - 1)) - 1)) != 0;
  }
  if(guard && (sticky || (m & 1))) m++; // reaching 2^23 is the smallest normal: the bit pattern is already right
  return __uint_as_float(s | m);
}

// (float)(SQRT3 * (long double)f)
__device__ __forceinline__ float mul_sqrt3(float f)
{
  if(!finite_f32(f)) return f * 1.7320508f; // NaN stays NaN, infinity keeps its sign
  return to_float(mul_const(from_float(f), SQRT3_MANT, SQRT3_EXP));
}
// (float)((long double)t + sign * (long double)c / SQRT12), sign = +1 or -1
__device__ __forceinline__ float add_div_sqrt12(float t, float c, int sign)
{
  if(!finite_f32(t) || !finite_f32(c)) return sign > 0 ? t + c : t - c; // the class of the result (NaN / infinity) is what a float sum gives
  ext q = div_const(from_float(c), SQRT3_MANT, SQRT12_EXP);
  if(sign < 0) q.sign ^= 1;
  return to_float(add(from_float(t), q));
}
} // namespace x87
