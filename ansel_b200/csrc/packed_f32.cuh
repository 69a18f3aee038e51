// Pairs of floats in the packed FP32 instructions of sm_100 (FADD2 / FMUL2 / FFMA2, one issue slot for two IEEE operations).
// Every operation is round-to-nearest, flush-to-zero, lane by lane identical to the scalar instruction the library's flags
// (--fmad=false -ftz=true) give.  ptxas contracts a packed multiply feeding a packed add or subtract into FFMA2 even under
// --fmad=false: callers must never chain the two packed (form the product or the sum per lane instead);
// tests/test_cpu_abi.py counts the FFMA2 in the SASS of the kernels using this header.
// With B200_KERNELS_ON_CPU (tests/emul) the same names are plain scalar code.
#pragma once
#ifdef B200_KERNELS_ON_CPU
struct f2
{
  float x, y;
};
static inline f2 mk2(float x, float y) { return f2{ x, y }; }
static inline f2 add2(f2 a, f2 b) { return f2{ a.x + b.x, a.y + b.y }; }
static inline f2 sub2(f2 a, f2 b) { return f2{ a.x - b.x, a.y - b.y }; }
static inline f2 mul2(f2 a, f2 b) { return f2{ a.x * b.x, a.y * b.y }; }
static inline f2 fma2(f2 a, f2 b, f2 c) { return f2{ fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y) }; }
static inline f2 neg2(f2 a) { return f2{ -a.x, -a.y }; }
static inline float min_nan(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }
static inline int __float2int_rz(float x) { return (int)x; } // cvttss2si: INT_MIN beyond the range, where the device saturates (same below -2^31)
#else
typedef float2 f2;
__device__ __forceinline__ f2 mk2(float x, float y) { return make_float2(x, y); }
__device__ __forceinline__ unsigned long long f2_bits(f2 a)
{
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ f2 bits_f2(unsigned long long r)
{
  f2 a;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
  return a;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b)
{
  unsigned long long c;
  asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(c) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(c);
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b)
{
  unsigned long long c;
  asm("sub.rn.ftz.f32x2 %0, %1, %2;" : "=l"(c) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(c);
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b)
{
  unsigned long long c;
  asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(c) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(c);
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c)
{
  unsigned long long d;
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
  return bits_f2(d);
}
__device__ __forceinline__ f2 neg2(f2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float min_nan(float a, float b) // a NaN stays a NaN (fminf drops it)
{
  float r;
  asm("min.NaN.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
#endif

