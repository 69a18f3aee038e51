/* Test infrastructure: just enough of the CUDA device vocabulary to compile the *kernels* of a .cu file with g++
 * and run their threads one after the other on the CPU.  Used by tests/test_cpu_kernel_emulation.py to check the
 * arithmetic and indexing of kernels against the oracle without a GPU.  What this cannot show is nvcc's code
 * generation (FTZ, division rewrites): that is what the `-m gpu` tests are for.  Never part of the product. */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstddef>
#include <xmmintrin.h>

struct float4
{
  float x, y, z, w;
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{ x, y, z, w }; }
struct uint2
{
  unsigned x, y;
};
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{ x, y }; }
struct dim3
{
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __grid_constant__
#define __ldg(p) (*(p))
#define __stcs(p, v) (*(p) = (v))
#define __ldcs(p) (*(p))
#define CUDART_NAN_F __builtin_nanf("")
#define CUDART_INF_F __builtin_inff()

static inline uint32_t __float_as_uint(float f)
{
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static inline float __uint_as_float(uint32_t u)
{
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline float __int_as_float(int i)
{
  float f;
  memcpy(&f, &i, 4);
  return f;
}
static inline long long __double_as_longlong(double d)
{
  long long u;
  memcpy(&u, &d, 8);
  return u;
}
static inline double __longlong_as_double(long long u)
{
  double d;
  memcpy(&d, &u, 8);
  return d;
}
static inline int __double2int_rz(double d) { return (int)d; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
/* block-wide primitives degrade to per-thread ones: a harness must not rely on what they reduce */
static inline void __syncthreads() {}
static inline int __syncthreads_count(int p) { return p; }
static inline int atomicAdd(int *p, int v)
{
  const int old = *p;
  *p += v;
  return old;
}
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v)
{
  const unsigned long long old = *p;
  *p += v;
  return old;
}
using std::isnan;

/* run `kernel(args...)` for every thread of a grid of 1-D blocks, sequentially; FTZ|DAZ like the device's -ftz=true */
template <typename K, typename... A> static void emulate(dim3 grid, unsigned threads, K kernel, A... args)
{
  _mm_setcsr(_mm_getcsr() | 0x8040u);
  blockDim = dim3(threads);
  gridDim = grid;
  for(unsigned bz = 0; bz < grid.z; bz++)
    for(unsigned by = 0; by < grid.y; by++)
      for(unsigned bx = 0; bx < grid.x; bx++)
        for(unsigned t = 0; t < threads; t++)
        {
          blockIdx = dim3(bx, by, bz);
          threadIdx = dim3(t);
          kernel(args...);
        }
}
