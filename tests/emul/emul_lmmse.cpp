/* Test infrastructure: the stages of ansel_b200/csrc/lmmse.cu compiled with g++ and run thread by thread on the CPU, tile after tile, in
 * the order the kernel runs them -- the threads of every stage in DESCENDING order (or ascending), so that a stage whose sites read what the
 * same stage writes elsewhere shows up as a difference to the oracle.  Not part of the product. */
#define B200_KERNELS_ON_CPU
#include "cuda_on_cpu.h"
#include "../../include/b200iop.h"
#include "../../ansel_b200/csrc/lmmse.cu"
#include <vector>

extern "C" int emul_lmmse(float *out, const float *in, int width, int height, unsigned filters, int mode, const float *processed_maximum, int nthreads, int ascending)
{
  _mm_setcsr(_mm_getcsr() | 0x8040u);
  if(width < 16 || height < 16) return 0;
  lm_args_t a;
  lm_plan(a, width, height, filters, mode, processed_maximum);
  std::vector<float> tables(2 * 65536), Q((size_t)6 * NP, __builtin_nanf(""));
  lm_gamma_tables(tables.data(), tables.data() + 65536);
  a.in = in;
  a.out = (float4 *)out;
  a.gamma_in = tables.data();
  a.gamma_out = tables.data() + 65536;
  for(int t = 0; t < a.nv * a.nh; t++)
  {
    const lm_tile_t T = lm_tile_of(a, t);
#define STAGE(call)                                                                                                    \
  for(int k = 0; k < nthreads; k++)                                                                                    \
  {                                                                                                                    \
    const int tid = ascending ? k : nthreads - 1 - k;                                                                  \
    call;                                                                                                              \
  }
    STAGE(lm_load(a, T, Q.data(), tid, nthreads))
    STAGE(lm_encode(a, T, Q.data(), tid, nthreads))
    STAGE(lm_differences(a, T, Q.data(), tid, nthreads))
    STAGE(lm_lowpass(a, T, Q.data(), tid, nthreads))
    STAGE(lm_interpolate(a, T, Q.data(), tid, nthreads))
    STAGE(lm_colours(a, T, Q.data(), tid, nthreads))
    STAGE(lm_rb_at_green(a, T, Q.data(), tid, nthreads))
    STAGE(lm_rb_at_rb(a, T, Q.data(), tid, nthreads))
    for(int pass = 0; pass < a.medians; pass++)
    {
      STAGE(lm_medians(a, T, Q.data(), tid, nthreads))
      STAGE(lm_rebuild(a, T, Q.data(), tid, nthreads))
    }
    STAGE(lm_restore(a, T, Q.data(), tid, nthreads))
    for(int step = 0; step < a.refine; step++)
      for(int which = 0; which < 3; which++) STAGE(lm_refine(a, T, Q.data(), which, tid, nthreads))
    STAGE(lm_store(a, T, Q.data(), tid, nthreads))
#undef STAGE
  }
  return 0;
}
