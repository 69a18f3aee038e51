/* Array entry points for the libm restatement (flt32_math.h) and for the system libm itself,
 * so tests/test_flt32_math.py can pin one against the other.  TEST INFRASTRUCTURE ONLY. */
#include "flt32_math.h"
#include <stddef.h>

#define ARRAY1(name, fn)                                                     \
  void name(const float *x, float *out, size_t n)                            \
  {                                                                          \
    _Pragma("omp parallel for schedule(static)") for(size_t k = 0; k < n; k++) out[k] = fn(x[k]); \
  }
#define ARRAY2(name, fn)                                                     \
  void name(const float *x, const float *y, float *out, size_t n)            \
  {                                                                          \
    _Pragma("omp parallel for schedule(static)") for(size_t k = 0; k < n; k++) out[k] = fn(x[k], y[k]); \
  }

ARRAY1(orc_expf_array, f32m_expf)
ARRAY1(orc_exp2f_array, f32m_exp2f)
ARRAY1(orc_logf_array, f32m_logf)
ARRAY1(orc_log2f_array, f32m_log2f)
ARRAY2(orc_powf_array, f32m_powf)
ARRAY1(orc_sinf_array, f32m_sinf)
ARRAY1(orc_cosf_array, f32m_cosf)
ARRAY1(orc_atanf_array, f32m_atanf)
ARRAY2(orc_atan2f_array, f32m_atan2f)
ARRAY2(orc_hypotf_array, f32m_hypotf)

/* the system libm, called through volatile function pointers so nothing is folded or vectorised */
static float (*volatile sys_expf)(float) = expf;
static float (*volatile sys_exp2f)(float) = exp2f;
static float (*volatile sys_logf)(float) = logf;
static float (*volatile sys_log2f)(float) = log2f;
static float (*volatile sys_powf)(float, float) = powf;
static float (*volatile sys_sinf)(float) = sinf;
static float (*volatile sys_cosf)(float) = cosf;
static float (*volatile sys_atanf)(float) = atanf;
static float (*volatile sys_atan2f)(float, float) = atan2f;
static float (*volatile sys_hypotf)(float, float) = hypotf;
ARRAY1(sys_expf_array, sys_expf)
ARRAY1(sys_exp2f_array, sys_exp2f)
ARRAY1(sys_logf_array, sys_logf)
ARRAY1(sys_log2f_array, sys_log2f)
ARRAY2(sys_powf_array, sys_powf)
ARRAY1(sys_sinf_array, sys_sinf)
ARRAY1(sys_cosf_array, sys_cosf)
ARRAY1(sys_atanf_array, sys_atanf)
ARRAY2(sys_atan2f_array, sys_atan2f)
ARRAY2(sys_hypotf_array, sys_hypotf)
