// Tone-curve evaluation shared by the colour conversions (color.cu) and the RGB <-> Lab glue (labglue.cu).
// colorprofiles/iop_profile.h: extrapolate_lut :536-545, eval_exp :559-562, dt_ioppr_eval_trc :577-580.
#pragma once
#include "flt32_math.cuh"
#include "runtime.h"

namespace
{
constexpr int LUTN = B200_LUT_SAMPLES;

// extrapolate_lut(), iop_profile.h:536-545
template <bool CONTRACT, bool DECODE_LOOP> __device__ __forceinline__ float lut_lerp(const float *lut, float v)
{
  const float scaled = v * (float)(LUTN - 1);
  const float ft = scaled > 0.0f ? (scaled < (float)(LUTN - 1) ? scaled : (float)(LUTN - 1)) : 0.0f; // CLAMPS: NaN -> 0
  const int t = (ft < (float)(LUTN - 2)) ? (int)ft : LUTN - 2;
  const float f = ft - (float)t;
  const float l1 = __ldg(lut + t), l2 = __ldg(lut + t + 1);
  if(!CONTRACT) return __fadd_rn(__fmul_rn(l1, 1.0f - f), __fmul_rn(l2, f));
  return DECODE_LOOP ? __fmaf_rn(l2, f, __fmul_rn(l1, 1.0f - f)) : __fmaf_rn(l1, 1.0f - f, __fmul_rn(l2, f));
}

// dt_ioppr_eval_trc(), iop_profile.h:577-580
template <bool CONTRACT, bool DECODE_LOOP>
__device__ __forceinline__ float eval_trc(const f32m::tables_t &tb, float x, const float *lut, const float *co)
{
  if(x < 1.0f) return lut_lerp<CONTRACT, DECODE_LOOP>(lut, x);
  return co[1] * f32m::powf_(tb, x * co[0], co[2]);
}
} // namespace
