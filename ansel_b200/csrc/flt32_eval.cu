// b200_flt32_eval_dev: evaluate the device libm (flt32_math.cuh) over arrays, so its bit-compatibility
// with the host's glibc -- the libm the reference's CPU path calls -- can be checked on the GPU box.
#include "runtime.h"
#include "flt32_math.cuh"

namespace
{
__global__ void eval_kernel(int fn, const float *x, const float *y, float *out, size_t n)
{
  __shared__ double tabs[f32m::SMEM_DOUBLES];
  // odd blocks read the tables from shared memory, even blocks from global: both paths are covered
  f32m::tables_t tb = f32m::global_tables();
  if(blockIdx.x & 1) tb = f32m::stage_tables(tabs, threadIdx.x, blockDim.x);
  __syncthreads();
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(k >= n) return;
  float r;
  switch(fn)
  {
    case B200_FLT32_EXPF: r = f32m::expf_(tb, x[k]); break;
    case B200_FLT32_EXP2F: r = f32m::exp2f_(tb, x[k]); break;
    case B200_FLT32_LOGF: r = f32m::logf_(tb, x[k]); break;
    case B200_FLT32_LOG2F: r = f32m::log2f_(tb, x[k]); break;
    case B200_FLT32_SINF: r = f32m::sinf_(x[k]); break;
    case B200_FLT32_COSF: r = f32m::cosf_(x[k]); break;
    case B200_FLT32_ATANF: r = f32m::atanf_(x[k]); break;
    case B200_FLT32_ATAN2F: r = f32m::atan2f_(x[k], y[k]); break;
    case B200_FLT32_HYPOTF: r = f32m::hypotf_(x[k], y[k]); break;
    default: r = f32m::powf_(tb, x[k], y[k]); break;
  }
  out[k] = r;
}
} // namespace

using namespace b200;
extern "C" int b200_flt32_eval_dev(int fn, const float *d_x, const float *d_y, float *d_out, size_t n, void *stream)
{
  const bool two = fn == B200_FLT32_POWF || fn == B200_FLT32_ATAN2F || fn == B200_FLT32_HYPOTF;
  if(fn < B200_FLT32_EXPF || fn > B200_FLT32_HYPOTF || !d_x || !d_out || (two && !d_y))
    return fail(B200_ERR_ARG, "flt32_eval: bad arguments");
  int rc = bind_device(-1);
  if(rc) return rc;
  if(!n) return B200_OK;
  eval_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(fn, d_x, d_y, d_out, n);
  B200_CUDA_TRY(cudaGetLastError());
  return B200_OK;
}
