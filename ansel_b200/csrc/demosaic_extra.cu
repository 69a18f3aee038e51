// The demosaic module's optional passes around the demosaicer: green equilibration of the mosaic before it, median
// colour smoothing of the result after it.
//
// Reference: src/iop/demosaic/basic.c color_smoothing :192-245, green_equilibration_lavg :248-293,
// green_equilibration_favg :296-329; called from iop/demosaic.c process() :1137-1170 (threshold = 1e-4 * ISO, :1049)
// and :1249-1250.  All streaming: 4 B/px in + 4 B/px out for the mosaic passes, 16 B in + 16 B out per smoothing
// sub-pass (two per pass: red, blue).
// The full average sums both green planes in double; the reference does so with an OpenMP reduction of undefined
// order, here a fixed two-stage tree -- the ratio may differ in its last bits, a pixel then by one ULP at most.
#include "runtime.h"

namespace
{
__device__ __forceinline__ int fc(int row, int col, uint32_t f)
{
  return (int)((f >> (((((unsigned)row << 1) & 14u) + ((unsigned)col & 1u)) << 1)) & 3u);
}

// ---- local average (:248-293) ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) green_eq_lavg_kernel(const float *__restrict__ in, float *__restrict__ out, int width, int height, int oj, int oi,
                                                            float thr)
{
  const int i = blockIdx.x * 256 + threadIdx.x, j = blockIdx.y;
  if(i >= width) return;
  const size_t k = (size_t)j * width + i;
  float v = __ldg(in + k);
  if(j >= oj && j + 2 < height && i >= oi && i + 2 < width && ((j - oj) & 1) == 0 && ((i - oi) & 1) == 0)
  {
    const float maximum = 1.0f;
    const float o1_1 = __ldg(in + k - width - 1), o1_2 = __ldg(in + k - width + 1), o1_3 = __ldg(in + k + width - 1), o1_4 = __ldg(in + k + width + 1);
    const float o2_1 = __ldg(in + k - 2 * (size_t)width), o2_2 = __ldg(in + k + 2 * (size_t)width), o2_3 = __ldg(in + k - 2), o2_4 = __ldg(in + k + 2);
    const float m1 = (o1_1 + o1_2 + o1_3 + o1_4) / 4.0f; // powers of two: exact either way
    const float m2 = (o2_1 + o2_2 + o2_3 + o2_4) / 4.0f;
    if((m2 > 0.0f) && (m1 > 0.0f) && (m1 / m2 < maximum * 2.0f))
    {
      const float c1 = __fdiv_rn(fabsf(o1_1 - o1_2) + fabsf(o1_1 - o1_3) + fabsf(o1_1 - o1_4) + fabsf(o1_2 - o1_3) + fabsf(o1_3 - o1_4) + fabsf(o1_2 - o1_4), 6.0f);
      const float c2 = __fdiv_rn(fabsf(o2_1 - o2_2) + fabsf(o2_1 - o2_3) + fabsf(o2_1 - o2_4) + fabsf(o2_2 - o2_3) + fabsf(o2_3 - o2_4) + fabsf(o2_2 - o2_4), 6.0f);
      if((v < maximum * 0.95f) && (c1 < maximum * thr) && (c2 < maximum * thr)) v = v * m1 / m2;
    }
  }
  out[k] = v;
}

// ---- full average (:296-329) ----------------------------------------------------------------------------
constexpr int FAVG_BLOCKS = 1024;
__global__ void __launch_bounds__(256) green_eq_favg_sum_kernel(const float *__restrict__ in, double *__restrict__ partial, int width, int height, int oi,
                                                                int g2_offset)
{
  __shared__ double s1[256], s2[256];
  const int ni = (width - 1 - g2_offset - oi + 1) / 2; // number of i values: oi, oi+2, ... < width-1-g2_offset
  const long long nj = (height - 1 + 1) / 2;           // j = 0, 2, ... < height-1
  const long long total = ni > 0 ? nj * ni : 0;
  double a = 0.0, b = 0.0;
  for(long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)FAVG_BLOCKS * 256)
  {
    const long long jj = e / ni;
    const int i = oi + 2 * (int)(e - jj * ni);
    const size_t j = (size_t)(2 * jj);
    a += (double)__ldg(in + j * width + i);
    b += (double)__ldg(in + (j + 1) * width + i + g2_offset);
  }
  s1[threadIdx.x] = a;
  s2[threadIdx.x] = b;
  __syncthreads();
  for(int st = 128; st > 0; st >>= 1)
  {
    if(threadIdx.x < st)
    {
      s1[threadIdx.x] += s1[threadIdx.x + st];
      s2[threadIdx.x] += s2[threadIdx.x + st];
    }
    __syncthreads();
  }
  if(threadIdx.x == 0)
  {
    partial[2 * blockIdx.x] = s1[0];
    partial[2 * blockIdx.x + 1] = s2[0];
  }
}
__global__ void green_eq_favg_ratio_kernel(double *partial)
{ // one thread: fixed order; partial[2*FAVG_BLOCKS] receives the ratio, or -1 when the reference returns early (:318-321)
  double a = 0.0, b = 0.0;
  for(int k = 0; k < FAVG_BLOCKS; k++)
  {
    a += partial[2 * k];
    b += partial[2 * k + 1];
  }
  partial[2 * FAVG_BLOCKS] = (a > 0.0 && b > 0.0) ? b / a : -1.0;
}
__global__ void __launch_bounds__(256) green_eq_favg_apply_kernel(const float *__restrict__ in, float *__restrict__ out, const double *__restrict__ partial,
                                                                  int width, int height, int oi, int g2_offset)
{
  const int i = blockIdx.x * 256 + threadIdx.x, j = blockIdx.y;
  if(i >= width) return;
  const size_t k = (size_t)j * width + i;
  const double ratio = partial[2 * FAVG_BLOCKS];
  float v = __ldg(in + k);
  if(ratio >= 0.0 && (j & 1) == 0 && j < height - 1 && i >= oi && ((i - oi) & 1) == 0 && i < width - 1 - g2_offset) v = (float)((double)v * ratio);
  out[k] = v;
}

// ---- colour smoothing (:192-245) ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) smooth_stash_kernel(float4 *__restrict__ px, size_t n, int c)
{
  const size_t k = (size_t)blockIdx.x * 256 + threadIdx.x;
  if(k >= n) return;
  float4 p = px[k];
  p.w = c == 0 ? p.x : p.z;
  px[k] = p;
}
#define SWAPMED(I, J)             \
  if(med[I] > med[J])             \
  {                               \
    const float tmp = med[J];     \
    med[J] = med[I];              \
    med[I] = tmp;                 \
  }
__global__ void __launch_bounds__(256) smooth_median_kernel(float *__restrict__ out, int width, int height, int c)
{
  const int i = 1 + blockIdx.x * 256 + threadIdx.x, j = 1 + blockIdx.y;
  if(i >= width - 1 || j >= height - 1) return;
  float *outp = out + 4 * ((size_t)j * width + i);
  const size_t w4 = 4 * (size_t)width;
  float med[9];
#pragma unroll
  for(int dj = -1; dj <= 1; dj++)
#pragma unroll
    for(int di = -1; di <= 1; di++)
    {
      const float *q = outp + dj * (long long)w4 + 4 * di;
      med[3 * (dj + 1) + (di + 1)] = q[3] - q[1];
    }
  SWAPMED(1, 2) SWAPMED(4, 5) SWAPMED(7, 8) SWAPMED(0, 1) SWAPMED(3, 4) SWAPMED(6, 7) SWAPMED(1, 2) SWAPMED(4, 5) SWAPMED(7, 8)
  SWAPMED(0, 3) SWAPMED(5, 8) SWAPMED(4, 7) SWAPMED(3, 6) SWAPMED(1, 4) SWAPMED(2, 5) SWAPMED(4, 7) SWAPMED(4, 2) SWAPMED(6, 4)
  SWAPMED(4, 2)
  outp[c] = fmaxf(med[4] + outp[1], 0.0f);
}
} // namespace

namespace b200
{
// green_eq: dt_iop_demosaic_greeneq_t (1 local, 2 full, 3 both); d_tmp0/d_tmp1: mosaic-sized device buffers.
// Returns in *d_result the buffer that holds the equalised mosaic.
int demosaic_green_eq_dev(const float *d_in, float *d_tmp0, float *d_tmp1, double *d_partial, int width, int height, uint32_t dsc_filters, int x, int y,
                          unsigned green_eq, float threshold, const float **d_result, cudaStream_t s)
{
  const dim3 grid((width + 255) / 256, height);
  const float *src = d_in;
  if(green_eq == 2 || green_eq == 3)
  { // green_equilibration_favg
    int oi = 0;
    if((b200_fc(0 + y, oi + x, dsc_filters) & 1) != 1) oi++;
    const int g2_offset = oi ? -1 : 1;
    green_eq_favg_sum_kernel<<<FAVG_BLOCKS, 256, 0, s>>>(src, d_partial, width, height, oi, g2_offset);
    green_eq_favg_ratio_kernel<<<1, 1, 0, s>>>(d_partial);
    green_eq_favg_apply_kernel<<<grid, 256, 0, s>>>(src, d_tmp0, d_partial, width, height, oi, g2_offset);
    B200_CUDA_TRY(cudaGetLastError());
    src = d_tmp0;
  }
  if(green_eq == 1 || green_eq == 3)
  { // green_equilibration_lavg
    int oj = 2, oi = 2;
    if(b200_fc(oj + y, oi + x, dsc_filters) != 1) oj++;
    if(b200_fc(oj + y, oi + x, dsc_filters) != 1) oi++;
    if(b200_fc(oj + y, oi + x, dsc_filters) != 1) oj--;
    float *dst = (src == d_tmp0) ? d_tmp1 : d_tmp0;
    green_eq_lavg_kernel<<<grid, 256, 0, s>>>(src, dst, width, height, oj, oi, threshold);
    B200_CUDA_TRY(cudaGetLastError());
    src = dst;
  }
  *d_result = src;
  return B200_OK;
}
int demosaic_green_eq_partial_doubles() { return 2 * FAVG_BLOCKS + 1; }

int demosaic_color_smoothing_dev(float *d_out, int width, int height, int passes, cudaStream_t s)
{
  const size_t n = (size_t)width * height;
  for(int pass = 0; pass < passes; pass++)
    for(int c = 0; c < 3; c += 2)
    {
      smooth_stash_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((float4 *)d_out, n, c);
      if(width > 2 && height > 2) smooth_median_kernel<<<dim3((width - 2 + 255) / 256, height - 2), 256, 0, s>>>(d_out, width, height, c);
    }
  B200_CUDA_TRY(cudaGetLastError());
  return B200_OK;
}
} // namespace b200
