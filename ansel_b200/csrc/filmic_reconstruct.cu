// filmic rgb: the wavelet highlight reconstruction in front of the tone mapping (deprecated in the reference and off
// by default -- `hl_deprecated` -- but still what existing edits with a reconstruction threshold below 16 EV run).
//
// Reference: src/iop/filmicrgb.c process() :2729-2838, mask_clipped_pixels :1201-1228, inpaint_noise :1230-1270,
// wavelets_reconstruct_RGB :1272-1324, wavelets_reconstruct_ratios :1326-1384, init_reconstruct :1387-1400,
// wavelets_detail_level :1403-1411, get_scales :1414-1431, reconstruct_highlights :1434-1532, compute_ratios
// :2604-2619, restore_ratios :2622-2639; iop/noise_generator.h (splitmix32, xoshiro128+, uniform / gaussian /
// poissonian `_simd` generators :129-204); pixel/bspline.h blur_2D_Bspline :330-350.
// Parity: bit-identical to oracle/restate/filmic_reconstruct_oracle.c (pinned to the reference functions).
//
// Everything is pointwise or a separable 5-tap stencil: HBM-bound streaming kernels, one thread per pixel (float4).
// Per scale: 2 blurs (4 passes) + detail split + accumulation = ~15 RGBA plane passes; 1 + iterations reconstructions.
#ifndef B200_KERNELS_ON_CPU // tests/emul compiles the kernels of this file with g++ to check them against the oracle without a GPU
#include "runtime.h"
#endif
#include "flt32_math.cuh"
#include <math.h>

namespace
{
constexpr int NT = 256;
constexpr int MAX_SCALES = 10;

// IEEE division by a constant: nvcc turns `x / c` into `x * (1/c)` under -ftz=true (see labglue.cu)
__device__ __forceinline__ float divc(float a, float b)
{
#ifdef __CUDA_ARCH__
  float q;
  asm("div.rn.ftz.f32 %0, %1, %2;" : "=f"(q) : "f"(a), "f"(b));
  return q;
#else
  return a / b; // host pass: never called by the product
#endif
}
__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
__device__ __forceinline__ float fmaxabsf(float a, float b) { return (fabsf(a) > fabsf(b) && !isnan(a)) ? a : (isnan(b) ? 0.f : b); }
__device__ __forceinline__ uint32_t splitmix32(uint64_t seed)
{
  uint64_t r = (seed ^ (seed >> 33)) * 0x62a9d9ed799705f5ull;
  r = (r ^ (r >> 28)) * 0xcb24d0a5c88c35b3ull;
  return (uint32_t)(r >> 32);
}
__device__ __forceinline__ float xoshiro128plus(uint32_t (&st)[4])
{
  const uint32_t result = st[0] + st[3];
  const uint32_t t = st[1] << 9;
  st[2] ^= st[0];
  st[3] ^= st[1];
  st[1] ^= st[2];
  st[0] ^= st[3];
  st[2] ^= t;
  st[3] = (st[3] << 11) | (st[3] >> 21);
  return (float)(result >> 8) * 0x1.0p-24f;
}
__device__ __forceinline__ float box_muller(const f32m::tables_t &tb, float u1, float u2, bool flip)
{
  const float radius = sqrtf(-2.0f * f32m::logf_(tb, u1));
  const float angle = (float)(6.283185307179586 * (double)u2); // 2.f * M_PI * u2: a double product in the source
  return flip ? radius * f32m::cosf_(angle) : radius * f32m::sinf_(angle);
}

// mask_clipped_pixels(): sigmoid weight per pixel, and the count of pixels worth recovering
__global__ void __launch_bounds__(NT) mask_kernel(const float4 *__restrict__ in, float *__restrict__ mask, int *__restrict__ clipped, size_t npx,
                                                  float normalize, float feathering)
{
  const size_t k = (size_t)blockIdx.x * NT + threadIdx.x;
  int mine = 0;
  if(k < npx)
  {
    const f32m::tables_t tb = f32m::global_tables();
    const float4 p = __ldg(in + k);
    const float pix_max = fmaxf(sqrtf(p.x * p.x + p.y * p.y + p.z * p.z), 0.f);
    const float argument = -pix_max * normalize + feathering;
    mask[k] = clamp01(1.0f / (1.0f + f32m::exp2f_(tb, argument)));
    mine = (4.f > argument) ? 1 : 0;
  }
  const int n = __syncthreads_count(mine);
  if(threadIdx.x == 0 && n) atomicAdd(clipped, n);
}
// inpaint_noise(): statistical noise blended in by the mask weight; all four lanes as the vectorised reference build
// computes them (lane 3 draws nothing: u1 = u2 = 0)
__global__ void __launch_bounds__(NT) inpaint_noise_kernel(const float4 *__restrict__ in, const float *__restrict__ mask, float4 *__restrict__ out, int width,
                                                           int height, float noise_level, float threshold, int distribution)
{
  const int j = blockIdx.x * NT + threadIdx.x, i = blockIdx.y;
  if(j >= width) return;
  const f32m::tables_t tb = f32m::global_tables();
  const size_t idx = (size_t)i * width + j;
  uint32_t st[4] = { splitmix32((uint64_t)j + 1), splitmix32(((uint64_t)j + 1) * ((uint64_t)i + 3)), splitmix32(1337), splitmix32(666) };
  xoshiro128plus(st);
  xoshiro128plus(st);
  xoshiro128plus(st);
  xoshiro128plus(st);
  const float weight = __ldg(mask + idx);
  const float4 pv = __ldg(in + idx);
  const float mu[4] = { pv.x, pv.y, pv.z, pv.w };
  float sigma[4], noise[4];
#pragma unroll
  for(int c = 0; c < 4; c++) sigma[c] = mu[c] * noise_level / threshold;
  float u1[4] = { 0.f, 0.f, 0.f, 0.f }, u2[4] = { 0.f, 0.f, 0.f, 0.f };
  if(distribution == 1)
  { // gaussian_noise_simd :141-171: three u1, then three u2
#pragma unroll
    for(int c = 0; c < 3; c++) u1[c] = fmaxf(xoshiro128plus(st), 1.17549435e-38f);
#pragma unroll
    for(int c = 0; c < 3; c++) u2[c] = xoshiro128plus(st);
#pragma unroll
    for(int c = 0; c < 4; c++) noise[c] = box_muller(tb, u1[c], u2[c], (c & 1) == 0) * sigma[c] + mu[c];
  }
  else if(distribution == 2)
  { // poisson_noise_simd :174-204: u1, u2 interleaved; Anscombe transform
#pragma unroll
    for(int c = 0; c < 3; c++)
    {
      u1[c] = fmaxf(xoshiro128plus(st), 1.17549435e-38f);
      u2[c] = xoshiro128plus(st);
    }
#pragma unroll
    for(int c = 0; c < 4; c++)
    {
      const float g = box_muller(tb, u1[c], u2[c], (c & 1) == 0);
      const float r = g * sigma[c] + 2.0f * sqrtf(fmaxf(mu[c] + 0.375f, 0.0f));
      noise[c] = (r * r - sigma[c] * sigma[c]) * 0.25f - 0.375f;
    }
  }
  else
  { // uniform_noise_simd :129-138
#pragma unroll
    for(int c = 0; c < 3; c++) u1[c] = xoshiro128plus(st);
#pragma unroll
    for(int c = 0; c < 4; c++) noise[c] = mu[c] + 2.0f * (u1[c] - 0.5f) * sigma[c];
  }
  float o[4];
#pragma unroll
  for(int c = 0; c < 4; c++) o[c] = fmaxf(mu[c] * (1.0f - weight) + weight * noise[c], 0.f);
  out[idx] = make_float4(o[0], o[1], o[2], o[3]);
}
// init_reconstruct()
__global__ void __launch_bounds__(NT) init_kernel(const float4 *__restrict__ in, const float *__restrict__ mask, float4 *__restrict__ rec, size_t npx)
{
  const size_t k = (size_t)blockIdx.x * NT + threadIdx.x;
  if(k >= npx) return;
  const float4 p = __ldg(in + k);
  const float w = 1.f - __ldg(mask + k);
  rec[k] = make_float4(fmaxf(p.x * w, 0.f), fmaxf(p.y * w, 0.f), fmaxf(p.z * w, 0.f), fmaxf(p.w * w, 0.f));
}
// sparse_scalar_product(), bspline.h:83-117
template <bool CLIP> __device__ __forceinline__ float bs5(float a, float b, float c, float d, float e)
{
  const float v = 0.0625f * a + 0.25f * b + 0.375f * c + 0.25f * d + 0.0625f * e;
  return CLIP ? (0.0f > v ? 0.0f : v) : v;
}
template <bool CLIP> __device__ __forceinline__ float4 bs5_4(const float4 &p0, const float4 &p1, const float4 &p2, const float4 &p3, const float4 &p4)
{
  return make_float4(bs5<CLIP>(p0.x, p1.x, p2.x, p3.x, p4.x), bs5<CLIP>(p0.y, p1.y, p2.y, p3.y, p4.y), bs5<CLIP>(p0.z, p1.z, p2.z, p3.z, p4.z),
                     bs5<CLIP>(p0.w, p1.w, p2.w, p3.w, p4.w));
}
template <bool CLIP> __global__ void __launch_bounds__(NT) blur_vertical_kernel(const float4 *__restrict__ in, float4 *__restrict__ tmp, int width, int height, int mult)
{
  const int j = blockIdx.x * NT + threadIdx.x, i = blockIdx.y;
  if(j >= width) return;
  const float4 *b = in + j;
  tmp[(size_t)width * i + j] = bs5_4<CLIP>(__ldg(b + (size_t)width * max(i - 2 * mult, 0)), __ldg(b + (size_t)width * max(i - mult, 0)), __ldg(b + (size_t)width * i),
                                           __ldg(b + (size_t)width * min(i + mult, height - 1)), __ldg(b + (size_t)width * min(i + 2 * mult, height - 1)));
}
// horizontal pass of the LF blur fused with wavelets_detail_level(): LF, and HF = texture = detail - LF.
// HFt may be the buffer `detail` itself (the reference aliases them from the second scale on): elementwise, same index.
__global__ void __launch_bounds__(NT) blur_horizontal_detail_kernel(const float4 *__restrict__ tmp, const float4 *detail, float4 *__restrict__ LF, float4 *HFt,
                                                                    float4 *__restrict__ texture, int width, int mult)
{
  const int j = blockIdx.x * NT + threadIdx.x;
  if(j >= width) return;
  const size_t row = (size_t)width * blockIdx.y;
  const float4 *t = tmp + row;
  const float4 lf = bs5_4<true>(__ldg(t + max(j - 2 * mult, 0)), __ldg(t + max(j - mult, 0)), __ldg(t + j), __ldg(t + min(j + mult, width - 1)),
                                __ldg(t + min(j + 2 * mult, width - 1)));
  const float4 v = detail[row + j];
  const float4 hf = make_float4(v.x - lf.x, v.y - lf.y, v.z - lf.z, v.w - lf.w);
  LF[row + j] = lf;
  HFt[row + j] = hf;
  texture[row + j] = hf;
}
// horizontal pass of the un-clipped HF blur (mult 1) fused with wavelets_reconstruct_RGB / _ratios: the blurred HF
// of a pixel is consumed by that pixel only, so it never reaches memory
struct rec_args_t
{
  float gamma, gamma_comp, beta, beta_comp, delta;
  int last; // s == scales - 1
};
template <int VARIANT>
__global__ void __launch_bounds__(NT) blur_horizontal_reconstruct_kernel(const float4 *__restrict__ tmp, const float4 *__restrict__ LFp, const float4 *__restrict__ texture,
                                                                         const float *__restrict__ mask, float4 *__restrict__ rec, int width, const rec_args_t a)
{
  const int j = blockIdx.x * NT + threadIdx.x;
  if(j >= width) return;
  const size_t row = (size_t)width * blockIdx.y;
  const float4 *t = tmp + row;
  const float4 hf4 = bs5_4<false>(__ldg(t + max(j - 2, 0)), __ldg(t + max(j - 1, 0)), __ldg(t + j), __ldg(t + min(j + 1, width - 1)), __ldg(t + min(j + 2, width - 1)));
  const float4 lf4 = __ldg(LFp + row + j), tt4 = __ldg(texture + row + j);
  const float alpha = __ldg(mask + row + j);
  const float HF[4] = { hf4.x, hf4.y, hf4.z, hf4.w }, LF[4] = { lf4.x, lf4.y, lf4.z, lf4.w }, TT[4] = { tt4.x, tt4.y, tt4.z, tt4.w };
  const float grey_texture = fmaxabsf(fmaxabsf(TT[0], TT[1]), TT[2]);
  const float grey_details = divc(HF[0] + HF[1] + HF[2], 3.f);
  float4 r4 = rec[row + j];
  float r[4] = { r4.x, r4.y, r4.z, r4.w };
  if(VARIANT == 0)
  {
    const float grey_HF = a.beta_comp * (a.gamma_comp * grey_details + a.gamma * grey_texture);
    const float grey_residual = divc(a.beta_comp * (LF[0] + LF[1] + LF[2]), 3.f);
#pragma unroll
    for(int c = 0; c < 4; c++)
    {
      const float details = (a.gamma_comp * HF[c] + a.gamma * TT[c]) * a.beta + grey_HF;
      const float residual = a.last ? (grey_residual + LF[c] * a.beta) : 0.f;
      r[c] += alpha * (a.delta * details + residual);
    }
  }
  else
  {
    const float grey_HF = (a.gamma_comp * grey_details + a.gamma * grey_texture);
#pragma unroll
    for(int c = 0; c < 4; c++)
    {
      const float details = 0.5f * ((a.gamma_comp * HF[c] + a.gamma * TT[c]) + grey_HF);
      const float residual = a.last ? LF[c] : 0.f;
      r[c] += alpha * (a.delta * details + residual);
    }
  }
  rec[row + j] = make_float4(r[0], r[1], r[2], r[3]);
}
// compute_ratios() with the euclidean norm (v1), restore_ratios()
__global__ void __launch_bounds__(NT) ratios_kernel(const float4 *__restrict__ rec, float *__restrict__ norms, float4 *__restrict__ ratios, size_t npx)
{
  const size_t k = (size_t)blockIdx.x * NT + threadIdx.x;
  if(k >= npx) return;
  const float4 p = __ldg(rec + k);
  const float norm = fmaxf(sqrtf(p.x * p.x + p.y * p.y + p.z * p.z), 1.52587890625e-05f);
  norms[k] = norm;
  ratios[k] = make_float4(p.x / norm, p.y / norm, p.z / norm, p.w / norm);
}
__global__ void __launch_bounds__(NT) restore_kernel(float4 *__restrict__ rec, const float *__restrict__ norms, size_t npx)
{
  const size_t k = (size_t)blockIdx.x * NT + threadIdx.x;
  if(k >= npx) return;
  const float4 p = rec[k];
  const float n = __ldg(norms + k);
  rec[k] = make_float4(clamp01(p.x) * n, clamp01(p.y) * n, clamp01(p.z) * n, clamp01(p.w) * n);
}
} // namespace

#ifndef B200_KERNELS_ON_CPU
namespace b200
{
// get_scales(), filmicrgb.c:1414-1431
int filmic_reconstruct_scales(const b200_piece_t *piece)
{
  const float module_scale = (float)((double)piece->iscale / piece->roi_in.scale); // dt_dev_get_module_scale: float / double
  const float scale = 1.0f / module_scale;
  const float bh = piece->buf_in_height * piece->iscale, bw = piece->buf_in_width * piece->iscale;
  const size_t size = (size_t)((bh > bw) ? bh : bw);
  const int scales = (int)floorf(log2f((2.0f * size * scale / ((5 - 1) * 5)) - 1.0f));
  return scales > MAX_SCALES ? MAX_SCALES : (scales < 1 ? 1 : scales);
}

// reconstruct_highlights(), :1434-1532
static int reconstruct(const float4 *in, const float *mask, float4 *rec, int variant, const b200_filmicrgb_data_t *d, int scales, int width, int height,
                       float4 *LF_even, float4 *LF_odd, float4 *HF_grey, float4 *vtmp, cudaStream_t st)
{
  const size_t npx = (size_t)width * height;
  const unsigned lin = (unsigned)((npx + NT - 1) / NT);
  const dim3 grid((width + NT - 1) / NT, height);
  init_kernel<<<lin, NT, 0, st>>>(in, mask, rec, npx);
  rec_args_t a;
  a.gamma = d->reconstruct_structure_vs_texture;
  a.gamma_comp = 1.0f - d->reconstruct_structure_vs_texture;
  a.beta = d->reconstruct_grey_vs_color;
  a.beta_comp = 1.f - d->reconstruct_grey_vs_color;
  a.delta = d->reconstruct_bloom_vs_details;
  for(int s = 0; s < scales; ++s)
  {
    const float4 *detail = s == 0 ? in : (s % 2 != 0 ? LF_odd : LF_even);
    float4 *LF = s == 0 ? LF_odd : (s % 2 != 0 ? LF_even : LF_odd);
    float4 *HF_temp = s == 0 ? LF_even : (s % 2 != 0 ? LF_odd : LF_even);
    blur_vertical_kernel<true><<<grid, NT, 0, st>>>(detail, vtmp, width, height, 1 << s);
    blur_horizontal_detail_kernel<<<grid, NT, 0, st>>>(vtmp, detail, LF, HF_temp, HF_grey, width, 1 << s);
    blur_vertical_kernel<false><<<grid, NT, 0, st>>>(HF_temp, vtmp, width, height, 1);
    a.last = (s == scales - 1) ? 1 : 0;
    if(variant == 0)
      blur_horizontal_reconstruct_kernel<0><<<grid, NT, 0, st>>>(vtmp, LF, HF_grey, mask, rec, width, a);
    else
      blur_horizontal_reconstruct_kernel<1><<<grid, NT, 0, st>>>(vtmp, LF, HF_grey, mask, rec, width, a);
  }
  B200_CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

// process() :2729-2838.  *d_use: what the tone mapping reads -- d_in when fewer than 10 pixels are clipped, else the
// reconstructed frame (device scratch of the calling thread, valid until the next filmic call on it).
int filmic_reconstruct_dev(const b200_piece_t *piece, const b200_filmicrgb_data_t *d, const float *d_in, const float **d_use, cudaStream_t st)
{
  const int width = piece->roi_out.width, height = piece->roi_out.height;
  const size_t npx = (size_t)width * height;
  *d_use = d_in;
  if(!npx) return B200_OK;
  void *base = nullptr;
  const size_t npx4 = (npx + 3) & ~(size_t)3; // keeps the RGBA planes behind the two scalar planes 16-byte aligned
  int rc = scratch(SLOT_TMP1, 256 + 2 * npx4 * sizeof(float) + 6 * npx * sizeof(float4), &base);
  if(rc) return rc;
  int *d_count = (int *)base;
  float *mask = (float *)((char *)base + 256), *norms = mask + npx4;
  float4 *planes = (float4 *)(norms + npx4);
  float4 *inpainted = planes, *ratios = planes, *rec = planes + npx, *LF_even = planes + 2 * npx, *LF_odd = planes + 3 * npx, *HF_grey = planes + 4 * npx,
         *vtmp = planes + 5 * npx;
  (void)MAX_SCALES;
  const unsigned lin = (unsigned)((npx + NT - 1) / NT);
  B200_CUDA_TRY(cudaMemsetAsync(d_count, 0, sizeof(int), st));
  mask_kernel<<<lin, NT, 0, st>>>((const float4 *)d_in, mask, d_count, npx, d->normalize, d->reconstruct_feather);
  B200_CUDA_TRY(cudaGetLastError());
  int clipped = 0; // the one host decision of the path: is the recovery worth running (:1226)
  B200_CUDA_TRY(cudaMemcpyAsync(&clipped, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
  B200_CUDA_TRY(cudaStreamSynchronize(st));
  if(!(clipped > 9)) return B200_OK;

  const float module_scale = (float)((double)piece->iscale / piece->roi_in.scale);
  const float scale = fmaxf(module_scale, 1.f);
  const dim3 grid((width + NT - 1) / NT, height);
  inpaint_noise_kernel<<<grid, NT, 0, st>>>((const float4 *)d_in, mask, inpainted, width, height, d->noise_level / scale, d->reconstruct_threshold,
                                            d->noise_distribution);
  B200_CUDA_TRY(cudaGetLastError());
  const int scales = filmic_reconstruct_scales(piece);
  if((rc = reconstruct(inpainted, mask, rec, 0, d, scales, width, height, LF_even, LF_odd, HF_grey, vtmp, st))) return rc;
  for(int i = 0; i < d->high_quality_reconstruction; i++)
  {
    ratios_kernel<<<lin, NT, 0, st>>>(rec, norms, ratios, npx);
    if((rc = reconstruct(ratios, mask, rec, 1, d, scales, width, height, LF_even, LF_odd, HF_grey, vtmp, st))) return rc;
    restore_kernel<<<lin, NT, 0, st>>>(rec, norms, npx);
  }
  B200_CUDA_TRY(cudaGetLastError());
  *d_use = (const float *)rec;
  return B200_OK;
}
} // namespace b200
#endif // B200_KERNELS_ON_CPU
