// Non-local means, exact replay of the reference's accumulation order.
//
// Reference: src/pixel/nlmeans_core.c  nlmeans_denoise :315-532 (scatter :95-104, define_patches :107-145,
// pixel_difference :156-165, diff_of_pixels_diff :168-180, init_column_sums :214-264,
// compute_slice_height/width :267-312), dt_fast_mexp2f math/math.h:290-301; callers
// iop/denoiseprofile.c process_nlmeans_cpu :1599-1648 (nlmeans_norm :1456-1470, nlmeans_scattering
// :1474-1499, nlmeans_precondition :1500-1533, nlmeans_backtransform :1580-1597).
//
// The reference's result is order dependent: per ~60x72 chunk and per patch offset, patch distances are
// float running column sums updated incrementally down the rows, and a float running sum along each row;
// contributions are accumulated into the output in patch order.  Bit parity therefore fixes three
// sequential axes.  Mapping: one CTA per reference chunk, patches in order, and per patch
//   phase A  threads = columns : the column-sum recurrence down the rows   -> S[row][col] (shared)
//   phase B1 threads = rows    : the running sum along each row            -> D[row][col] (in place of S)
//   phase B2 threads = pixels  : weights and accumulation into the chunk's RGBA tile (shared)
// so every lane does independent work while each recurrence is evaluated in the reference's order.
// Not HBM bound by construction (~9 kflop/px at K=7); reported against FP32 issue.
#include "runtime.h"
#include <math.h>

namespace
{
constexpr int SLICE_WIDTH = 72, SLICE_HEIGHT = 60; // nlmeans_core.c:55-56
constexpr int MAX_RADIUS = 4;                      // patch radius 1..4 (:35)
constexpr int MAX_CH = SLICE_HEIGHT + 9;           // compute_slice_height() returns 51..69
constexpr int MAX_CW = SLICE_WIDTH;
constexpr int SW = MAX_CW + 2 * MAX_RADIUS + 2;    // 82 columns of column sums, +1 -> odd stride below
constexpr int SSTRIDE = SW + 1;                    // 83: odd, so lanes = rows hit distinct banks
#ifndef NLM_NT
#define NLM_NT 256
#endif
#ifndef NLM_MINB
#define NLM_MINB 2
#endif
#ifndef NLM_UNR
#define NLM_UNR 2
#endif
#ifndef NLM_UB
#define NLM_UB 2
#endif
constexpr int NT = NLM_NT;

struct patch_t
{
  short rows, cols;
};

struct nlm_args_t
{
  const float4 *in;
  float4 *out;
  const patch_t *patches;
  int n_patches;
  int width, height, chk_h, chk_w, n_cl;
  int radius;
  float center_weight, sharpness, cp_norm;
  float norm[4];
  float weight[4], invert[4];
  int skip_blend;
};

__device__ __forceinline__ float fast_mexp2(float x) // math/math.h:290-301
{
  const int i1 = 0x3f800000, i2 = 0x3f000000;
  const int k0 = i1 + (int)(x * (float)(i2 - i1));
  return __int_as_float(k0 >= 0x800000 ? k0 : 0);
}
__device__ __forceinline__ float pixdiff(const float4 a, const float4 b, const float *n) // :156-165
{
  const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z;
  return d0 * d0 * n[0] + d1 * d1 * n[1] + d2 * d2 * n[2];
}
__device__ __forceinline__ float diff_of_diffs(const float4 p1, const float4 p2, const float4 p3, const float4 p4, const float *n) // :168-180
{
  const float a0 = p1.x - p2.x, a1 = p1.y - p2.y, a2 = p1.z - p2.z;
  const float b0 = p3.x - p4.x, b1 = p3.y - p4.y, b2 = p3.z - p4.z;
  return (a0 * a0 - b0 * b0) * n[0] + (a1 * a1 - b1 * b1) * n[1] + (a2 * a2 - b2 * b2) * n[2];
}

__global__ void __launch_bounds__(NT, NLM_MINB) nlm_chunks_kernel(const __grid_constant__ nlm_args_t a)
{
  extern __shared__ __align__(16) float smem[];
  float4 *const tile = reinterpret_cast<float4 *>(smem);           // [MAX_CH][MAX_CW] accumulated RGBA
  float *const S = smem + 4 * MAX_CH * MAX_CW;                     // [MAX_CH][SSTRIDE] column sums / distortions
  const int tid = threadIdx.x;
  const int it = blockIdx.x / a.n_cl, il = blockIdx.x - it * a.n_cl;
  const int chunk_top = it * a.chk_h, chunk_left = il * a.chk_w;
  const int chunk_bot = min(chunk_top + a.chk_h, a.height), chunk_right = min(chunk_left + a.chk_w, a.width);
  const int width = a.width, height = a.height, radius = a.radius;
  const int cbase = chunk_left - radius - 1;              // column of S[.][0]
  const int ncols = (chunk_right + radius) - cbase;       // <= SW
  const float4 *const in = a.in;

  for(int k = tid; k < MAX_CH * MAX_CW; k += NT) tile[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  for(int p = 0; p < a.n_patches; p++)
  {
    const int srow = a.patches[p].rows, scol = a.patches[p].cols;
    const int row_min = max(chunk_top, max(0, -srow)), row_max = min(chunk_bot, height - max(0, srow));
    if(row_min >= row_max) continue; // uniform
    const int row_top = max(row_min, max(radius, radius - srow));
    const int row_bot = min(row_max, height - 1 - max(radius, radius + srow));
    const int col_min = max(chunk_left, -scol), col_max = min(chunk_right, width - scol);
    const int pcol_min = chunk_left - min(radius, min(chunk_left, chunk_left + scol));
    const int pcol_max = chunk_right + min(radius, min(width - chunk_right, width - (chunk_right + scol)));
    const long long poff = (long long)srow * width + scol; // patch offset in pixels
    const int nrows = row_max - row_min;

    // ---- phase A: column sums down the rows, one thread per column ---------------------------------
    if(tid < ncols)
    {
      const int col = cbase + tid;
      const bool live = col >= pcol_min && col < pcol_max;
      float cs = 0.0f;
      if(live)
      { // init_column_sums(), :214-264, at row = row_min
        const int row = row_min;
        const int rmin = row - min(radius, min(row, row + srow));
        const int rmax = row + min(radius, min(height - 1 - row, height - 1 - (row + srow)));
        float sum = 0.0f;
        for(int r = rmin; r <= rmax; r++)
        {
          const float4 *px = in + (size_t)r * width + col;
          sum += pixdiff(__ldg(px), __ldg(px + poff), a.norm);
        }
        cs = sum;
      }
      const int lim_a = min(row_top, row_bot);
      // The recurrence is sequential in `row`, its operands are not: fetch UNR rows' worth of pixels
      // first (independent loads in flight), then apply the updates in the reference's order.
      constexpr int UNR = NLM_UNR;
      for(int row0 = row_min; row0 < row_max; row0 += UNR)
      {
        float4 B0[UNR], B1[UNR], T0[UNR], T1[UNR];
        int kind[UNR];
#pragma unroll
        for(int u = 0; u < UNR; u++)
        {
          const int row = row0 + u;
          kind[u] = row >= row_max ? -1 : (row < lim_a ? 1 : (row < row_bot ? 2 : ((row >= row_top && row + 1 < row_max) ? 3 : 0)));
          if(live && (kind[u] == 1 || kind[u] == 2))
          {
            const float4 *b = in + (size_t)(row + 1 + radius) * width + col;
            B0[u] = __ldg(b);
            B1[u] = __ldg(b + poff);
          }
          if(live && (kind[u] == 2 || kind[u] == 3))
          {
            const float4 *t = in + (size_t)(row - radius) * width + col;
            T0[u] = __ldg(t);
            T1[u] = __ldg(t + poff);
          }
        }
#pragma unroll
        for(int u = 0; u < UNR; u++)
        {
          if(kind[u] < 0) break;
          S[(row0 + u - row_min) * SSTRIDE + tid] = cs;
          if(!live) continue;
          if(kind[u] == 1)
            cs += pixdiff(B0[u], B1[u], a.norm); // :424-440
          else if(kind[u] == 2)
            cs += diff_of_diffs(B0[u], B1[u], T0[u], T1[u], a.norm); // :441-466
          else if(kind[u] == 3)
            cs -= pixdiff(T0[u], T1[u], a.norm); // :467-483
        }
      }
    }
    __syncthreads();

    // ---- phase B1: running distortion along each row, one thread per row (:384-387,409) -----------
    if(tid < nrows && col_min < col_max)
    {
      float *const Sr = S + tid * SSTRIDE - cbase; // Sr[col] = column sum of `col` for this row
      float distortion = 0.0f;
      for(int i = col_min - radius; i < min(col_min + radius, col_max); i++) distortion += Sr[i];
      for(int col = col_min; col < col_max; col++)
      {
        distortion += (Sr[col + radius] - Sr[col - radius - 1]);
        Sr[col - radius - 1] = distortion; // that slot is never read again: keep D[col] there
      }
    }
    __syncthreads();

    // ---- phase B2: weights and accumulation, all threads over the chunk's pixels ------------------
    {
      const int nc = col_max - col_min;
      const int total = nc > 0 ? nrows * nc : 0;
      // operands of UB pixels first (independent loads in flight), then the arithmetic
      constexpr int UB = NLM_UB;
      for(int idx0 = tid; idx0 < total; idx0 += UB * NT)
      {
        float4 q[UB], c[UB];
        float dist[UB];
        int slot[UB];
#pragma unroll
        for(int u = 0; u < UB; u++)
        {
          const int idx = idx0 + u * NT;
          slot[u] = -1;
          if(idx < total)
          {
            const int rr = idx / nc, cc = idx - rr * nc;
            const int row = row_min + rr, col = col_min + cc;
            const float4 *px = in + (size_t)row * width + col;
            q[u] = __ldg(px + poff);
            if(a.center_weight >= 0) c[u] = __ldg(px);
            dist[u] = S[rr * SSTRIDE + (col - radius - 1 - cbase)];
            slot[u] = (row - chunk_top) * MAX_CW + (col - chunk_left);
          }
        }
#pragma unroll
        for(int u = 0; u < UB; u++)
        {
          if(slot[u] < 0) continue;
          float wt;
          if(a.center_weight < 0)
            wt = fast_mexp2(dist[u] * a.sharpness); // :389-402
          else
          { // :404-420
            const float d0 = c[u].x - q[u].x, d1 = c[u].y - q[u].y, d2 = c[u].z - q[u].z;
            const float pd = d0 * d0 * a.cp_norm + d1 * d1 * a.cp_norm + d2 * d2 * a.cp_norm;
            const float dissimilarity = (dist[u] + pd) / (1.0f + a.center_weight);
            wt = fast_mexp2(fmaxf(0.0f, dissimilarity * a.sharpness - 2.0f));
          }
          float4 *o = tile + slot[u];
          float4 v = *o;
          v.x += q[u].x * wt;
          v.y += q[u].y * wt;
          v.z += q[u].z * wt;
          v.w += 1.0f * wt;
          *o = v;
        }
      }
    }
    __syncthreads();
  }

  // ---- normalise (and blend) : :485-519 ---------------------------------------------------------------
  const int cw = chunk_right - chunk_left, ch = chunk_bot - chunk_top;
  for(int idx = tid; idx < cw * ch; idx += NT)
  {
    const int rr = idx / cw, cc = idx - rr * cw;
    const float4 v = tile[rr * MAX_CW + cc];
    const size_t g = (size_t)(chunk_top + rr) * width + chunk_left + cc;
    float4 o;
    if(a.skip_blend)
      o = make_float4(v.x / v.w, v.y / v.w, v.z / v.w, v.w / v.w);
    else
    {
      const float4 i4 = __ldg(in + g);
      o.x = (i4.x * a.invert[0]) + (v.x / v.w * a.weight[0]);
      o.y = (i4.y * a.invert[1]) + (v.y / v.w * a.weight[1]);
      o.z = (i4.z * a.invert[2]) + (v.z / v.w * a.weight[2]);
      o.w = (i4.w * a.invert[3]) + (v.w / v.w * a.weight[3]);
    }
    a.out[g] = o;
  }
}

// scatter(), :95-104: evaluated in double, truncated to int
int scatter(float scale, float scattering, int i1, int i2)
{
  const int a1 = abs(i1), a2 = abs(i2);
  const int sg = (i1 > 0) - (i1 < 0);
  return (int)(scale * ((a1 * a1 * a1 + 7.0 * a1 * sqrt((double)a2)) * sg * scattering / 6.0 + i1));
}
int slice_height(int height) // :267-296
{
  if(height % SLICE_HEIGHT == 0) return SLICE_HEIGHT;
  int best = height % SLICE_HEIGHT, best_incr = 0;
  for(int incr = 1; incr < 10; incr++)
  {
    const int plus_rem = height % (SLICE_HEIGHT + incr);
    if(plus_rem == 0) return SLICE_HEIGHT + incr;
    if(plus_rem > best)
    {
      best_incr = +incr;
      best = plus_rem;
    }
    const int minus_rem = height % (SLICE_HEIGHT - incr);
    if(minus_rem == 0) return SLICE_HEIGHT - incr;
    if(minus_rem > best)
    {
      best_incr = -incr;
      best = minus_rem;
    }
  }
  return SLICE_HEIGHT + best_incr;
}
int slice_width(int width) // :299-312
{
  int sl = SLICE_WIDTH;
  int rem = width % sl;
  if(rem < SLICE_WIDTH / 2 && (width % (sl - 4)) > rem)
  {
    sl -= 4;
    rem = width % sl;
    if(rem < SLICE_WIDTH / 2 && (width % (sl - 4)) > rem) sl -= 4;
  }
  return sl;
}
} // namespace

namespace b200
{
// nlmeans_denoise(), :315-532, on device RGBA buffers.  in != out.
int nlmeans_denoise_dev(const float *d_in, float *d_out, int width, int height, float scattering, float scale, float luma,
                        float chroma, float center_weight, float sharpness, int radius, int search_radius, int decimate,
                        const float norm[4], cudaStream_t stream)
{
  if(radius < 0 || radius > MAX_RADIUS) return fail(B200_ERR_UNSUPPORTED, "nlmeans: patch radius %d outside 0..%d", radius, MAX_RADIUS);
  if(search_radius < 0 || search_radius > 64) return fail(B200_ERR_ARG, "nlmeans: search radius %d", search_radius);
  int n_patches = (2 * search_radius + 1) * (2 * search_radius + 1);
  if(decimate) n_patches = (n_patches + 1) / 2;
  patch_t *h_patches = (patch_t *)malloc(sizeof(patch_t) * (size_t)n_patches);
  if(!h_patches) return fail(B200_ERR_NOMEM, "nlmeans: host allocation");
  { // define_patches(), :107-145
    int k = 0, dec = decimate;
    for(int ri = -search_radius; ri <= search_radius; ri++)
      for(int ci = -search_radius; ci <= search_radius; ci++)
      {
        if(dec && (++dec & 1)) continue;
        h_patches[k].rows = (short)scatter(scale, scattering, ri, ci);
        h_patches[k].cols = (short)scatter(scale, scattering, ci, ri);
        k++;
      }
  }
  void *d_patches = nullptr;
  int rc = scratch(SLOT_SMALL + 1, sizeof(patch_t) * (size_t)n_patches, &d_patches);
  if(rc)
  {
    free(h_patches);
    return rc;
  }
  cudaError_t e = cudaMemcpyAsync(d_patches, h_patches, sizeof(patch_t) * (size_t)n_patches, cudaMemcpyHostToDevice, stream);
  if(e == cudaSuccess) e = cudaStreamSynchronize(stream); // h_patches is freed below
  free(h_patches);
  if(e != cudaSuccess) return fail(B200_ERR_CUDA, "nlmeans: patch upload: %s", cudaGetErrorString(e));

  nlm_args_t a;
  a.in = (const float4 *)d_in;
  a.out = (float4 *)d_out;
  a.patches = (const patch_t *)d_patches;
  a.n_patches = n_patches;
  a.width = width;
  a.height = height;
  a.chk_h = slice_height(height);
  a.chk_w = slice_width(width);
  a.n_cl = (width + a.chk_w - 1) / a.chk_w;
  const int n_ct = (height + a.chk_h - 1) / a.chk_h;
  a.radius = radius;
  a.center_weight = center_weight;
  a.sharpness = sharpness;
  const int pw = 2 * radius + 1;
  a.cp_norm = center_weight * pw * pw; // compute_center_pixel_norm(), :147-153
  for(int c = 0; c < 4; c++) a.norm[c] = norm[c];
  a.weight[0] = luma;
  a.weight[1] = a.weight[2] = chroma;
  a.weight[3] = 1.0f;
  a.invert[0] = 1.0f - luma;
  a.invert[1] = a.invert[2] = 1.0f - chroma;
  a.invert[3] = 0.0f;
  a.skip_blend = (luma == 1.0 && chroma == 1.0) ? 1 : 0;
  if(a.chk_h > MAX_CH || a.chk_w > MAX_CW) return fail(B200_ERR_ARG, "nlmeans: chunk %dx%d exceeds the kernel's tile", a.chk_w, a.chk_h);

  const int smem_bytes = (4 * MAX_CH * MAX_CW + MAX_CH * SSTRIDE) * (int)sizeof(float);
  static bool attr_set[16] = { false };
  int dev = 0;
  B200_CUDA_TRY(cudaGetDevice(&dev));
  if(!attr_set[dev & 15])
  {
    B200_CUDA_TRY(cudaFuncSetAttribute(nlm_chunks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set[dev & 15] = true;
  }
  nlm_chunks_kernel<<<(unsigned)(n_ct * a.n_cl), NT, smem_bytes, stream>>>(a);
  B200_CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

int denoise_vst_forward(const b200_piece_t *piece, const b200_denoiseprofile_data_t *d, const float *d_in, float *d_out,
                        size_t npx, bool nlm, cudaStream_t s);
int denoise_vst_backward(const b200_piece_t *piece, const b200_denoiseprofile_data_t *d, float *d_buf, size_t npx, bool nlm,
                         cudaStream_t s);

// process_nlmeans_cpu(), denoiseprofile.c:1599-1648
int denoiseprofile_nlmeans_dev(const b200_piece_t *piece, const b200_denoiseprofile_data_t *d, const float *d_in, float *d_out,
                               cudaStream_t s)
{
  const int width = piece->roi_in.width, height = piece->roi_in.height;
  const size_t npx = (size_t)width * height;
  const float scale = fminf(fminf((float)piece->roi_in.scale, 2.0f), 1.0f);
  const int P = (int)ceilf(d->radius * scale);
  int K = (int)d->nbhood;
  float scattering = d->scattering;
  { // nlmeans_scattering(), :1474-1499.  dt_dev_pixelpipe_has_preview_output() is read as "preview pipe".
    const bool has_preview = piece->pipe_type == B200_PIPE_PREVIEW;
    if(has_preview || piece->pipe_type == B200_PIPE_THUMBNAIL)
    {
      const int maxk = (int)((K * K * K + 7.0 * K * sqrt((double)K)) * scattering / 6.0 + K);
      K = K < 3 ? K : 3;
      scattering = (float)((maxk - K) * 6.0 / (K * K * K + 7.0 * K * sqrt((double)K)));
    }
    if(!has_preview)
    {
      const int maxk = (int)((K * K * K + 7.0 * K * sqrt((double)K)) * scattering / 6.0 + K);
      const int k4 = K < 4 ? K : 4;
      const float ks = K * scale;
      K = (int)((float)k4 > ks ? (float)k4 : ks);
      scattering = (float)((maxk - K) * 6.0 / (K * K * K + 7.0 * K * sqrt((double)K)));
    }
  }
  float norm = .045f / ((2 * P + 1) * (2 * P + 1)); // nlmeans_norm(), :1456-1470
  if(!d->fix_anscombe_and_nlmeans_norm) norm = .015f / (2 * P + 1);
  const float central_pixel_weight = d->central_pixel_weight * scale;

  void *pre = nullptr;
  int rc = scratch(SLOT_TMP0, npx * 16, &pre);
  if(rc) return rc;
  if((rc = denoise_vst_forward(piece, d, d_in, (float *)pre, npx, true, s))) return rc;
  const float norm2[4] = { 1.0f, 1.0f, 1.0f, 1.0f };
  if((rc = nlmeans_denoise_dev((const float *)pre, d_out, width, height, scattering, scale, 1.0f, 1.0f, central_pixel_weight, norm, P,
                               K, 0, norm2, s)))
    return rc;
  return denoise_vst_backward(piece, d, d_out, npx, true, s);
}
} // namespace b200
