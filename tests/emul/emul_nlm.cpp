/* Test infrastructure: the phases of the non-local-means group kernel (ansel_b200/csrc/nlm_group.cuh) compiled with g++
 * and run thread by thread on the CPU, phase after phase as the kernel's barriers order them, with the product's own
 * host-side planning (grp_plan, grp_define_patches).  Not part of the product. */
#define B200_KERNELS_ON_CPU
#include "cuda_on_cpu.h"
#include "../../include/b200iop.h"
#include "../../ansel_b200/csrc/nlm.cu"
#include <vector>

template <int R, int WP, bool NORM1, bool PROFILED, bool DIVC, int KP> static void run_chunks(const grp_args_t &a, int n_chunks)
{
  std::vector<float> smem((size_t)a.wrows * 3 * WP + (size_t)a.G * a.splane + GRP_MAXG);
  std::vector<grp_thread_t<KP>> st(GRP_NT);
  float *const W = smem.data(), *const S = W + a.wrows * 3 * WP;
  int *const shifts = reinterpret_cast<int *>(S + a.G * a.splane);
  for(int b = 0; b < n_chunks; b++)
  {
    for(auto &v : smem) v = __builtin_nanf(""); // whatever a phase reads must have been written by an earlier one
    const chunk_t c = chunk_of(a, b);
    for(int t = 0; t < GRP_NT; t++) grp_fill<WP>(a, c, W, t);
    for(int t = 0; t < GRP_NT; t++) grp_own_init<WP>(a, c, W, st[t], t);
    for(int p0 = 0; p0 < a.n_patches; p0 += a.G)
    {
      for(int t = 0; t < GRP_NT; t++) grp_phase_a<WP, R, NORM1>(a, c, W, S, shifts, p0, t);
      for(int t = 0; t < GRP_NT; t++) grp_phase_b1<R>(a, c, S, p0, t);
      for(int t = 0; t < GRP_NT; t++) grp_phase_b2<WP, PROFILED, DIVC, KP>(a, c, W, S, shifts, st[t], p0, t);
    }
    for(int t = 0; t < GRP_NT; t++) grp_finish(a, c, st[t], t);
  }
}

template <int R, int WP, int KP> static void run_rk(const grp_args_t &a, int n, bool norm1, bool profiled, bool divc)
{
  if(!profiled)
    return norm1 ? run_chunks<R, WP, true, false, false, KP>(a, n) : run_chunks<R, WP, false, false, false, KP>(a, n);
  if(divc) return norm1 ? run_chunks<R, WP, true, true, true, KP>(a, n) : run_chunks<R, WP, false, true, true, KP>(a, n);
  return norm1 ? run_chunks<R, WP, true, true, false, KP>(a, n) : run_chunks<R, WP, false, true, false, KP>(a, n);
}
template <int R> static void run_r(const grp_args_t &a, int n, bool norm1, bool profiled, bool divc)
{
  const bool tall = grp_pairs_per_thread(a) > GRP_KP_MIN;
  if(a.wp == GRP_WP_NARROW)
    return tall ? run_rk<R, GRP_WP_NARROW, GRP_KP_MAX>(a, n, norm1, profiled, divc) : run_rk<R, GRP_WP_NARROW, GRP_KP_MIN>(a, n, norm1, profiled, divc);
  return tall ? run_rk<R, GRP_WP_WIDE, GRP_KP_MAX>(a, n, norm1, profiled, divc) : run_rk<R, GRP_WP_WIDE, GRP_KP_MIN>(a, n, norm1, profiled, divc);
}

/* same arguments as b200_nlmeans_denoise_dev on host buffers; smem_bytes = what an SM offers, g_cap = patches in flight at most,
 * ieee_div: the plain division instead of Markstein's sequence.  Returns the patches in flight, 0 = the plan does not fit. */
extern "C" int emul_nlmeans_group(const float *in, float *out, int width, int height, float scattering, float scale, float luma, float chroma,
                                  float center_weight, float sharpness, int radius, int search_radius, int decimate, const float *norm,
                                  int smem_bytes, int g_cap, int ieee_div)
{
  _mm_setcsr(_mm_getcsr() | 0x8040u);
  int n_patches = (2 * search_radius + 1) * (2 * search_radius + 1);
  if(decimate) n_patches = (n_patches + 1) / 2;
  std::vector<patch_t> patches(n_patches);
  const int shift_max = grp_define_patches(patches.data(), search_radius, scale, scattering, decimate);
  const float weight[4] = { luma, chroma, chroma, 1.0f }, invert[4] = { 1.0f - luma, 1.0f - chroma, 1.0f - chroma, 0.0f };
  const int pw = 2 * radius + 1;
  (void)pw;
  grp_args_t g;
  g.in = (const float4 *)in;
  g.out = (float4 *)out;
  g.patches = patches.data();
  if(!grp_plan(g, n_patches, width, height, radius, center_weight, sharpness, norm, weight, invert, (luma == 1.0 && chroma == 1.0) ? 1 : 0,
               shift_max, smem_bytes, g_cap))
    return 0;
  const int n_ct = (height + g.chk_h - 1) / g.chk_h;
  const bool profiled = !(center_weight < 0), norm1 = norm[0] == 1.0f && norm[1] == 1.0f && norm[2] == 1.0f;
  const bool divc = grp_division_by_constant(g) && !ieee_div;
  if(radius == 1)
    run_r<1>(g, n_ct * g.n_cl, norm1, profiled, divc);
  else
    run_r<2>(g, n_ct * g.n_cl, norm1, profiled, divc);
  return g.G;
}
