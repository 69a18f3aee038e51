/* Test infrastructure: the kernels of ansel_b200/csrc/bilateral.cu compiled with g++ and run thread by thread on the CPU,
 * in the order bilateral_grid_dev() launches them.  Not part of the product. */
#define B200_KERNELS_ON_CPU
#include "cuda_on_cpu.h"
#include "../../include/b200iop.h"
#include "../../ansel_b200/csrc/bilateral.cu"
#include <vector>

/* grid != NULL: the grid after the splat (blur == 0) or after the three blurs */
extern "C" int emul_bilateral(const float *in, float *out, int width, int height, float sigma_s, float sigma_r, float detail, float *grid, int max_floats,
                              int dims[3], int blur)
{
  bgrid_t G;
  bilateral_grid_size(&G, width, height, 100.0f, sigma_s, sigma_r);
  if(G.size_z > MAX_Z || G.size_z < 4 || G.size_x < 4 || G.size_y < 4) return 3;
  const size_t cells = (size_t)G.size_x * G.size_y, n = cells * G.size_z;
  std::vector<float> buf(n);
  emulate(dim3((unsigned)((cells + BNT - 1) / BNT)), BNT, bilateral_splat_kernel, (const float4 *)in, buf.data(), G);
  if(grid && !blur && n <= (size_t)max_floats) memcpy(grid, buf.data(), n * sizeof(float));
  const size_t ox = (size_t)G.size_z, oy = (size_t)G.size_x * G.size_z, oz = 1;
  emulate(dim3((unsigned)(((size_t)G.size_z * G.size_y + BNT - 1) / BNT)), BNT, bilateral_blur_kernel<false>, buf.data(), oz, oy, ox, G.size_z, G.size_y, G.size_x);
  emulate(dim3((unsigned)(((size_t)G.size_z * G.size_x + BNT - 1) / BNT)), BNT, bilateral_blur_kernel<false>, buf.data(), oz, ox, oy, G.size_z, G.size_x, G.size_y);
  emulate(dim3((unsigned)(((size_t)G.size_x * G.size_y + BNT - 1) / BNT)), BNT, bilateral_blur_kernel<true>, buf.data(), ox, oy, oz, G.size_x, G.size_y, G.size_z);
  if(grid && blur && n <= (size_t)max_floats) memcpy(grid, buf.data(), n * sizeof(float));
  if(dims) dims[0] = G.size_x, dims[1] = G.size_y, dims[2] = G.size_z;
  emulate(dim3((unsigned)((width + BNT - 1) / BNT), (unsigned)height), BNT, bilateral_slice_kernel, (const float4 *)in, (float4 *)out, (const float *)buf.data(), G,
          -detail * G.sigma_r * 0.04f);
  return 0;
}
