/* Test infrastructure: the stages of ansel_b200/csrc/markesteijn.cu compiled with g++ and run thread by thread on the CPU, tile after
 * tile, in the order the kernel runs them -- the threads of every stage in DESCENDING order, so that a stage whose pixels read what
 * the same stage writes elsewhere (there must be none) shows up as a difference to the oracle.  Not part of the product. */
#define B200_KERNELS_ON_CPU
#include "cuda_on_cpu.h"
#include "../../include/b200iop.h"
#include "../../ansel_b200/csrc/markesteijn.cu"
#include <vector>

template <int PASSES> static int emul_run(float *out, const float *in, int width, int height, int x0, int y0, const unsigned char *xtrans36, int nthreads, int ascending)
{
  using geo = mk_geo<PASSES>;
  _mm_setcsr(_mm_getcsr() | 0x8040u);
  mk_args_t a;
  memset(&a, 0, sizeof(a));
  a.width = width;
  a.height = height;
  a.pad = geo::PAD;
  a.step = geo::STEP;
  for(int r = 0; r < 6; r++)
    for(int c = 0; c < 6; c++) a.xt[r * 6 + c] = xtrans36[((r + y0 + 600) % 6) * 6 + (c + x0 + 600) % 6];
  mk_hexagons(a);
  a.ntx = (width + geo::STEP - 1) / geo::STEP;
  const int nty = (height + geo::STEP - 1) / geo::STEP;
  a.ntiles = a.ntx * nty;
  a.in = in;
  a.out = (float4 *)out;
  std::vector<float> P((size_t)geo::PLANES * NPX, __builtin_nanf(""));
  std::vector<uint8_t> homo(geo::NDIR * NPX, 0xff);
  std::vector<short> start(NPX);
  a.start = start.data();
  for(int t = 0; t < a.ntiles; t++)
  {
    mk_tile_t T = mk_tile_of(a, t);
    T.cls = 0; /* one map, rebuilt per tile: the classes of the product are checked by the GPU tests and test_cpu_markesteijn's class test */
    mk_walk(start.data(), a.xt, a.sgrow, T.top, T.left, T.nrow, T.ncol);
#define STAGE(call)                                                                                                    \
  for(int k = 0; k < nthreads; k++)                                                                                    \
  {                                                                                                                    \
    const int tid = ascending ? k : nthreads - 1 - k;                                                                  \
    call;                                                                                                              \
  }
    STAGE(mk_load(a, T, P.data(), tid, nthreads))
    STAGE(mk_green(a, T, P.data(), tid, nthreads))
    for(int pass = 0; pass < PASSES; pass++)
    {
      float *const Pp = pass ? P.data() + 12 * NPX : P.data();
      if(pass == 1) STAGE(mk_copy_planes(P.data(), tid, nthreads))
      if(pass) STAGE(mk_recalc_green(a, T, Pp, P.data() + NPX, tid, nthreads))
      STAGE(mk_solitary(a, T, Pp, geo::PAD_SG, tid, nthreads))
      STAGE(mk_red_blue(a, T, Pp, geo::PAD_RB, tid, nthreads))
      STAGE(mk_green_blocks<geo::NDIR / 2>(a, T, Pp, geo::PAD_G22, tid, nthreads))
    }
    STAGE(mk_derivatives<PASSES>(a, T, P.data(), tid, nthreads))
    STAGE(mk_homogeneity<PASSES>(a, T, P.data(), homo.data(), tid, nthreads))
    STAGE(mk_average<PASSES>(a, T, P.data(), homo.data(), tid, nthreads))
#undef STAGE
  }
  return 0;
}

extern "C" int emul_markesteijn(float *out, const float *in, int width, int height, int x0, int y0, const unsigned char *xtrans36, int nthreads, int ascending, int passes)
{
  return passes == 3 ? emul_run<3>(out, in, width, height, x0, y0, xtrans36, nthreads, ascending) : emul_run<1>(out, in, width, height, x0, y0, xtrans36, nthreads, ascending);
}

/* the classes of tiles the product builds for a frame: the class of every tile and, per class, the record of the walk -- compared by the
 * test with a walk of each tile on its own.  Returns the number of classes, -1 if two tiles of one class walk differently. */
extern "C" int emul_markesteijn_classes(int width, int height, int x0, int y0, const unsigned char *xtrans36, int passes)
{
  mk_args_t a;
  memset(&a, 0, sizeof(a));
  a.width = width;
  a.height = height;
  a.pad = passes == 3 ? mk_geo<3>::PAD : mk_geo<1>::PAD;
  a.step = passes == 3 ? mk_geo<3>::STEP : mk_geo<1>::STEP;
  for(int r = 0; r < 6; r++)
    for(int c = 0; c < 6; c++) a.xt[r * 6 + c] = xtrans36[((r + y0 + 600) % 6) * 6 + (c + x0 + 600) % 6];
  mk_hexagons(a);
  std::vector<short> maps;
  if(mk_build_classes(a, maps)) return -2;
  std::vector<short> own(NPX);
  int unvisited = 0;
  for(int t = 0; t < a.ntiles; t++)
  {
    const mk_tile_t T = mk_tile_of(a, t);
    mk_walk(own.data(), a.xt, a.sgrow, T.top, T.left, T.nrow, T.ncol);
    if(memcmp(own.data(), maps.data() + (size_t)T.cls * NPX, NPX * sizeof(short))) return -1;
    for(int r = 3; r < T.nrow - 3; r++)
      for(int c = 3; c < T.ncol - 3; c++)
        if(mk_fc(a.xt, T.top + r, T.left + c) != 1 && own[r * TS + c] < 0) unvisited++;
  }
  return unvisited ? -3 : (int)(maps.size() / NPX);
}
