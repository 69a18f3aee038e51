// Row-band planner for sharding ONE frame over several GPUs.  Host code only.
//
// The reference shards a frame through its tiling engine: src/develop/tiling.c
// _default_process_tiling_ptp :723-1075 cuts tiles with the overlap/alignment a module's
// tiling_callback() reports, runs process() on each tile with roi_in.x/y set to the tile origin, and
// keeps the tile interior.  A band here is exactly such a tile that spans the full width.
//
// Two ways to place the cuts:
//  * overlap mode (grid = 1): what tiling.c does -- the result equals the reference's own tiled result
//    for the same cuts (not its untiled one wherever a module's arithmetic depends on the tile origin);
//  * grid mode: cuts on the module's internal block grid so the banded result is bit-identical to the
//    UNTILED frame.  For RCD the grid is its 94-row tile pitch (iop/demosaic/rcd.c:71-75: RCD_TILESIZE 112,
//    RCD_BORDER 9, RCD_TILEVALID 94): a band whose input starts at a multiple of 94 and ends 18 rows past
//    one replays the same tiles as the full frame, and dropping 9 rows at each cut removes the rows its first /
//    last tile treated as frame edge.
#include "runtime.h"

using namespace b200;

extern "C" int b200_band_plan(int height, int n_bands, int grid, int halo, int align, b200_band_t *bands)
{
  if(!bands || height <= 0 || n_bands <= 0 || grid <= 0 || halo < 0 || align <= 0) return fail(B200_ERR_ARG, "band_plan: bad argument");
  if(grid > 1 && (grid % align)) return fail(B200_ERR_ARG, "band_plan: the block grid must be a multiple of the alignment");
  int cut[B200_MAX_BANDS + 1];
  if(n_bands > B200_MAX_BANDS) return fail(B200_ERR_ARG, "band_plan: more than B200_MAX_BANDS bands");
  cut[0] = 0;
  cut[n_bands] = height;
  for(int i = 1; i < n_bands; i++)
  {
    const long long ideal = (long long)height * i / n_bands;
    int c;
    if(grid > 1)
    { // nearest grid line, then the halo rows the previous band's last block does not own
      const int k = (int)((ideal + grid / 2) / grid);
      c = k * grid + halo;
    }
    else
      c = (int)(ideal / align) * align;
    if(c < cut[i - 1]) c = cut[i - 1];
    if(c > height) c = height;
    cut[i] = c;
  }
  for(int i = 0; i < n_bands; i++)
  {
    b200_band_t *b = bands + i;
    b->out_y0 = cut[i];
    b->out_y1 = cut[i + 1] < cut[i] ? cut[i] : cut[i + 1];
    if(b->out_y1 == b->out_y0)
    { // empty band (more GPUs than block rows): nothing to read
      b->in_y0 = b->in_y1 = b->out_y0;
      continue;
    }
    int y0 = b->out_y0 - halo, y1 = b->out_y1 + halo;
    if(grid <= 1)
    { // tiling.c aligns the tile origin down, the end up
      y0 = (y0 / align) * align;
      y1 = ((y1 + align - 1) / align) * align;
    }
    b->in_y0 = y0 < 0 ? 0 : y0;
    b->in_y1 = y1 > height ? height : y1;
  }
  return B200_OK;
}

// the block grid and halo under which banded RCD equals the untiled frame bit for bit (rcd.c:71-75)
extern "C" void b200_demosaic_band_grid(const b200_piece_t *piece, int *grid, int *halo, int *align)
{
  if(piece && piece->data)
  {
    const b200_demosaic_data_t *d = (const b200_demosaic_data_t *)piece->data;
    if((d->demosaicing_method & ~2048u) != B200_DEMOSAIC_RCD) /* 2048 = DEMOSAIC_DUAL: the VNG4 half blends in pointwise */
    { // AMaZE mirrors at its own tile origin: no cut reproduces the untiled frame; bands are tiling.c tiles
      b200_tiling_t t;
      b200_demosaic_tiling(piece, &t);
      if(grid) *grid = 1;
      if(halo) *halo = (int)t.overlap;
      if(align) *align = (int)t.yalign;
      return;
    }
  }
  if(grid) *grid = 94;
  if(halo) *halo = 9;
  if(align) *align = 2;
}
