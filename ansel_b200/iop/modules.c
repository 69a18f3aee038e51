/* C module adapters: the bodies a reference maintainer puts behind each module's
 * process() / process_cl() / tiling_callback() (src/iop/iop_api.h:265-266, 292-293, 121-122).
 *
 * The reference prefixes a module's plain symbols with dt_iop_<op>__ through asm labels
 * (src/common/module_api.h:139-154); the same names are exported here so the adapters can be
 * exercised from tests exactly as lib_ansel would call them.  Everything else of each module
 * (params, commit_params, GUI, introspection) stays the reference's own C.
 *
 * Return conventions: process() 0 = success; process_cl() TRUE = success.
 */
#include "dt_surface.h"

#define ADAPT(op)                                                                                         \
  int dt_iop_##op##__process(struct dt_iop_module_t *self, const dt_dev_pixelpipe_t *pipe,                \
                             const dt_dev_pixelpipe_iop_t *piece, const void *const i, void *const o)     \
  {                                                                                                       \
    b200_piece_t p;                                                                                       \
    b200_piece_from_dt(&p, self, pipe, piece);                                                            \
    return b200_##op##_process_host(&p, i, o);                                                            \
  }                                                                                                       \
  int dt_iop_##op##__process_cl(struct dt_iop_module_t *self, const dt_dev_pixelpipe_t *pipe,             \
                                const dt_dev_pixelpipe_iop_t *piece, cl_mem dev_in, cl_mem dev_out)       \
  {                                                                                                       \
    b200_piece_t p;                                                                                       \
    b200_piece_from_dt(&p, self, pipe, piece);                                                            \
    return b200_##op##_process_dev(&p, dev_in, dev_out, pipe->stream) == 0 ? TRUE : FALSE;                \
  }                                                                                                       \
  void dt_iop_##op##__tiling_callback(struct dt_iop_module_t *self, const dt_dev_pixelpipe_t *pipe,       \
                                      const dt_dev_pixelpipe_iop_t *piece, dt_develop_tiling_t *tiling)   \
  {                                                                                                       \
    b200_piece_t p;                                                                                       \
    b200_piece_from_dt(&p, self, pipe, piece);                                                            \
    b200_##op##_tiling(&p, tiling);                                                                       \
  }

ADAPT(demosaic) /* src/iop/demosaic.c:1043 (process), rcd.c:568 (process_rcd_cl), :1916 (tiling_callback) */
ADAPT(colorin)  /* src/iop/colorin.c:711, :590-681 (process_cl) */
ADAPT(colorout) /* src/iop/colorout.c:373, :288-371 (process_cl) */

/* layout probes so non-C callers (tests, bench.py) can verify their mirror of dt_surface.h */
#include <stddef.h>
size_t b200_dt_surface_probe(int which)
{
  switch(which)
  {
    case 0: return sizeof(dt_iop_buffer_dsc_t);
    case 1: return offsetof(dt_iop_buffer_dsc_t, temperature.coeffs);
    case 2: return offsetof(dt_iop_buffer_dsc_t, processed_maximum);
    case 3: return offsetof(dt_iop_buffer_dsc_t, cst);
    case 4: return sizeof(dt_dev_pixelpipe_iop_t);
    case 5: return offsetof(dt_dev_pixelpipe_iop_t, roi_in);
    case 6: return offsetof(dt_dev_pixelpipe_iop_t, dsc_in);
    case 7: return offsetof(dt_dev_pixelpipe_iop_t, dsc_out);
    case 8: return sizeof(dt_dev_pixelpipe_t);
    case 9: return sizeof(dt_iop_module_t);
    case 10: return sizeof(b200_piece_t);
    case 11: return sizeof(b200_conversion_t);
    default: return 0;
  }
}
