/* oracle/_ref wrapper: the float -> integer conversions at the end of the pipe.  TEST INFRASTRUCTURE ONLY.
 * oracle/Makefile cuts verbatim: iop/gamma.c :352-364 (_copy_output, what gamma's process() runs when no mask or channel
 * is displayed, :367-377); imageio/imageio_core.c :706-738 (_clamp_float_to_uint8, _swap_byteorder_float_to_uint8,
 * _export_final_buffer_to_uint16: the export's down-conversion of the pipe's float backbuffer). */
#include "ref_piece.h"
#include "gen_pipe_end.c"

void ref_gamma_copy_output(const float *in, uint8_t *out, size_t npixels) { _copy_output(in, out, npixels * 4); }
void ref_clamp_float_to_uint8(const float *in, uint8_t *out, size_t w, size_t h) { _clamp_float_to_uint8(in, out, w, h); }
void ref_swap_byteorder_float_to_uint8(const float *in, uint8_t *out, size_t w, size_t h) { _swap_byteorder_float_to_uint8(in, out, w, h); }
void ref_export_final_buffer_to_uint16(const float *in, uint16_t *out, size_t w, size_t h) { _export_final_buffer_to_uint16(in, out, w, h); }
