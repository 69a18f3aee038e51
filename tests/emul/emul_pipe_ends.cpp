/* Test infrastructure: the kernels of ansel_b200/csrc/pipe_ends.cu compiled with g++ and run thread by thread on the
 * CPU, launched the way the b200_*_process_dev entry points of that file launch them.  Not part of the product. */
#define B200_KERNELS_ON_CPU
#include "cuda_on_cpu.h"
#include "../../include/b200iop.h"
#include "../../ansel_b200/csrc/pipe_ends.cu"
#include <vector>

static dim3 row_grid(int width, int height) { return dim3((unsigned)((width + 4 * NT - 1) / (4 * NT)), (unsigned)height); }
static unsigned flat_grid(size_t n) { return (unsigned)((n + 4 * NT - 1) / (4 * NT)); }

template <int STAGES, bool COUNT>
static void front(bool u16, const void *in, float *out, int w, int h, const prepare_t &P, const balance_t &B, const clip_t &H, unsigned long long *counter)
{
  if(u16)
    emulate(row_grid(w, h), NT, raw_front_kernel<unsigned short, STAGES, COUNT>, (const unsigned short *)in, out, w, h, P, B, H, counter);
  else
    emulate(row_grid(w, h), NT, raw_front_kernel<float, STAGES, COUNT>, (const float *)in, out, w, h, P, B, H, counter);
}

/* the gain maps as make_prepare() lays them out on the device: four planes */
static int prepare(const b200_piece_t *piece, prepare_t *P, std::vector<float> &maps)
{
  const int rc = fill_prepare(piece, P);
  if(rc) return rc;
  if(P->gain)
  {
    const b200_rawprepare_data_t *d = (const b200_rawprepare_data_t *)piece->data;
    const size_t plane = (size_t)P->map_w * P->map_h;
    maps.resize(4 * plane);
    for(int f = 0; f < 4; f++) memcpy(maps.data() + f * plane, d->gainmaps[f]->map_gain, plane * sizeof(float));
    P->maps = maps.data();
  }
  return 0;
}

extern "C" int emul_rawprepare(const b200_piece_t *piece, const void *in, float *out)
{
  const b200_rawprepare_data_t *d = (const b200_rawprepare_data_t *)piece->data;
  const int w = piece->roi_out.width, h = piece->roi_out.height;
  prepare_t P;
  std::vector<float> maps;
  if(prepare(piece, &P, maps)) return 1;
  if(bayer_typed(piece))
  {
    front<1, false>(piece->datatype == B200_TYPE_UINT16, in, out, w, h, P, balance_t{}, clip_t{}, nullptr);
    return 0;
  }
  const size_t n = (size_t)w * h * piece->channels;
  emulate(dim3((unsigned)((n + NT - 1) / NT)), NT, predownsampled_kernel, (const float *)in, out, w, h, (int)piece->channels, piece->roi_in.width, P.csx, P.csy,
          d->sub[0], d->div[0]);
  return 0;
}

extern "C" int emul_temperature(const b200_piece_t *piece, const float *in, float *out)
{
  const b200_temperature_data_t *d = (const b200_temperature_data_t *)piece->data;
  const int w = piece->roi_out.width, h = piece->roi_out.height;
  if(piece->filters == 9u)
  {
    float lut[36];
    for(int r = 0; r < 6; r++)
      for(int c = 0; c < 6; c++) lut[r * 6 + c] = d->coeffs[piece->xtrans[r][c]];
    emulate(dim3((unsigned)((w + NT - 1) / NT), (unsigned)h), NT, balance_xtrans_kernel, in, out, w, h, piece->roi_out.x, piece->roi_out.y, (const float *)lut);
  }
  else if(piece->filters)
  {
    balance_t B;
    make_balance(piece, &B);
    front<2, false>(false, in, out, w, h, prepare_t{}, B, clip_t{}, nullptr);
  }
  else
  {
    const size_t npx = (size_t)w * h;
    emulate(dim3((unsigned)((npx + NT - 1) / NT)), NT, balance_pixels_kernel, in, out, npx, (int)piece->channels, d->coeffs[0], d->coeffs[1], d->coeffs[2]);
  }
  return 0;
}

/* -> 0 ok, 3 = a reconstruction mode past its bypass (B200_ERR_UNSUPPORTED) */
/* shifted_filters: b200_roi_filters(piece->filters, roi_in.x, roi_in.y), a host function of the product the caller evaluates */
// the launch sequence of inpaint_sequence() (pipe_ends.cu); the planes start as NaN: only what a direction wrote may reach the output
static void emul_inpaint(const float *in, float *out, const inpaint_t &A, const xtable_t *X, const unsigned long long *cnt)
{
  const int w = A.width, h = A.height;
  const size_t npx = (size_t)w * h;
  std::vector<float> t_in(npx, NAN), t_fwd(npx, NAN), t_bwd(npx, NAN), up(npx, NAN);
  float *rows = t_in.data(), *down = out;
  const dim3 by_row((unsigned)((h + 127) / 128), 2), by_col((unsigned)((w + 127) / 128), 2), px((unsigned)((w + NT - 1) / NT), (unsigned)h);
  emulate(dim3((unsigned)((w + 31) / 32), (unsigned)((h + 31) / 32)), 256, transpose_kernel<false>, in, (const float *)nullptr, t_in.data(), w, h, cnt);
  if(X)
  {
    emulate(by_row, 128, inpaint_rows_xtrans_kernel, (const float *)t_in.data(), t_fwd.data(), t_bwd.data(), A, *X, cnt);
    emulate(by_col, 128, inpaint_cols_xtrans_kernel, in, down, up.data(), A, *X, cnt);
  }
  else
  {
    emulate(by_row, 128, inpaint_rows_kernel, (const float *)t_in.data(), t_fwd.data(), t_bwd.data(), A, cnt);
    emulate(by_col, 128, inpaint_cols_kernel, in, down, up.data(), A, cnt);
  }
  emulate(dim3((unsigned)((h + 31) / 32), (unsigned)((w + 31) / 32)), 256, transpose_kernel<true>, (const float *)t_fwd.data(), (const float *)t_bwd.data(), rows, h, w, cnt);
  if(X)
    emulate(px, NT, inpaint_sum_xtrans_kernel, in, (const float *)rows, (const float *)down, (const float *)up.data(), out, A, *X, cnt);
  else
    emulate(px, NT, inpaint_sum_kernel, in, (const float *)rows, (const float *)down, (const float *)up.data(), out, A, cnt);
}

extern "C" int emul_highlights(const b200_piece_t *piece, const float *in, float *out, unsigned long long *n_clipped, unsigned shifted_filters)
{
  const b200_highlights_data_t *d = (const b200_highlights_data_t *)piece->data;
  const int w = piece->roi_out.width, h = piece->roi_out.height;
  const bool mosaic = piece->filters != 0;
  const int ch = mosaic ? 1 : (int)piece->channels;
  float thresholds[4], clip;
  make_thresholds(piece, thresholds, &clip);
  const size_t npx = (size_t)w * h, n = npx * ch;
  unsigned long long counter = 0;
  if(mosaic)
    front<0, true>(false, in, nullptr, w, h, prepare_t{}, balance_t{}, clip_t{ clip, fminf(fminf(thresholds[0], thresholds[1]), thresholds[2]) }, &counter);
  else
    emulate(dim3((unsigned)((npx + NT - 1) / NT)), NT, count_pixels_kernel, in, npx, ch, thresholds[0], thresholds[1], thresholds[2], &counter);
  *n_clipped = counter;
  if(mosaic && piece->filters == 9u && (d->mode == B200_HIGHLIGHTS_LCH || d->mode == B200_HIGHLIGHTS_INPAINT))
  {
    xtable_t X;
    for(int r = 0; r < 6; r++)
      for(int c = 0; c < 6; c++) X.v[r * 6 + c] = piece->xtrans[r][c];
    X.x0 = piece->roi_in.x;
    X.y0 = piece->roi_in.y;
    if(d->mode == B200_HIGHLIGHTS_LCH)
      emulate(dim3((unsigned)((w + NT - 1) / NT), (unsigned)h), NT, lch_xtrans_kernel, in, out, w, h, X, clip, (const unsigned long long *)&counter);
    else
    {
      float pmax[4];
      for(int c = 0; c < 4; c++) pmax[c] = (piece->processed_maximum[c] > 0.f) ? piece->processed_maximum[c] : 1.0f;
      const inpaint_t A = { { 0.987f * d->clip * pmax[0], 0.987f * d->clip * pmax[1], 0.987f * d->clip * pmax[2], clip }, 9u, w, h };
      emul_inpaint(in, out, A, &X, (const unsigned long long *)&counter);
    }
    return 0;
  }
  if(mosaic && piece->filters != 9u && d->mode == B200_HIGHLIGHTS_LCH)
  {
    emulate(dim3((unsigned)((w + NT - 1) / NT), (unsigned)h), NT, lch_bayer_kernel, in, out, w, h, piece->roi_out.x, piece->roi_out.y, (unsigned)piece->filters, clip,
            (const unsigned long long *)&counter);
    return 0;
  }
  if(mosaic && piece->filters != 9u && d->mode == B200_HIGHLIGHTS_INPAINT)
  {
    float pmax[4];
    for(int c = 0; c < 4; c++) pmax[c] = (piece->processed_maximum[c] > 0.f) ? piece->processed_maximum[c] : 1.0f;
    inpaint_t A = { { 0.987f * d->clip * pmax[0], 0.987f * d->clip * pmax[1], 0.987f * d->clip * pmax[2], clip }, shifted_filters, w, h };
    emul_inpaint(in, out, A, nullptr, (const unsigned long long *)&counter);
    return 0;
  }
  const bool clip_mode = d->mode == B200_HIGHLIGHTS_CLIP || (!mosaic && (d->mode == B200_HIGHLIGHTS_LCH || d->mode == B200_HIGHLIGHTS_INPAINT));
  if(clip_mode)
    emulate(dim3(flat_grid(n)), NT, flat_kernel<OP_CLIP>, in, out, n, clip, 0.0f, (piece->mask_display & 1) ? 1 : 0, (const unsigned long long *)&counter);
  else if(counter >= 25ull)
    return 3;
  else
    emulate(dim3(flat_grid(n)), NT, flat_kernel<OP_COPY>, in, out, n, 0.0f, 0.0f, 0, (const unsigned long long *)&counter);
  return 0;
}

extern "C" int emul_rawfront(const b200_piece_t *rp, const b200_piece_t *tp, const b200_piece_t *hp, const void *in, float *out)
{
  const int w = rp->roi_out.width, h = rp->roi_out.height;
  prepare_t P;
  std::vector<float> maps;
  if(prepare(rp, &P, maps)) return 1;
  balance_t B = {};
  if(tp) make_balance(tp, &B);
  clip_t H = {};
  unsigned long long counter = 0;
  const bool u16 = rp->datatype == B200_TYPE_UINT16;
  if(hp)
  {
    float thresholds[4];
    make_thresholds(hp, thresholds, &H.clip);
    H.raw_threshold = fminf(fminf(thresholds[0], thresholds[1]), thresholds[2]);
    if(tp)
    {
      front<3, true>(u16, in, out, w, h, P, B, H, &counter);
      front<7, false>(u16, in, out, w, h, P, B, H, &counter);
    }
    else
    {
      front<1, true>(u16, in, out, w, h, P, B, H, &counter);
      front<5, false>(u16, in, out, w, h, P, B, H, &counter);
    }
  }
  else if(tp)
    front<3, false>(u16, in, out, w, h, P, B, H, &counter);
  else
    front<1, false>(u16, in, out, w, h, P, B, H, &counter);
  return 0;
}

extern "C" int emul_exposure(const b200_piece_t *piece, const float *in, float *out)
{
  const b200_exposure_data_t *d = (const b200_exposure_data_t *)piece->data;
  const size_t n = (size_t)piece->roi_out.width * piece->roi_out.height * piece->channels;
  emulate(dim3(flat_grid(n)), NT, flat_kernel<OP_EXPOSURE>, in, out, n, d->black, d->scale, (piece->mask_display & 1) ? 1 : 0, (const unsigned long long *)nullptr);
  return 0;
}
extern "C" void emul_gamma(const float *in, unsigned *out, size_t npx)
{
  emulate(dim3((unsigned)((npx + NT - 1) / NT)), NT, gamma_kernel, (const float4 *)in, out, npx);
}
extern "C" void emul_export(const float *in, void *out, size_t npx, int format)
{
  const dim3 grid((unsigned)((npx + NT - 1) / NT));
  if(format == B200_EXPORT_UINT8) emulate(grid, NT, export_kernel<B200_EXPORT_UINT8>, (const float4 *)in, out, npx);
  if(format == B200_EXPORT_UINT8_SWAP) emulate(grid, NT, export_kernel<B200_EXPORT_UINT8_SWAP>, (const float4 *)in, out, npx);
  if(format == B200_EXPORT_UINT16) emulate(grid, NT, export_kernel<B200_EXPORT_UINT16>, (const float4 *)in, out, npx);
}
