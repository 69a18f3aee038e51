/* CPU restatement of the demosaic module's optional passes around the demosaicer.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/src/iop/demosaic/basic.c: color_smoothing :192-245 (median of nine colour differences
 * through the fixed 19-exchange network, the alpha lane used as scratch), green_equilibration_lavg :248-293,
 * green_equilibration_favg :296-329, as called from iop/demosaic.c process() :1137-1170 (threshold :1049) and
 * :1249-1250.  Pinned against those lines cut verbatim (oracle/_ref, ref_demosaic_extra.c): bit-exact for the
 * smoothing and the local average; the full average sums two greens planes with an OpenMP `reduction(+)` in double,
 * whose order the reference does not define -- the oracle sums in raster order.
 */
#include "oracle_common.h"
#include <string.h>

#define SWAPMED(I, J)          \
  if(med[I] > med[J])          \
  {                            \
    const float tmp = med[J];  \
    med[J] = med[I];           \
    med[I] = tmp;              \
  }

void orc_color_smoothing(float *out, int width, int height, int passes)
{
  const int w4 = 4 * width;
  for(int pass = 0; pass < passes; pass++)
    for(int c = 0; c < 3; c += 2)
    {
      for(size_t k = 0; k < (size_t)width * height; k++) out[4 * k + 3] = out[4 * k + c];
#pragma omp parallel for
      for(int j = 1; j < height - 1; j++)
      {
        float *outp = out + (size_t)4 * j * width + 4;
        for(int i = 1; i < width - 1; i++, outp += 4)
        {
          float med[9] = { outp[-w4 - 4 + 3] - outp[-w4 - 4 + 1], outp[-w4 + 3] - outp[-w4 + 1], outp[-w4 + 4 + 3] - outp[-w4 + 4 + 1],
                           outp[-4 + 3] - outp[-4 + 1],           outp[3] - outp[1],             outp[4 + 3] - outp[4 + 1],
                           outp[w4 - 4 + 3] - outp[w4 - 4 + 1],   outp[w4 + 3] - outp[w4 + 1],   outp[w4 + 4 + 3] - outp[w4 + 4 + 1] };
          SWAPMED(1, 2) SWAPMED(4, 5) SWAPMED(7, 8) SWAPMED(0, 1) SWAPMED(3, 4) SWAPMED(6, 7) SWAPMED(1, 2) SWAPMED(4, 5) SWAPMED(7, 8)
          SWAPMED(0, 3) SWAPMED(5, 8) SWAPMED(4, 7) SWAPMED(3, 6) SWAPMED(1, 4) SWAPMED(2, 5) SWAPMED(4, 7) SWAPMED(4, 2) SWAPMED(6, 4)
          SWAPMED(4, 2)
          outp[c] = fmaxf(med[4] + outp[1], 0.0f);
        }
      }
    }
}

void orc_green_eq_lavg(float *out, const float *in, int width, int height, uint32_t filters, int x, int y, float thr)
{
  const float maximum = 1.0f;
  int oj = 2, oi = 2;
  if(orc_fc(oj + y, oi + x, filters) != 1) oj++;
  if(orc_fc(oj + y, oi + x, filters) != 1) oi++;
  if(orc_fc(oj + y, oi + x, filters) != 1) oj--;
  memcpy(out, in, sizeof(float) * (size_t)width * height);
  for(size_t j = oj; j + 2 < (size_t)height; j += 2)
    for(size_t i = oi; i + 2 < (size_t)width; i += 2)
    {
      const float o1_1 = in[(j - 1) * width + i - 1], o1_2 = in[(j - 1) * width + i + 1];
      const float o1_3 = in[(j + 1) * width + i - 1], o1_4 = in[(j + 1) * width + i + 1];
      const float o2_1 = in[(j - 2) * width + i], o2_2 = in[(j + 2) * width + i];
      const float o2_3 = in[j * width + i - 2], o2_4 = in[j * width + i + 2];
      const float m1 = (o1_1 + o1_2 + o1_3 + o1_4) / 4.0f;
      const float m2 = (o2_1 + o2_2 + o2_3 + o2_4) / 4.0f;
      if((m2 > 0.0f) && (m1 > 0.0f) && (m1 / m2 < maximum * 2.0f))
      {
        const float c1 = (fabsf(o1_1 - o1_2) + fabsf(o1_1 - o1_3) + fabsf(o1_1 - o1_4) + fabsf(o1_2 - o1_3) + fabsf(o1_3 - o1_4)
                          + fabsf(o1_2 - o1_4)) / 6.0f;
        const float c2 = (fabsf(o2_1 - o2_2) + fabsf(o2_1 - o2_3) + fabsf(o2_1 - o2_4) + fabsf(o2_2 - o2_3) + fabsf(o2_3 - o2_4)
                          + fabsf(o2_2 - o2_4)) / 6.0f;
        if((in[j * width + i] < maximum * 0.95f) && (c1 < maximum * thr) && (c2 < maximum * thr))
          out[j * width + i] = in[j * width + i] * m1 / m2;
      }
    }
}

void orc_green_eq_favg(float *out, const float *in, int width, int height, uint32_t filters, int x, int y)
{
  int oj = 0, oi = 0;
  double sum1 = 0.0, sum2 = 0.0, gr_ratio;
  if((orc_fc(oj + y, oi + x, filters) & 1) != 1) oi++;
  const int g2_offset = oi ? -1 : 1;
  memcpy(out, in, sizeof(float) * (size_t)width * height);
  for(size_t j = oj; j + 1 < (size_t)height; j += 2)
    for(size_t i = oi; (long)i < (long)(width - 1 - g2_offset); i += 2)
    {
      sum1 += in[j * width + i];
      sum2 += in[(j + 1) * width + i + g2_offset];
    }
  if(sum1 > 0.0 && sum2 > 0.0)
    gr_ratio = sum2 / sum1;
  else
    return;
  for(int j = oj; j < height - 1; j += 2)
    for(int i = oi; i < width - 1 - g2_offset; i += 2) out[(size_t)j * width + i] = in[(size_t)j * width + i] * gr_ratio;
}
