/* shadow of src/develop/imageop_math.h for the oracle/_ref build of amaze.cc: FC() :190-193 and MIN. */
#pragma once
#include <stddef.h>
#include <stdint.h>
static inline int FC(const size_t row, const size_t col, const uint32_t filters)
{
  return filters >> (((row << 1 & 14) + (col & 1)) << 1) & 3;
}
#ifndef MIN
#define MIN(a, b) (((a) < (b)) ? (a) : (b))
#endif
