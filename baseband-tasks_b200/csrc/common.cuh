// Shared host-side helpers of the C-ABI implementation files.
#pragma once
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/bbt_b200.h"
#include "kernels_fft.cuh"

namespace bbt {

// Error reporting (thread-local message, status code returned).
int fail(int code, const std::string& msg);
const std::string& last_error();
int check_launch(const char* what);

inline int ilog2(int64_t n) {
  int l = 0;
  while ((int64_t(1) << l) < n) ++l;
  return l;
}
inline bool is_pow2(int64_t n) { return n > 0 && (n & (n - 1)) == 0; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline bbt_stream_t as_stream(void* s) { return static_cast<bbt_stream_t>(s); }

// exp(-2 pi i m / denom), m < count, on the device (computed in float64).
cf* make_roots(int64_t count, double denom);
// Per-device table of the 2^log2n-th roots of unity, exp(-2 pi i m / 2^log2n).
const cf* twiddle_table(int log2n);

// Tuning knob ``key`` (bbt_tune_set / BBT_TUNE), or ``dflt`` if not set.
int tune(const char* key, int dflt);
void tune_set(const char* key, int value);

inline unsigned grid_for(int64_t total, int threads) {
  return (unsigned)std::max<int64_t>(
      1, std::min<int64_t>(ceil_div(total, threads), (int64_t)sm_count() * 16));
}

}  // namespace bbt

#define BBT_FOR_LOG2(L, F)                                                    \
  switch (L) {                                                                \
    case 1: F(1); break;                                                      \
    case 2: F(2); break;                                                      \
    case 3: F(3); break;                                                      \
    case 4: F(4); break;                                                      \
    case 5: F(5); break;                                                      \
    case 6: F(6); break;                                                      \
    case 7: F(7); break;                                                      \
    case 8: F(8); break;                                                      \
    case 9: F(9); break;                                                      \
    case 10: F(10); break;                                                    \
    case 11: F(11); break;                                                    \
    case 12: F(12); break;                                                    \
    case 13: F(13); break;                                                    \
    case 14: F(14); break;                                                    \
    default: break;                                                           \
  }
