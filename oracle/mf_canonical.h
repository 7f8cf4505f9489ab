/* TEST INFRASTRUCTURE -- canonical matched filter used by the oracle.
 *
 * The reference instantiates GNU Radio's own filter.fir_filter_ccc(5,[1]*25)
 * (reference: gr-rfid/apps/reader.py:65,75); its source (gr-filter + VOLK
 * volk_32fc_x2_dot_prod_32fc, GNU Radio >= 3.7.2, gr-rfid/CMakeLists.txt:93-95)
 * is NOT under /root/reference and its float summation order depends on the
 * SIMD kernel VOLK selects at run time.  No reference test pins that order, so
 * at the bit level this stage is "parity unpinned" and the build DEFINES the
 * canonical order here:
 *
 *   y[n] = ((..((0 + x[D*n-(K-1)]) + x[D*n-(K-2)]) + ...) + x[D*n]),   x[<0] = +0
 *
 * i.e. one float accumulator per component, starting from +0, input index
 * ascending (the same direction GNU Radio's fir_filter walks its reversed-tap
 * dot product), taps all exactly 1.0 so the products are exact.  Output count is
 * floor(n_in / D) (an output needs its newest sample x[D*n] to exist).
 */
#ifndef ORACLE_MF_CANONICAL_H
#define ORACLE_MF_CANONICAL_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* x: interleaved I,Q float32, n_in complex samples. y: interleaved, capacity n_in/decim.
 * returns number of outputs written. */
static inline size_t oracle_mf_boxcar(const float* x, size_t n_in, int ntaps, int decim, float* y)
{
  size_t n_out = n_in / (size_t)decim;
  for (size_t n = 0; n < n_out; n++) {
    float ar = 0.0f, ai = 0.0f;
    long newest = (long)(n * (size_t)decim);
    for (long k = newest - (ntaps - 1); k <= newest; k++) {
      float xr = 0.0f, xi = 0.0f;
      if (k >= 0) { xr = x[2 * k]; xi = x[2 * k + 1]; }
      ar = ar + xr;
      ai = ai + xi;
    }
    y[2 * n] = ar;
    y[2 * n + 1] = ai;
  }
  return n_out;
}

#ifdef __cplusplus
}
#endif
#endif
