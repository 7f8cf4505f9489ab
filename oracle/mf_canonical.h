/* TEST INFRASTRUCTURE -- canonical matched filter used by the oracle.
 *
 * The reference instantiates GNU Radio's own filter.fir_filter_ccc(5,[1]*25)
 * (reference: gr-rfid/apps/reader.py:65,75); its source (gr-filter + VOLK
 * volk_32fc_x2_dot_prod_32fc, GNU Radio >= 3.7.2, gr-rfid/CMakeLists.txt:93-95)
 * is NOT under /root/reference and its float summation order depends on the
 * SIMD kernel VOLK selects at run time.  No reference test pins that order, so
 * at the bit level this stage is "parity unpinned" and the build DEFINES the
 * canonical order here -- the polyphase "integrate-and-dump, then moving sum"
 * structure of a decimating boxcar (D = decimation, K = ntaps, q = K / D,
 * rem = K % D, x[<0] = +0.0f, no leading zero terms):
 *
 *   B(m) = (((x[D*m-D+1] + x[D*m-D+2]) + ...) + x[D*m])          one block of D raw samples
 *   P(m) = ((x[D*m-rem+1] + ...) + x[D*m])                        newest `rem` samples of block m
 *   y[n] = (((P(n-q) + B(n-q+1)) + B(n-q+2)) + ...) + B(n)        (P term only if rem > 0)
 *
 * i.e. float32 adds, one accumulator per component, ascending input index
 * inside a block and ascending blocks (GNU Radio's fir_filter also walks its
 * reversed-tap dot product in ascending input order; VOLK's SIMD kernels then
 * split it over 2/4/8 strided accumulators, so no order is "the" reference
 * order).  Taps are exactly 1.0, so the products are exact.  For the reference
 * configuration (D = 5, K = 25): five block sums of five samples each.
 * tests/test_oracle.py checks that the decoded bits, sync indices and T on
 * misc/data/file_source_test do not depend on this choice (sequential float,
 * blocked float and float64 accumulation give identical decode results).
 * Output count is floor(n_in / D) (an output needs its newest sample x[D*n]).
 */
#ifndef ORACLE_MF_CANONICAL_H
#define ORACLE_MF_CANONICAL_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* x: interleaved I,Q float32, n_in complex samples. y: interleaved, capacity n_in/decim.
 * returns number of outputs written. */
static inline void oracle_mf_partial(const float* x, long lo, long hi, float* sr, float* si)
{
  /* ((x[lo] + x[lo+1]) + ...) + x[hi], indices < 0 read as +0 */
  float ar = 0.0f, ai = 0.0f;
  for (long k = lo; k <= hi; k++) {
    float xr = 0.0f, xi = 0.0f;
    if (k >= 0) { xr = x[2 * k]; xi = x[2 * k + 1]; }
    if (k == lo) { ar = xr; ai = xi; }
    else { ar = ar + xr; ai = ai + xi; }
  }
  *sr = ar; *si = ai;
}

static inline size_t oracle_mf_boxcar(const float* x, size_t n_in, int ntaps, int decim, float* y)
{
  const long D = decim, q = ntaps / decim, rem = ntaps % decim;
  size_t n_out = n_in / (size_t)decim;
  for (size_t n = 0; n < n_out; n++) {
    float ar = 0.0f, ai = 0.0f, br, bi;
    int first = 1;
    long nn = (long)n;
    if (rem > 0) {
      oracle_mf_partial(x, D * (nn - q) - rem + 1, D * (nn - q), &ar, &ai);
      first = 0;
    }
    for (long m = nn - q + 1; m <= nn; m++) {
      oracle_mf_partial(x, D * m - D + 1, D * m, &br, &bi);
      if (first) { ar = br; ai = bi; first = 0; }
      else { ar = ar + br; ai = ai + bi; }
    }
    y[2 * n] = ar;
    y[2 * n + 1] = ai;
  }
  return n_out;
}


/* Same values, bit for bit, as oracle_mf_boxcar, computed the way a FIR implementation would: every block sum B(m) is
 * formed once (the plain form above recomputes each of them ntaps/decim times) and the outputs are sums of q
 * neighbouring block sums in the canonical ascending order; both loops are simple enough for the compiler to
 * vectorise ACROSS outputs, which leaves the order of the additions inside each output untouched.  This is what the
 * CPU reference arm of bench.py times (a fair stand-in for GNU Radio's VOLK-vectorised fir_filter_ccc, whose source is
 * not in the reference tree); tests/test_oracle.py checks it against oracle_mf_boxcar.
 * scratch: 2 * (n_in / decim + q + 1) floats (block sums, interleaved). */
#if defined(__x86_64__) && defined(__GNUC__)
#define ORACLE_MF_TARGET __attribute__((target("avx2")))   /* additions only: no contraction can occur */
#else
#define ORACLE_MF_TARGET
#endif
ORACLE_MF_TARGET static inline size_t oracle_mf_boxcar_blocked(const float* x, size_t n_in, int ntaps, int decim, float* y,
                                                                float* scratch)
{
  const long D = decim, q = ntaps / decim, rem = ntaps % decim;
  const size_t n_out = n_in / (size_t)decim;
  if (rem != 0 || q < 1 || D < 2) return oracle_mf_boxcar(x, n_in, ntaps, decim, y);  /* partial blocks: plain form */
  /* B(m), m = -(q-1) .. n_out-1, stored at scratch[2 * (m + q - 1)];  blocks that start before sample 0 */
  float* Bs = scratch;
  for (long m = -(q - 1); m <= 0 && m < (long)n_out; m++) {
    float ar, ai;
    oracle_mf_partial(x, D * m - D + 1, D * m, &ar, &ai);
    if (D * m < 0) { ar = 0.0f; ai = 0.0f; }   /* entirely before the capture: the sum of zero terms */
    Bs[2 * (m + q - 1)] = ar; Bs[2 * (m + q - 1) + 1] = ai;
  }
  for (long m = 1; m < (long)n_out; m++) {
    const float* p = x + 2 * (D * m - D + 1);
    float ar = p[0], ai = p[1];
    for (long k = 1; k < D; k++) { ar = ar + p[2 * k]; ai = ai + p[2 * k + 1]; }
    Bs[2 * (m + q - 1)] = ar; Bs[2 * (m + q - 1) + 1] = ai;
  }
  for (long n = 0; n < (long)n_out; n++) {
    const float* b = Bs + 2 * n;   /* B(n-q+1) */
    float ar = b[0], ai = b[1];
    for (long k = 1; k < q; k++) { ar = ar + b[2 * k]; ai = ai + b[2 * k + 1]; }
    y[2 * n] = ar; y[2 * n + 1] = ai;
  }
  return n_out;
}

/* plain sequential-ascending and float64 variants, used only to show that the decode result
 * does not depend on the summation order */
static inline size_t oracle_mf_boxcar_sequential(const float* x, size_t n_in, int ntaps, int decim, float* y, int use_double)
{
  size_t n_out = n_in / (size_t)decim;
  for (size_t n = 0; n < n_out; n++) {
    float ar = 0.0f, ai = 0.0f;
    double dr = 0.0, di = 0.0;
    long newest = (long)(n * (size_t)decim);
    for (long k = newest - (ntaps - 1); k <= newest; k++) {
      if (k < 0) continue;
      ar = ar + x[2 * k]; ai = ai + x[2 * k + 1];
      dr += x[2 * k]; di += x[2 * k + 1];
    }
    y[2 * n] = use_double ? (float)dr : ar;
    y[2 * n + 1] = use_double ? (float)di : ai;
  }
  return n_out;
}

#ifdef __cplusplus
}
#endif
#endif
