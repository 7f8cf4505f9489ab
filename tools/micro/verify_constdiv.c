/* Exhaustive check (developer aid + tests): for a constant divisor D, does
 *     c = RN(1/D);  q = RN(x*c);  r = fma(-q, D, x);  q' = fma(r, c, q)
 * equal the IEEE quotient RN(x/D) for EVERY binary32 mantissa?  (Scaling x by a power of two scales every
 * intermediate exactly, so one binade covers all normal x whose intermediates stay normal.)
 * Usage: verify_constdiv D [D ...]; exit code 0 iff all pass. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int check(float D)
{
  const float c = 1.0f / D;
  long bad = 0;
  for (uint32_t m = 0; m < (1u << 23); m++) {
    for (int e = 0; e < 2; e++) { /* two binades: also exercises the binade crossing of the quotient */
      uint32_t bits = ((uint32_t)(127 + 20 * e) << 23) | m;
      float x;
      memcpy(&x, &bits, 4);
      for (int sgn = 0; sgn < 2; sgn++) {
        float xs = sgn ? -x : x;
        volatile float want = xs / D;
        float q = xs * c;
        float r = fmaf(-q, D, xs);
        float q2 = fmaf(r, c, q);
        if (q2 != want) bad++;
      }
    }
  }
  printf("D = %g: %ld mismatches over 2^25 inputs\n", (double)D, bad);
  return bad == 0;
}

int main(int argc, char** argv)
{
  int ok = 1;
  for (int i = 1; i < argc; i++) ok &= check((float)atof(argv[i]));
  return ok ? 0 : 1;
}
