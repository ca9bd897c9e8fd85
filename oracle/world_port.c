/* TEST INFRASTRUCTURE ONLY -- C restatement of the two strictly sequential pieces of the WORLD
 * path that numpy cannot vectorise (the rest of the restatement is oracle/world_np.py).
 * Only tests/ may load this; the product never does.
 *
 *   port_randn_stream   the reference's randn() after randn_reseed(): xorshift128, twelve draws
 *                       per variate (externs/WORLD_v2/src/matlabfunctions.cpp:247-277)
 *   port_phase_scan     the running phase sum and its wrap of the pulse time base
 *                       (externs/WORLD_v2/src/synthesis.cpp:248-255): a sequential FP64 sum whose
 *                       rounding decides pulse positions
 */
#include <math.h>
#include <stdint.h>

void port_randn_stream(double *out, long long n) {
  uint32_t x = 123456789u, y = 362436069u, z = 521288629u, w = 88675123u;   /* randn_reseed */
  for (long long k = 0; k < n; ++k) {
    uint32_t acc = 0;
    for (int j = 0; j < 12; ++j) {
      const uint32_t t = x ^ (x << 11);
      x = y; y = z; z = w;
      w = (w ^ (w >> 19)) ^ (t ^ (t >> 8));
      acc += w >> 4;
    }
    out[k] = (double)acc / 268435456.0 - 6.0;
  }
}

void port_phase_scan(const double *f0_per_sample, long long n, int fs, double *total, double *wrapped) {
  const double two_pi = 2.0 * 3.1415926535897932384;
  double run = 0.0;
  for (long long i = 0; i < n; ++i) {
    run += two_pi * f0_per_sample[i] / fs;
    total[i] = run;
    wrapped[i] = fmod(run, two_pi);
  }
}
