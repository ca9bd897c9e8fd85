/* world-b200 drop-in for externs/WORLD_v2/src/world/fft.h:16-57 (the FFTW-style plan API the
 * reference wraps around its vendored FFT).  Same struct layout and signatures so third-party
 * callers of libworld.a link.  Link-compatibility helper, HOST code: it is not
 * on the batched path -- Dio / StoneMask / CheapTrick / D4C / Synthesis / Harvest / the codec
 * run their transforms in the CUDA kernels and never call this API. */
#ifndef WORLD_FFT_H_
#define WORLD_FFT_H_
#include "world/macrodefinitions.h"
WORLD_BEGIN_C_DECLS
#define FFT_FORWARD 1
#define FFT_BACKWARD 2
#define FFT_ESTIMATE 3
typedef double fft_complex[2];   /* (re, im) */
/* Field order and types are the reference's (the struct travels by value); what the work arrays
 * hold is private to the implementation. */
typedef struct {
  int n;               /* transform length (a power of two) */
  int sign;            /* FFT_FORWARD / FFT_BACKWARD */
  unsigned int flags;  /* FFT_ESTIMATE, ignored */
  fft_complex *c_in;   /* complex input, or NULL (r2c) */
  double *in;          /* real input, or NULL */
  fft_complex *c_out;  /* complex output, or NULL (c2r) */
  double *out;         /* real output, or NULL */
  double *input;       /* work buffer, 2 n doubles */
  int *ip;             /* bit-reversal table, n ints */
  double *w;           /* twiddle table, 5 n / 4 doubles */
} fft_plan;
/* replaces W/src/fft.cpp:76-97: c2c; FFT_FORWARD computes DFT(conj(in)), FFT_BACKWARD
 * n * IDFT(conj(in)) (the reference's conventions, W/src/fft.cpp:36-45,61-71) */
WORLD_API fft_plan fft_plan_dft_1d(int n, fft_complex *in, fft_complex *out, int sign,
                                   unsigned int flags);
/* replaces W/src/fft.cpp:99-121: unnormalised inverse of bins 0..n/2 (imaginary parts of bins 0
 * and n/2 ignored) */
WORLD_API fft_plan fft_plan_dft_c2r_1d(int n, fft_complex *in, double *out, unsigned int flags);
/* replaces W/src/fft.cpp:123-145: forward DFT, bins 0..n/2 written */
WORLD_API fft_plan fft_plan_dft_r2c_1d(int n, double *in, fft_complex *out, unsigned int flags);
/* replaces W/src/fft.cpp:147-153 */
WORLD_API void fft_execute(fft_plan p);
/* replaces W/src/fft.cpp:155-166 */
WORLD_API void fft_destroy_plan(fft_plan p);
WORLD_END_C_DECLS
#endif
