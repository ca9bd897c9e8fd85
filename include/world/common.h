/* world-b200 drop-in for externs/WORLD_v2/src/world/common.h:18-143 (helpers shared by the
 * reference's stages and exported by libworld.a).  Same struct layouts and signatures so
 * third-party callers link.  Link-compatibility helpers, HOST code: the batched
 * path has its own device versions (wb_spectral.cuh) and never calls these. */
#ifndef WORLD_COMMON_H_
#define WORLD_COMMON_H_
#include "world/fft.h"
#include "world/macrodefinitions.h"
WORLD_BEGIN_C_DECLS

/* ---- spectral helpers ------------------------------------------------------------------------ */
/* replaces W/src/common.cpp:51-54: the power of two above `sample` */
WORLD_API int GetSuitableFFTSize(int sample);
/* replaces W/src/common.cpp:56-75: bins below f0 receive the spectrum mirrored about f0 / 2;
 * writes output[0 .. 1 + int(f0 fft_size / fs)) only */
WORLD_API void DCCorrection(const double *input, double current_f0, int fs, int fft_size,
                            double *output);
/* replaces W/src/common.cpp:77-111: rectangular smoothing of `width` Hz over fft_size/2 + 1 bins */
WORLD_API void LinearSmoothing(const double *input, double width, int fs, int fft_size,
                               double *output);
/* replaces W/src/common.cpp:113-121: 4-term cosine window */
WORLD_API void NuttallWindow(int y_length, double *y);

static inline int MyMaxInt(int x, int y) { return x > y ? x : y; }
static inline int MyMinInt(int x, int y) { return x < y ? x : y; }
static inline double MyMaxDouble(double x, double y) { return x > y ? x : y; }
static inline double MyMinDouble(double x, double y) { return x < y ? x : y; }
/* aperiodicity clamped to [0.001, 1 - 1e-12] (W/src/world/common.h:111-113) */
static inline double GetSafeAperiodicity(double x) {
  return MyMaxDouble(0.001, MyMinDouble(0.999999999999, x));
}

/* ---- transform work areas: Initialize* allocates, Destroy* frees (W/src/common.cpp:125-226) ----- */
typedef struct {
  int fft_size;
  double *waveform;       /* fft_size reals in */
  fft_complex *spectrum;  /* bins 0 .. fft_size/2 out */
  fft_plan forward_fft;
} ForwardRealFFT;
WORLD_API void InitializeForwardRealFFT(int fft_size, ForwardRealFFT *forward_real_fft);
WORLD_API void DestroyForwardRealFFT(ForwardRealFFT *forward_real_fft);

typedef struct {
  int fft_size;
  double *waveform;       /* fft_size reals out (unnormalised) */
  fft_complex *spectrum;  /* bins 0 .. fft_size/2 in */
  fft_plan inverse_fft;
} InverseRealFFT;
WORLD_API void InitializeInverseRealFFT(int fft_size, InverseRealFFT *inverse_real_fft);
WORLD_API void DestroyInverseRealFFT(InverseRealFFT *inverse_real_fft);

typedef struct {
  int fft_size;
  fft_complex *input;
  fft_complex *output;
  fft_plan inverse_fft;
} InverseComplexFFT;
WORLD_API void InitializeInverseComplexFFT(int fft_size, InverseComplexFFT *inverse_complex_fft);
WORLD_API void DestroyInverseComplexFFT(InverseComplexFFT *inverse_complex_fft);

typedef struct {
  int fft_size;
  double *log_spectrum;                /* bins 0 .. fft_size/2 in (log amplitude) */
  fft_complex *minimum_phase_spectrum; /* bins 0 .. fft_size/2 out */
  fft_complex *cepstrum;
  fft_plan inverse_fft;
  fft_plan forward_fft;
} MinimumPhaseAnalysis;
WORLD_API void InitializeMinimumPhaseAnalysis(int fft_size, MinimumPhaseAnalysis *minimum_phase);
/* replaces W/src/common.cpp:182-217 */
WORLD_API void GetMinimumPhaseSpectrum(const MinimumPhaseAnalysis *minimum_phase);
WORLD_API void DestroyMinimumPhaseAnalysis(MinimumPhaseAnalysis *minimum_phase);

WORLD_END_C_DECLS
#endif
