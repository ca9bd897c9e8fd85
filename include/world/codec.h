/* world-b200 drop-in for externs/WORLD_v2/src/world/codec.h:16-92.  Same signatures.  The
 * spectral-envelope codec (the two functions the analysis / synth tools import,
 * W/test/analysis.cpp:304, W/test/synth.cpp:186) runs on the GPU (libworld_b200.so); the
 * band-aperiodicity codec, which neither tool imports, is a host-side link-compatibility helper. */
#ifndef WORLD_CODEC_H_
#define WORLD_CODEC_H_
#include "world/macrodefinitions.h"
WORLD_BEGIN_C_DECLS
/* replaces W/src/codec.cpp:211-214 */
WORLD_API int GetNumberOfAperiodicities(int fs);
/* replaces W/src/codec.cpp:216-235: 20 log10(ap) sampled at 3 kHz, 6 kHz, ... (interp1Q) */
WORLD_API void CodeAperiodicity(const double * const *aperiodicity, int f0_length, int fs,
                                int fft_size, int number_of_aperiodicities,
                                double **coded_aperiodicity);
/* replaces W/src/codec.cpp:237-264.  The reference DECLARES (fs, fft_size,
 * number_of_aperiodicities) but DEFINES (fs, number_of_aperiodicities, fft_size); what a linked
 * caller gets is the definition's order, kept here and spelled out. */
WORLD_API void DecodeAperiodicity(const double * const *coded_aperiodicity, int f0_length, int fs,
                                  int number_of_aperiodicities, int fft_size,
                                  double **aperiodicity);
/* replaces W/src/codec.cpp:266-295 */
WORLD_API void CodeSpectralEnvelope(const double * const *spectrogram, int f0_length, int fs,
                                    int fft_size, int number_of_dimensions,
                                    double **coded_spectral_envelope);
/* replaces W/src/codec.cpp:297-324 */
WORLD_API void DecodeSpectralEnvelope(const double * const *coded_spectral_envelope, int f0_length,
                                      int fs, int fft_size, int number_of_dimensions,
                                      double **spectrogram);
WORLD_END_C_DECLS
#endif
