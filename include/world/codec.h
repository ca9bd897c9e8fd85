/* world-b200 drop-in for externs/WORLD_v2/src/world/codec.h:71-92 (the two functions the
 * analysis / synth tools import, W/test/analysis.cpp:304, W/test/synth.cpp:186).  Same
 * signatures; the mel-DCT runs on the GPU (libworld_b200.so). */
#ifndef WORLD_CODEC_H_
#define WORLD_CODEC_H_
#include "world/macrodefinitions.h"
WORLD_BEGIN_C_DECLS
/* replaces W/src/codec.cpp:211-214 */
WORLD_API int GetNumberOfAperiodicities(int fs);
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
