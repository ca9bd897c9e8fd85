/* world-b200 drop-in for externs/WORLD_v2/src/world/cheaptrick.h:16-80 (this fork's variant:
 * CheapTrickOption carries fft_size and InitializeCheapTrickOption takes fs). */
#ifndef WORLD_CHEAPTRICK_H_
#define WORLD_CHEAPTRICK_H_
#include "world/macrodefinitions.h"
WORLD_BEGIN_C_DECLS
typedef struct {
  double q1;
  double f0_floor;
  int fft_size;
} CheapTrickOption;
/* replaces W/src/cheaptrick.cpp:200-228; spectrogram = f0_length caller-owned rows of
 * option->fft_size/2+1 doubles */
WORLD_API void CheapTrick(const double *x, int x_length, int fs, const double *temporal_positions,
                          const double *f0, int f0_length, const CheapTrickOption *option,
                          double **spectrogram);
/* replaces W/src/cheaptrick.cpp:230-239 */
WORLD_API void InitializeCheapTrickOption(int fs, CheapTrickOption *option);
/* replaces W/src/cheaptrick.cpp:191-194 */
WORLD_API int GetFFTSizeForCheapTrick(int fs, const CheapTrickOption *option);
/* replaces W/src/cheaptrick.cpp:196-198 */
WORLD_API double GetF0FloorForCheapTrick(int fs, int fft_size);
WORLD_END_C_DECLS
#endif
