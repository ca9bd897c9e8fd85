/* world-b200 drop-in for externs/WORLD_v2/src/world/d4c.h:16-46. */
#ifndef WORLD_D4C_H_
#define WORLD_D4C_H_
#include "world/macrodefinitions.h"
WORLD_BEGIN_C_DECLS
typedef struct {
  double threshold;
} D4COption;
/* replaces W/src/d4c.cpp:337-397; aperiodicity = f0_length caller-owned rows of fft_size/2+1 */
WORLD_API void D4C(const double *x, int x_length, int fs, const double *temporal_positions,
                   const double *f0, int f0_length, int fft_size, const D4COption *option,
                   double **aperiodicity);
/* replaces W/src/d4c.cpp:399-401 */
WORLD_API void InitializeD4COption(D4COption *option);
WORLD_END_C_DECLS
#endif
