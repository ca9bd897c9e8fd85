/* world-b200 drop-in for externs/WORLD_v2/src/world/stonemask.h:27-29. */
#ifndef WORLD_STONEMASK_H_
#define WORLD_STONEMASK_H_
#include "world/macrodefinitions.h"
WORLD_BEGIN_C_DECLS
/* replaces W/src/stonemask.cpp:211-217 */
WORLD_API void StoneMask(const double *x, int x_length, int fs, const double *temporal_positions,
                         const double *f0, int f0_length, double *refined_f0);
WORLD_END_C_DECLS
#endif
