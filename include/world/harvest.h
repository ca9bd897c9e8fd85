/* world-b200 drop-in for externs/WORLD_v2/src/world/harvest.h:16-59. */
#ifndef WORLD_HARVEST_H_
#define WORLD_HARVEST_H_
#include "world/macrodefinitions.h"
WORLD_BEGIN_C_DECLS
typedef struct {
  double f0_floor;
  double f0_ceil;
  double frame_period;
} HarvestOption;
/* replaces W/src/harvest.cpp:1223-1255 */
WORLD_API void Harvest(const double *x, int x_length, int fs, const HarvestOption *option,
                       double *temporal_positions, double *f0);
/* replaces W/src/harvest.cpp:1257-1262 */
WORLD_API void InitializeHarvestOption(HarvestOption *option);
/* replaces W/src/harvest.cpp:1217-1221 */
WORLD_API int GetSamplesForHarvest(int fs, int x_length, double frame_period);
WORLD_END_C_DECLS
#endif
