/* world-b200 drop-in for externs/WORLD_v2/src/world/dio.h:16-43.
 * Same struct layout, same signatures; the work runs on the GPU (libworld_b200.so). */
#ifndef WORLD_DIO_H_
#define WORLD_DIO_H_
#include "world/macrodefinitions.h"
WORLD_BEGIN_C_DECLS
typedef struct {
  double f0_floor;
  double f0_ceil;
  double channels_in_octave;
  double frame_period; /* msec */
  int speed;           /* 1..12 */
  double allowed_range;
} DioOption;
/* replaces W/src/dio.cpp:642-647 */
WORLD_API void Dio(const double *x, int x_length, int fs, const DioOption *option,
                   double *temporal_positions, double *f0);
/* replaces W/src/dio.cpp:649-665 */
WORLD_API void InitializeDioOption(DioOption *option);
/* replaces W/src/dio.cpp:638-640 */
WORLD_API int GetSamplesForDIO(int fs, int x_length, double frame_period);
WORLD_END_C_DECLS
#endif
