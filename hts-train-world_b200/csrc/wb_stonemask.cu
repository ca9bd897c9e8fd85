#include "wb_batch.h"
namespace wb {
bool stonemask_run(const UttView& u, int fs, int total_frames, const int* frame_utt, const double* frame_t, const double* f0_in, double* f0_out) { set_error("stonemask: not implemented yet"); return false; }
}
