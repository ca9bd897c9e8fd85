#include "wb_batch.h"
namespace wb {
bool d4c_run(const UttView& u, int fs, int total_frames, const int* frame_utt, const double* frame_t, const double* f0, int fft_size, double threshold, double* ap) { set_error("d4c: not implemented yet"); return false; }
}
