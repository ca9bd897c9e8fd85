#include "wb_batch.h"
namespace wb {
bool harvest_run(Batch* b, const HarvestParams& p, double* d_f0_out) { set_error("harvest: not implemented yet"); return false; }
}
