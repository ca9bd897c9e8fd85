#include "wb_batch.h"
namespace wb {
bool dio_run(Batch* b, const DioParams& p, double* d_f0_out) { set_error("dio: not implemented yet"); return false; }
}
