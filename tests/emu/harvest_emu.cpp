// TEST INFRASTRUCTURE: the refinement kernel of hts-train-world_b200/csrc/wb_harvest.cu (harvest_refine_thread_kernel:
// one thread per (1 ms frame, base candidate), Goertzel recurrences at the <= 6 harmonic bins) compiled for the CPU
// (tests/emu/cuda_emu.h).  One utterance; the caller supplies the decimated signal and the base candidates and gets
// the refined candidates / scores of every (frame, overlapped slot) back -- tests/test_kernel_emulation.py compares
// them with the restatement of GetRefinedF0 in oracle/harvest_np.py.
#define WB_HOST_EMU 1
#include "cuda_emu.h"
#include "../../hts-train-world_b200/csrc/wb_harvest.cu"

extern "C" int emu_harvest_refine(const double* y, int y_len, double actual_fs, double f0_floor, double f0_ceil,
                                  const double* base, int n_fr, int nc, int max_base, double* cand_out, double* score_out) {
  using namespace wb;
  if (nc < 1 || nc > max_base || n_fr < 1) return 2;
  HarvestConst c = {};
  c.fs = static_cast<int>(actual_fs); c.r = 1; c.nch = 0; c.lag = 0;
  c.actual_fs = actual_fs; c.f0_floor = f0_floor; c.f0_ceil = f0_ceil;
  // the concatenated compact FP64 twiddle tables (Context::d_twiddle_c)
  std::vector<double2> twc(Context::tw_c_offset(kTwLog2 + 1));
  for (int L = 4; L <= kTwLog2; ++L)
    for (int k = 0; k <= (1 << (L - 1)); ++k) {
      const long double a = -2.0L * 3.14159265358979323846264338327950288L * k / (1 << L);
      twc[Context::tw_c_offset(L) + k] = make_double2((double)cosl(a), (double)sinl(a));
    }
  double mean = 0.0;
  for (int i = 0; i < y_len; ++i) mean += y[i];
  mean /= y_len;
  const long long y_off = 0, cand_off = 0;
  const int g_off = 0;
  const long long total_work = (long long)n_fr * nc;
  const int blocks = (int)((total_work + 127) / 128);
  std::vector<int> all(blocks);
  for (int i = 0; i < blocks; ++i) all[i] = i;
  wbemu::launch(all, blocks, 128, 0, [&]() {
    harvest_refine_thread_kernel(y, &y_off, &y_len, &mean, base, &g_off, &n_fr, &nc, &cand_off, max_base, c, 1, total_work,
                                 twc.data(), cand_out, score_out);
  });
  return 0;
}
