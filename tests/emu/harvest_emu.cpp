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

// The contour logic (FixStep1 - FixStep4, W/src/harvest.cpp:711-1113) on given refined candidates / scores of one
// utterance: harvest_fix_a_kernel, then the warp-cooperative harvest_fix_b_kernel AND the one-lane
// harvest_fix_b_serial_kernel on copies of the same state.  out_warp / out_serial: the merged contour (step 4) of each;
// step3_*: the contour after FixStep3.  The test requires them to be identical bit for bit and runs this under
// ThreadSanitizer (the lanes of the warp version exchange data through global arrays between warp barriers).
extern "C" int emu_harvest_contour(const double* cand, const double* score, int n_fr, int nc, double* out_warp,
                                   double* out_serial, double* step3_warp, double* step3_serial) {
  using namespace wb;
  const int g_off = 0;
  const long long cand_off = 0;
  const size_t nbl = 2 * ((size_t)n_fr + 2 * kSmoothLag) + 8;
  std::vector<double> tmp1(n_fr, 0.0), tmp2(n_fr, 0.0);
  std::vector<int> bl(nbl, 0);
  int nsec = 0;
  wbemu::launch_grid(1, 1, 256, 0, [&]() {
    harvest_fix_a_kernel(cand, score, &g_off, &n_fr, &nc, &cand_off, tmp1.data(), tmp2.data(), bl.data(), &nsec);
  });
  if (nsec < 1) return 3;
  const long long mc_off = 0;
  for (int ver = 0; ver < 2; ++ver) {
    std::vector<double> t1 = tmp1, t2 = tmp2, mc((size_t)nsec * n_fr + 1, 0.0);
    std::vector<int> b2 = bl, chan(n_fr, 0), order(n_fr, 0);
    if (ver == 0)
      wbemu::launch_grid(1, 1, 32, 0, [&]() {
        harvest_fix_b_kernel(cand, score, &g_off, &n_fr, &nc, &cand_off, t1.data(), t2.data(), b2.data(), &mc_off, mc.data(),
                             chan.data(), order.data());
      });
    else
      wbemu::launch_grid(1, 1, 32, 0, [&]() {
        harvest_fix_b_serial_kernel(cand, score, &g_off, &n_fr, &nc, &cand_off, t1.data(), t2.data(), b2.data(), &mc_off, mc.data(),
                                    chan.data(), order.data());
      });
    memcpy(ver == 0 ? out_warp : out_serial, t1.data(), n_fr * sizeof(double));
    memcpy(ver == 0 ? step3_warp : step3_serial, t2.data(), n_fr * sizeof(double));
  }
  return 0;
}

#ifdef EMU_HARVEST_MAIN      // stand-alone run for ThreadSanitizer: emu_main-style, synthetic candidates from a file
#include <stdio.h>
int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  int hdr[2];
  if (fread(hdr, sizeof(int), 2, f) != 2) return 2;
  const int n_fr = hdr[0], nc = hdr[1], slots = nc * 7;
  std::vector<double> cand((size_t)n_fr * slots), score((size_t)n_fr * slots);
  if (fread(cand.data(), sizeof(double), cand.size(), f) != cand.size()) return 2;
  if (fread(score.data(), sizeof(double), score.size(), f) != score.size()) return 2;
  fclose(f);
  std::vector<double> a(n_fr), b(n_fr), c3(n_fr), d3(n_fr);
  const int rc = emu_harvest_contour(cand.data(), score.data(), n_fr, nc, a.data(), b.data(), c3.data(), d3.data());
  if (rc) return rc;
  return memcmp(a.data(), b.data(), n_fr * sizeof(double)) == 0 && memcmp(c3.data(), d3.data(), n_fr * sizeof(double)) == 0 ? 0 : 9;
}
#endif
