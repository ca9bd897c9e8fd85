// TEST INFRASTRUCTURE: the StoneMask kernel of hts-train-world_b200/csrc/wb_stonemask.cu compiled for
// the CPU (tests/emu/cuda_emu.h).  Host side as in stonemask_run: the largest transform of the batch
// sizes the shared memory (computed here on the host), then stonemask_kernel for the listed frames.
#define WB_HOST_EMU 1
#include "cuda_emu.h"
#include "../../hts-train-world_b200/csrc/wb_stonemask.cu"

extern "C" int emu_stonemask(const double* x, int x_len, int fs, const double* t, const double* f0, int F, const int* rows,
                             int n_rows, double* f0_rows) {
  using namespace wb;
  int h_max = 3;
  for (int f = 0; f < F; ++f)
    if (stonemask_in_range(f0[f], fs)) h_max = std::max(h_max, stonemask_log2fft(stonemask_hwl(f0[f], fs)));
  if (h_max > 13) return 2;
  // the concatenated compact FP32 twiddle tables (Context::d_twiddle_cf)
  std::vector<float2> twcf(Context::tw_c_offset(kTwLog2 + 1));
  for (int L = 4; L <= kTwLog2; ++L)
    for (int k = 0; k <= (1 << (L - 1)); ++k) {
      const long double a = -2.0L * 3.14159265358979323846264338327950288L * k / (1 << L);
      twcf[Context::tw_c_offset(L) + k] = make_float2((float)(double)cosl(a), (float)(double)sinl(a));
    }
  std::vector<double> xs(x, x + x_len);
  xs.push_back(0.0);
  xs.push_back(0.0);
  const long long x_off = 0;
  const int f_off = 0;
  UttView u{xs.data(), &x_off, &x_len, &f_off, &F, 1};
  std::vector<int> frame_utt(F, 0);
  std::vector<double> out(F, -1.0);
  const size_t smem = ((cpad_size(1 << h_max) + 1) & ~1) * sizeof(float2) + ((size_t)(1 << h_max) / 2 + 8) * sizeof(double) +
                      ((size_t)(1 << h_max) / 2) * sizeof(int);                 // as stonemask_run
  wbemu::smem_overruns = 0;
  wbemu::launch(std::vector<int>(rows, rows + n_rows), F, 256, smem,
                [&]() { stonemask_kernel(u, frame_utt.data(), t, f0, twcf.data(), fs, h_max, out.data()); });
  for (int r = 0; r < n_rows; ++r) f0_rows[r] = out[rows[r]];
  return wbemu::smem_overruns ? 4 : 0;
}

// the default path: stonemask_dft_kernel (one warp per frame, direct evaluation of the harmonic bins)
extern "C" int emu_stonemask_dft(const double* x, int x_len, int fs, const double* t, const double* f0, int F, const int* rows,
                                 int n_rows, double* f0_rows) {
  using namespace wb;
  std::vector<double2> twc(Context::tw_c_offset(kTwLog2 + 1));
  for (int L = 4; L <= kTwLog2; ++L)
    for (int k = 0; k <= (1 << (L - 1)); ++k) {
      const long double a = -2.0L * 3.14159265358979323846264338327950288L * k / (1 << L);
      twc[Context::tw_c_offset(L) + k] = make_double2((double)cosl(a), (double)sinl(a));
    }
  std::vector<double> xs(x, x + x_len);
  xs.push_back(0.0);
  xs.push_back(0.0);
  const long long x_off = 0;
  const int f_off = 0;
  UttView u{xs.data(), &x_off, &x_len, &f_off, &F, 1};
  std::vector<int> frame_utt(F, 0);
  std::vector<double> out(F, -1.0);
  std::vector<int> blocks;                                  // the CTAs that hold the requested frames
  for (int r = 0; r < n_rows; ++r)
    if (blocks.empty() || blocks.back() != rows[r] / kSmWarps) blocks.push_back(rows[r] / kSmWarps);
  wbemu::smem_overruns = 0;
  wbemu::launch(blocks, (F + kSmWarps - 1) / kSmWarps, kSmWarps * 32, 0,
                [&]() { stonemask_dft_kernel(u, frame_utt.data(), t, f0, twc.data(), fs, F, out.data()); });
  for (int r = 0; r < n_rows; ++r) f0_rows[r] = out[rows[r]];
  return wbemu::smem_overruns ? 4 : 0;
}
