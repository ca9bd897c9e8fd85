// TEST INFRASTRUCTURE: the CheapTrick kernel of hts-train-world_b200/csrc/wb_cheaptrick.cu compiled
// for the CPU (tests/emu/cuda_emu.h).  Host orchestration as in cheaptrick_run: draw counts ->
// exclusive scan -> cheaptrick_kernel<11, 128> (48 kHz: fft_size 2048).  One utterance; only the
// frames listed in `rows` are launched.
#define WB_HOST_EMU 1
#include "cuda_emu.h"
#include "../../hts-train-world_b200/csrc/wb_cheaptrick.cu"

extern "C" int emu_cheaptrick(const double* x, int x_len, int fs, const double* t, const double* f0, int F, int fft_size,
                              double q1, const int* rows, int n_rows, double* sp_rows) {
  using namespace wb;
  if (fft_size != 2048) return 2;
  const double f0_floor = 3.0 * fs / (fft_size - 3.0);
  std::vector<long long> offs(F);
  long long tot = 0;
  for (int f = 0; f < F; ++f) {
    offs[f] = tot;
    tot += 2LL * cheaptrick_hwl(fs, cheaptrick_f0(f0[f], f0_floor)) + 1 + fft_size / 2 + 1;
  }
  std::vector<uint32_t> randn_tab((size_t)tot + 16);
  {
    uint32_t sx = 123456789u, sy = 362436069u, sz = 521288629u, sw = 88675123u;
    for (auto& v : randn_tab) {
      uint32_t acc = 0;
      for (int j = 0; j < 12; ++j) {
        const uint32_t tt = sx ^ (sx << 11);
        sx = sy; sy = sz; sz = sw;
        sw = (sw ^ (sw >> 19)) ^ (tt ^ (tt >> 8));
        acc += sw >> 4;
      }
      v = acc;
    }
  }
  std::vector<double2> tw(1025);
  std::vector<float2> twf(1025);
  for (int k = 0; k <= 1024; ++k) {
    const long double a = -2.0L * 3.14159265358979323846264338327950288L * k / 2048;
    tw[k] = make_double2((double)cosl(a), (double)sinl(a));
    twf[k] = make_float2((float)tw[k].x, (float)tw[k].y);
  }
  std::vector<double> xs(x, x + x_len);
  xs.push_back(0.0);
  xs.push_back(0.0);
  const long long x_off = 0;
  const int f_off = 0;
  UttView u{xs.data(), &x_off, &x_len, &f_off, &F, 1};
  std::vector<int> frame_utt(F, 0);
  std::vector<double> sp((size_t)F * (fft_size / 2 + 1), -1.0);
  const size_t smem = cpad_size(fft_size / 2) * sizeof(double2) + (fft_size + 16 + 128) * sizeof(double);   // as cheaptrick_run
  wbemu::smem_overruns = 0;
  wbemu::launch(std::vector<int>(rows, rows + n_rows), F, 128, smem, [&]() {
    cheaptrick_kernel<11, 128>(u, frame_utt.data(), t, f0, offs.data(), randn_tab.data(), tw.data(), twf.data(), fs, 11, q1,
                               f0_floor, sp.data());
  });
  for (int r = 0; r < n_rows; ++r)
    memcpy(sp_rows + (size_t)r * (fft_size / 2 + 1), sp.data() + (size_t)rows[r] * (fft_size / 2 + 1), (fft_size / 2 + 1) * sizeof(double));
  return wbemu::smem_overruns ? 4 : 0;
}
