// TEST INFRASTRUCTURE: the D4C kernels of hts-train-world_b200/csrc/wb_d4c.cu compiled for the CPU
// (tests/emu/cuda_emu.h) behind one C entry point.  The host orchestration mirrors d4c_run:
// LoveTrain draw counts -> exclusive scan -> LoveTrain -> main draw counts -> scan -> main kernel
// (mode 0) or the split pair d4c_gd_kernel + d4c_tail_kernel (mode 1); bit 1 of mode selects the
// FP32 LoveTrain kernel.  One utterance; only the frames listed in `rows` are launched.
#define WB_HOST_EMU 1
#include "cuda_emu.h"
#include "../../hts-train-world_b200/csrc/wb_d4c.cu"

namespace {
struct X128 { uint32_t x = 123456789u, y = 362436069u, z = 521288629u, w = 88675123u; };
uint32_t next(X128& s) {
  const uint32_t t = s.x ^ (s.x << 11);
  s.x = s.y; s.y = s.z; s.z = s.w;
  s.w = (s.w ^ (s.w >> 19)) ^ (t ^ (t >> 8));
  return s.w;
}
}  // namespace

extern "C" int emu_d4c(const double* x, int x_len, int fs, const double* t, const double* f0, int F, int fft_size,
                       double threshold, int mode, const int* rows, int n_rows, double* ap_rows, double* ap0_out) {
  using namespace wb;
  D4CConst c;
  c.fs = fs;
  c.threshold = threshold;
  c.out_half = fft_size / 2;
  const int nd = static_cast<int>(pow(2.0, 1.0 + static_cast<int>(log(4.0 * fs / kFloorF0D4C + 1) / kLog2)));
  const int nlt = static_cast<int>(pow(2.0, 1.0 + static_cast<int>(log(3.0 * fs / 40.0 + 1) / kLog2)));
  c.log2nd = 0; while ((1 << c.log2nd) < nd) ++c.log2nd;
  c.log2lt = 0; while ((1 << c.log2lt) < nlt) ++c.log2lt;
  if (c.log2nd != 12 || c.log2lt != 12) return 2;                      // the emulation covers the 48 kHz sizes
  c.nbands = static_cast<int>(fmin(kUpperLimit, fs / 2.0 - kFrequencyInterval) / kFrequencyInterval);
  c.window_length = static_cast<int>(kFrequencyInterval * nd / fs) * 2 + 1;
  c.sel_boundary = matlab_round(nd * 8.0 / c.window_length);
  c.band_top = 0;
  for (int i = 0; i < c.nbands; ++i) {
    c.centers[i] = static_cast<int>(kFrequencyInterval * (i + 1) * nd / fs);
    c.band_top = std::max(c.band_top, c.centers[i] - c.window_length / 2 + c.window_length - 1);
  }
  c.lt_b0 = static_cast<int>(ceil(100.0 * nlt / fs));
  c.lt_b1 = static_cast<int>(ceil(4000.0 * nlt / fs));
  c.lt_b2 = static_cast<int>(ceil(7900.0 * nlt / fs));
  std::vector<double> win(c.window_length);
  for (int i = 0; i < c.window_length; ++i) {
    const double tmp = i / (c.window_length - 1.0);
    win[i] = 0.355768 - 0.487396 * cos(2.0 * kPi * tmp) + 0.144232 * cos(4.0 * kPi * tmp) - 0.012604 * cos(6.0 * kPi * tmp);
  }
  // tables of size 2^12: exp(-2 pi i k / 4096), k = 0 .. 2048
  std::vector<double2> tw(2049);
  std::vector<float2> twf(2049);
  for (int k = 0; k <= 2048; ++k) {
    const long double a = -2.0L * 3.14159265358979323846264338327950288L * k / 4096;
    tw[k] = make_double2((double)cosl(a), (double)sinl(a));
    twf[k] = make_float2((float)tw[k].x, (float)tw[k].y);
  }
  // utterance table (one utterance, 16-byte aligned samples with one pad sample)
  std::vector<double> xs(x, x + x_len);
  xs.push_back(0.0);
  xs.push_back(0.0);
  const long long x_off = 0;
  const int f_off = 0;
  UttView u{xs.data(), &x_off, &x_len, &f_off, &F, 1};
  std::vector<int> frame_utt(F, 0);
  // dynamic shared memory exactly as d4c_run requests it
  const int hd = nd / 2, nl = 1 << c.log2lt;
  const size_t smem_lt = (size_t)(2 * cpad_size(nl / 2) + 96) * sizeof(double);
  const size_t smem_main = d4c_cbuf_slots(nd, c.nbands) * sizeof(double2) + (size_t)(2 * (hd + 8) + 160) * sizeof(double) +
                           sizeof(SelectScratch) + (kMaxBands + 2) * sizeof(double);
  const size_t smem_lt32 = (size_t)((cpadf(nl / 2) + 4 + 1) & ~1) * sizeof(float2) + (96 + kLtStage) * sizeof(double);
#ifdef WB_D4C_HAS_SPLIT
  const size_t smem_gd = (size_t)(2 * (hd + 8) + 160) * sizeof(double) + (size_t)cpad_size(hd) * sizeof(double2);
  const size_t smem_tail = (size_t)d4c_tail_fb_slots(nd) * sizeof(float2) + (size_t)((c.nbands * (hd + 1) + 1) & ~1) * sizeof(float) +
                           (kMaxBands + 2) * sizeof(double);
#endif
  wbemu::smem_overruns = 0;
  // LoveTrain
  std::vector<long long> offs_lt(F), offs_main(F);
  long long tot_lt = 0, tot_main = 0;
  for (int f = 0; f < F; ++f) {
    offs_lt[f] = tot_lt;
    tot_lt += f0[f] == 0.0 ? 0 : 2LL * matlab_round(div_rn(mul_rn(1.5, (double)fs), fmax(f0[f], 40.0))) + 1;
  }
  std::vector<double> ap0(F, 0.0);
  std::vector<int> all(F);
  for (int f = 0; f < F; ++f) all[f] = f;
  // draws: LoveTrain of every voiced frame first, then 3 windows per processed frame -- the main
  // offsets need ap0 of every frame when threshold > 0; with threshold <= 0 only voicing matters
  std::vector<int> lt_rows = threshold > 0.0 ? all : std::vector<int>(rows, rows + n_rows);
  long long need = tot_lt;
  {
    long long m = 0;
    for (int f = 0; f < F; ++f) m += f0[f] == 0.0 ? 0 : 3LL * (2LL * d4c_hwl(4.0, fs, fmax(kFloorF0D4C, f0[f])) + 1);
    need += m;
  }
  std::vector<uint32_t> randn_tab((size_t)need + 16);
  {
    X128 s;
    for (auto& v : randn_tab) { uint32_t acc = 0; for (int j = 0; j < 12; ++j) acc += next(s) >> 4; v = acc; }
  }
#ifndef WB_D4C_HAS_SPLIT
  if (mode & 1) return 3;                                              // this source tree has no split main kernels
#endif
  if (mode & 2) {                                                      // LoveTrain with the FP32 transform (the GPU default)
    wbemu::launch(lt_rows, F, 256, smem_lt32, [&]() { d4c_lovetrain32_kernel<12>(u, frame_utt.data(), t, f0, offs_lt.data(), randn_tab.data(), twf.data(), c, ap0.data()); });
  } else {
    wbemu::launch(lt_rows, F, 256, smem_lt, [&]() { d4c_lovetrain_kernel<12>(u, frame_utt.data(), t, f0, offs_lt.data(), randn_tab.data(), tw.data(), c, ap0.data()); });
  }
  if (!(threshold > 0.0))
    for (int f = 0; f < F; ++f) if (f0[f] != 0.0 && ap0[f] == 0.0) ap0[f] = 1.0;      // not launched: passes `ap0 <= threshold`
  for (int f = 0; f < F; ++f) {
    offs_main[f] = tot_main;
    tot_main += (f0[f] == 0.0 || ap0[f] <= threshold) ? 0 : 3LL * (2LL * d4c_hwl(4.0, fs, fmax(kFloorF0D4C, f0[f])) + 1);
  }
  std::vector<double> ap((size_t)F * (c.out_half + 1), -1.0);
  const std::vector<int> sel(rows, rows + n_rows);
  if (mode & 1) {
#ifdef WB_D4C_HAS_SPLIT
    std::vector<float> slices((size_t)F * c.nbands * c.window_length, 0.f);
    wbemu::launch(sel, F, 256, smem_gd, [&]() { d4c_gd_kernel<12, 256>(u, frame_utt.data(), t, f0, ap0.data(), offs_main.data(), &tot_lt, randn_tab.data(), tw.data(), win.data(), c, slices.data()); });
    wbemu::launch(sel, F, 256, smem_tail, [&]() { d4c_tail_kernel<12, 256>(f0, ap0.data(), slices.data(), twf.data(), c, ap.data()); });
#endif
  } else {
    wbemu::launch(sel, F, 256, smem_main, [&]() { d4c_main_kernel<12, 256, 4>(u, frame_utt.data(), t, f0, ap0.data(), offs_main.data(), &tot_lt, randn_tab.data(), tw.data(), twf.data(), win.data(), c, ap.data()); });
  }
  if (ap0_out) memcpy(ap0_out, ap0.data(), F * sizeof(double));
  for (int r = 0; r < n_rows; ++r)
    memcpy(ap_rows + (size_t)r * (c.out_half + 1), ap.data() + (size_t)rows[r] * (c.out_half + 1), (c.out_half + 1) * sizeof(double));
  return wbemu::smem_overruns ? 4 : 0;                                 // 4: a kernel wrote past its shared-memory allocation
}
