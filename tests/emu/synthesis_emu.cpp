// TEST INFRASTRUCTURE: the Synthesis kernels of hts-train-world_b200/csrc/wb_synthesis.cu compiled for
// the CPU (tests/emu/cuda_emu.h).  Host orchestration as in synthesis_run, one utterance: pulse
// bound -> time base (increments, running phase, pulse count / scan / write) -> classification ->
// synth_item_kernel<11, float2> (48 kHz: fft_size 2048, the FP32 channel the library runs).
#define WB_HOST_EMU 1
#include "cuda_emu.h"
#include "../../hts-train-world_b200/csrc/wb_synthesis.cu"

extern "C" int emu_synthesis(const double* f0, int F, const double* sp, const double* ap, int fft_size, double frame_period_ms,
                             int fs, int y_length, double* y_out) {
  using namespace wb;
  if (fft_size != 2048 || F < 2) return 2;
  const int N = fft_size, n_utt = 1;
  SynthConst c;
  c.fs = fs;
  c.log2n = 11;
  c.frame_period_s = frame_period_ms / 1000.0;
  c.lowest_f0 = fs / N + 1.0;
  c.f0_max_len = F;
  const int f_off = 0, f_len = F;
  const long long y_off = 0;
  const int y_len = y_length;
  std::vector<double> y((size_t)((y_length + 1) & ~1) + 2, 0.0);
  int cap = 0, cnt = 0, poff = 0;
  wbemu::smem_overruns = 0;
  wbemu::launch_grid(n_utt, 1, 128, 0, [&]() { synth_pulse_bound_kernel(f0, &f_off, &f_len, c, &cap); });
  const int total_p = cap;
  std::vector<uint32_t> randn_tab((size_t)y_length + 32);
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
  std::vector<int> p_index(total_p + 1), p_utt(total_p + 1, -1), list_per(total_p + 1), list_aper(total_p + 1);
  std::vector<double> p_shift(total_p + 1);
  std::vector<unsigned char> p_vuv(total_p + 1);
  {
    const int n_chunks_max = (y_length + kTbChunk - 1) / kTbChunk + 1;
    std::vector<double> inc((size_t)y.size() + 1), tot((size_t)y.size() + 1);
    std::vector<unsigned char> vuv((size_t)y.size() + 1);
    std::vector<int> counts((size_t)n_utt * n_chunks_max, 0);
    wbemu::launch_grid((y_length + kTbChunk - 1) / kTbChunk, n_utt, 256, 0,
                       [&]() { synth_inc_kernel(f0, &f_off, &f_len, &y_len, &y_off, c, inc.data(), vuv.data()); });
    wbemu::launch_grid(n_utt, 1, kTbChunk, 0, [&]() { synth_phase_kernel(inc.data(), &y_off, &y_len, tot.data()); });
    {   // the four-samples-per-thread version (the library's default) must give the same bits
      std::vector<double> tot4(tot.size(), 0.0);
      wbemu::launch_grid(n_utt, 1, kTbChunk / kTbPer, 0, [&]() { synth_phase4_kernel(inc.data(), &y_off, &y_len, tot4.data()); });
      if (memcmp(tot4.data(), tot.data(), (size_t)y_length * sizeof(double)) != 0) return 7;
    }
    wbemu::launch_grid(n_chunks_max, n_utt, 256, 0, [&]() {
      synth_pulses_kernel<false>(tot.data(), vuv.data(), &y_off, &y_len, c, n_chunks_max, counts.data(), &poff, &cap, p_index.data(),
                                 p_shift.data(), p_vuv.data(), p_utt.data());
    });
    wbemu::launch_grid(1, 1, 128, 0, [&]() { synth_pulse_scan_kernel(counts.data(), n_utt, n_chunks_max, &cnt); });
    wbemu::launch_grid(n_chunks_max, n_utt, 256, 0, [&]() {
      synth_pulses_kernel<true>(tot.data(), vuv.data(), &y_off, &y_len, c, n_chunks_max, counts.data(), &poff, &cap, p_index.data(),
                                p_shift.data(), p_vuv.data(), p_utt.data());
    });
  }
  if (cnt > cap) return 5;
  std::vector<double> rem(N);
  double dc_component = 0.0;
  for (int i = 0; i < N / 2; ++i) {
    rem[i] = 0.5 - 0.5 * cos(2.0 * kPi * (i + 1.0) / (1.0 + N));
    rem[N - i - 1] = rem[i];
    dc_component += rem[i] * 2.0;
  }
  for (int i = 0; i < N / 2; ++i) { rem[i] /= dc_component; rem[N - i - 1] = rem[i]; }
  int cnt2[2] = {0, 0};
  {
    const int ncb = (total_p + 255) / 256;                         // count pass, scan, write pass: lists in pulse order
    std::vector<int> blk(2 * (size_t)std::max(1, ncb), 0);
    wbemu::launch_grid(ncb, 1, 256, 0, [&]() {
      synth_classify_kernel<false>(ap, &f_off, &f_len, p_index.data(), p_vuv.data(), p_utt.data(), total_p, c, blk.data(), list_per.data(), list_aper.data());
    });
    wbemu::launch_grid(1, 1, 64, 0, [&]() { synth_classify_scan_kernel(blk.data(), ncb, cnt2); });
    wbemu::launch_grid(ncb, 1, 256, 0, [&]() {
      synth_classify_kernel<true>(ap, &f_off, &f_len, p_index.data(), p_vuv.data(), p_utt.data(), total_p, c, blk.data(), list_per.data(), list_aper.data());
    });
  }
  const int n_per = cnt2[0], n_aper = cnt2[1];
  const int n_items = n_per + (n_aper + 1) / 2;
  if (n_items > 0) {
    std::vector<float2> twf(1025);
    for (int k = 0; k <= 1024; ++k) {
      const long double a = -2.0L * 3.14159265358979323846264338327950288L * k / 2048;
      twf[k] = make_float2((float)(double)cosl(a), (float)(double)sinl(a));
    }
    const size_t smem = 2 * cpad_size(N) * sizeof(float2) + 96 * sizeof(double);        // as synthesis_run
    wbemu::launch_grid(n_items, 1, 256, smem, [&]() {
      synth_item_kernel<11, float2>(sp, ap, &f_off, &f_len, &y_off, &y_len, &poff, &cnt, p_index.data(), p_shift.data(), p_vuv.data(),
                                    p_utt.data(), list_per.data(), list_aper.data(), n_per, n_aper, randn_tab.data(), twf.data(),
                                    rem.data(), c, y.data());
    });
    wbemu::launch_grid(1, 1, 256, 0, [&]() { synth_ola_finish_kernel(y.data(), (long long)y_length); });   // integer sums -> samples
  }
  memcpy(y_out, y.data(), (size_t)y_length * sizeof(double));
  return wbemu::smem_overruns ? 4 : 0;
}
