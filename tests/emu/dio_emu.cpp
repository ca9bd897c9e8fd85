// TEST INFRASTRUCTURE: the Dio kernels of hts-train-world_b200/csrc/wb_dio.cu / wb_zerocross.cuh
// compiled for the CPU (tests/emu/cuda_emu.h).  Host orchestration as in dio_run (speed 1, one
// utterance): filter bank -> mean -> overlap-save band filters -> zero-crossing count / scan /
// write -> candidates -> best contour + FixF0Contour.
#define WB_HOST_EMU 1
#include "cuda_emu.h"
#define cudaMemcpy(dst, src, n, kind) (memcpy((dst), (src), (n)), cudaSuccess)
#include "../../hts-train-world_b200/csrc/wb_dio.cu"
#undef cudaMemcpy

namespace wb {                                  // what wb_context.cu provides in the library
void set_error(const char* fmt, ...) { fprintf(stderr, "[emu] %s\n", fmt); }
bool check_cuda(cudaError_t e, const char*, const char*, int) { return e == cudaSuccess; }
cudaStream_t pool_stream() { return nullptr; }
}  // namespace wb

extern "C" int emu_dio(const double* x, int x_len, int fs, double f0_floor, double f0_ceil, double channels_in_octave,
                       double frame_period, double allowed_range, double* f0_out) {
  using namespace wb;
  if (fs != 48000) return 2;                    // the compile-time 8 192-point block of the 48 kHz path
  DioParams p{f0_floor, f0_ceil, channels_in_octave, frame_period, 1, allowed_range};
  DioFilterBank fb;
  if (!build_filter_bank((double)fs, p, &fb) || fb.log2bn != 13) return 3;
  DioConst c;
  c.nb = fb.nb; c.bn = fb.bn; c.log2bn = fb.log2bn; c.hN = fb.hN; c.D = fb.D; c.V = fb.V;
  for (int i = 0; i < fb.nb; ++i) { c.hal[i] = fb.hal[i]; c.boundary_f0[i] = fb.boundary_f0[i]; }
  c.actual_fs = fs; c.f0_floor = f0_floor; c.f0_ceil = f0_ceil; c.allowed_range = allowed_range;
  c.voice_range_minimum = static_cast<int>(0.5 + 1000.0 / frame_period / f0_floor) * 2 + 1;
  const int y_len = 1 + x_len;
  const int sample = y_len + 4 * static_cast<int>(1.0 + fs / fb.boundary_f0[0] / 2.0);
  const int mask = static_cast<int>(pow(2.0, static_cast<int>(log(static_cast<double>(sample)) / kLog2) + 1.0)) - 1;
  const int TF = static_cast<int>(1000.0 * x_len / fs / frame_period) + 1;       // GetSamplesForDIO
  std::vector<double> xs(x, x + x_len);
  xs.resize(x_len + 16, 0.0);
  const long long x_off = 0, f_off_ll = 0;
  const int f_off = 0;
  std::vector<double> frame_t(TF);
  for (int i = 0; i < TF; ++i) frame_t[i] = i * frame_period / 1000.0;
  std::vector<double> cand((size_t)c.nb * TF), score((size_t)c.nb * TF), tmp1(TF), tmp2(TF), f0(TF, 0.0);
  std::vector<int> pos(TF), neg(TF);
  double mean = 0.0;
  wbemu::smem_overruns = 0;
  wbemu::launch_grid(1, 1, 1024, 0, [&]() { dio_mean_kernel(xs.data(), &x_off, &x_len, &y_len, &mean); });
  // twiddles of the 8 192-point block
  std::vector<double2> tw(4097);
  for (int k = 0; k <= 4096; ++k) {
    const long double a = -2.0L * 3.14159265358979323846264338327950288L * k / 8192;
    tw[k] = make_double2((double)cosl(a), (double)sinl(a));
  }
  OlsConst oc = {c.nb, c.bn, c.log2bn, c.D, c.V};
  std::vector<int> shift(c.nb);
  for (int i = 0; i < c.nb; ++i) shift[i] = c.D + 2 * c.hal[i] + c.hN;
  const size_t tot = (size_t)c.nb * y_len + 8;
  std::vector<double> Fbuf(tot, 0.0);
  const int n_blocks = (y_len + c.V - 1) / c.V, n_chunks = (y_len + kZcChunk - 1) / kZcChunk, n_lists = c.nb * 4;
  const size_t smem = 2 * cpad_size(c.bn / 2) * sizeof(double2);                  // as dio_run
  wbemu::launch_grid(n_blocks, 1, 512, smem, [&]() {
    ols_filter_kernel<13, 512, 4>(xs.data(), &x_off, &x_len, &y_len, &mask, &mean, &f_off_ll, fb.G.p, tw.data(), oc, shift.data(), 0,
                                  Fbuf.data());
  });
  std::vector<int> counts((size_t)n_lists * n_chunks, 0), ltot(n_lists, 0);
  wbemu::launch_grid(n_chunks, c.nb, 256, 0,
                     [&]() { zc_kernel<false>(Fbuf.data(), &f_off_ll, &y_len, c.nb, 0, n_chunks, counts.data(), nullptr, nullptr); });
  wbemu::launch_grid((n_lists + 127) / 128, 1, 128, 0, [&]() { zc_scan_kernel(counts.data(), n_lists, n_chunks, ltot.data()); });
  std::vector<long long> loff(n_lists);
  long long etot = 0;
  for (int l = 0; l < n_lists; ++l) { loff[l] = etot; etot += ltot[l]; }
  std::vector<double> edges((size_t)etot + 2, 0.0);
  wbemu::launch_grid(n_chunks, c.nb, 256, 0,
                     [&]() { zc_kernel<true>(Fbuf.data(), &f_off_ll, &y_len, c.nb, 0, n_chunks, counts.data(), loff.data(), edges.data()); });
  // the default path of the library: zero crossings taken inside the filter kernel (ols_filter_zc_kernel),
  // block segments -> scan -> gather.  Its edge lists must hold the same events as the two-pass ones.
  {
    OlsConst ocz = oc;
    ocz.V = c.V - 2;
    const int nbz = (y_len - 1 + ocz.V - 1) / ocz.V;
    std::vector<int> segcnt((size_t)n_lists * nbz, 0), segoff((size_t)n_lists * nbz, 0), ltot2(n_lists + 1, 0);
    std::vector<double> seg((size_t)n_lists * nbz * kZcSegCap, -1.0);
    wbemu::launch_grid(nbz, 1, 512, smem, [&]() {
      ols_filter_zc_kernel<13, 512, 4>(xs.data(), &x_off, &x_len, &y_len, &mask, &mean, fb.G.p, tw.data(), ocz, shift.data(), 0, nbz, kZcSegCap,
                                       segcnt.data(), seg.data());
    });
    wbemu::launch_grid((n_lists + 127) / 128, 1, 128, 0,
                       [&]() { zc_seg_scan_kernel(segcnt.data(), n_lists, nbz, kZcSegCap, segoff.data(), ltot2.data(), ltot2.data() + n_lists); });
    if (ltot2[n_lists] != 0) return 5;                                           // a segment overflowed
    for (int l = 0; l < n_lists; ++l) if (ltot2[l] != ltot[l]) return 6;
    std::vector<double> edges2((size_t)etot + 2, 0.0);
    wbemu::launch_grid(n_lists, 1, 128, 0,
                       [&]() { zc_seg_gather_kernel(segcnt.data(), segoff.data(), seg.data(), nbz, kZcSegCap, loff.data(), edges2.data()); });
    // (a sample's block and position inside the block differ between the two partitions, so the transforms
    // round differently: the sub-sample positions agree to ~1e-12, not bit for bit)
    for (long long i = 0; i < etot; ++i)
      if (fabs(edges[i] - edges2[i]) > 1e-9 * (1.0 + fabs(edges[i]))) return 7;
    edges.swap(edges2);
  }
  wbemu::launch_grid((TF + 127) / 128, c.nb, 128, 0, [&]() {
    dio_candidates_kernel(edges.data(), loff.data(), ltot.data(), &f_off, &TF, frame_t.data(), c, 0, TF, cand.data(), score.data());
  });
  wbemu::launch_grid(1, 1, 256, 0, [&]() {
    dio_fix_kernel(cand.data(), score.data(), &f_off, &TF, c, 0, TF, tmp1.data(), tmp2.data(), pos.data(), neg.data(), f0.data());
  });
  memcpy(f0_out, f0.data(), TF * sizeof(double));
  return wbemu::smem_overruns ? 4 : 0;
}
