// TEST INFRASTRUCTURE: stand-alone driver of tests/emu/{d4c,cheaptrick}_emu.cpp for the
// ThreadSanitizer builds (an instrumented executable is simpler to run than an instrumented library
// inside python).  Reads <dir>/{x,t,f0}.f64 and <dir>/rows.i32 (48 kHz, fft_size 2048):
//   emu_main <dir> <mode> <threshold>      D4C (default build): writes <dir>/ap.f64
//   emu_main <dir>                         CheapTrick (-DEMU_CHEAPTRICK): writes <dir>/sp.f64
//   emu_main <dir> <fs>                    StoneMask (-DEMU_STONEMASK): f0.f64 = raw F0, writes <dir>/f0_refined.f64
//   emu_main <dir>                         Dio (-DEMU_DIO): x.f64 only, writes <dir>/f0_raw.f64
//   emu_main <dir>                         Synthesis (-DEMU_SYNTHESIS): f0.f64, sp.f64, ap.f64 (1025 bins), writes <dir>/y.f64
// and exits with the harness's return code.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
extern "C" int emu_cheaptrick(const double* x, int x_len, int fs, const double* t, const double* f0, int F, int fft_size,
                              double q1, const int* rows, int n_rows, double* sp_rows);
extern "C" int emu_stonemask(const double* x, int x_len, int fs, const double* t, const double* f0, int F, const int* rows,
                             int n_rows, double* f0_rows);
extern "C" int emu_synthesis(const double* f0, int F, const double* sp, const double* ap, int fft_size, double frame_period_ms,
                             int fs, int y_length, double* y_out);
extern "C" int emu_dio(const double* x, int x_len, int fs, double f0_floor, double f0_ceil, double channels_in_octave,
                       double frame_period, double allowed_range, double* f0_out);
extern "C" int emu_d4c(const double* x, int x_len, int fs, const double* t, const double* f0, int F, int fft_size,
                       double threshold, int mode, const int* rows, int n_rows, double* ap_rows, double* ap0_out);
template <typename T>
static std::vector<T> slurp(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { perror(path.c_str()); exit(90); }
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<T> v(n / sizeof(T));
  if (fread(v.data(), 1, n, f) != (size_t)n) exit(91);
  fclose(f);
  return v;
}
static void dump(const std::string& path, const std::vector<double>& v) {
  FILE* f = fopen(path.c_str(), "wb");
  fwrite(v.data(), sizeof(double), v.size(), f);
  fclose(f);
}
int main(int argc, char** argv) {
  if (argc < 2) return 92;
  const std::string dir = argv[1];
#ifdef EMU_DIO
  {
    const auto x = slurp<double>(dir + "/x.f64");
    std::vector<double> f0(static_cast<int>(1000.0 * x.size() / 48000 / 5.0) + 1);
    const int rc = emu_dio(x.data(), (int)x.size(), 48000, 71.0, 800.0, 2.0, 5.0, 0.1, f0.data());
    dump(dir + "/f0_raw.f64", f0);
    return rc;
  }
#endif
#ifdef EMU_SYNTHESIS
  {
    const auto f0 = slurp<double>(dir + "/f0.f64"), sp = slurp<double>(dir + "/sp.f64"), ap = slurp<double>(dir + "/ap.f64");
    const int F = (int)f0.size(), y_length = static_cast<int>((F - 1) * 5.0 / 1000.0 * 48000) + 1;
    std::vector<double> y(y_length);
    const int rc = emu_synthesis(f0.data(), F, sp.data(), ap.data(), 2048, 5.0, 48000, y_length, y.data());
    dump(dir + "/y.f64", y);
    return rc;
  }
#endif
  const auto x = slurp<double>(dir + "/x.f64"), t = slurp<double>(dir + "/t.f64"), f0 = slurp<double>(dir + "/f0.f64");
  const auto rows = slurp<int>(dir + "/rows.i32");
  std::vector<double> out(rows.size() * 1025);
#if defined(EMU_SYNTHESIS) || defined(EMU_DIO)
  const int rc = 0;
#elif defined(EMU_STONEMASK)
  if (argc < 3) return 92;
  out.resize(rows.size());
  const int rc = emu_stonemask(x.data(), (int)x.size(), atoi(argv[2]), t.data(), f0.data(), (int)f0.size(), rows.data(),
                               (int)rows.size(), out.data());
  dump(dir + "/f0_refined.f64", out);
#elif defined(EMU_CHEAPTRICK)
  const int rc = emu_cheaptrick(x.data(), (int)x.size(), 48000, t.data(), f0.data(), (int)f0.size(), 2048, -0.15, rows.data(),
                                (int)rows.size(), out.data());
  dump(dir + "/sp.f64", out);
#else
  if (argc < 4) return 92;
  std::vector<double> ap0(f0.size());
  const int rc = emu_d4c(x.data(), (int)x.size(), 48000, t.data(), f0.data(), (int)f0.size(), 2048, atof(argv[3]), atoi(argv[2]),
                         rows.data(), (int)rows.size(), out.data(), ap0.data());
  dump(dir + "/ap.f64", out);
#endif
  return rc;
}
