// TEST INFRASTRUCTURE: stand-alone driver of tests/emu/d4c_emu.cpp for the ThreadSanitizer build
// (an instrumented executable is simpler to run than an instrumented library inside python).
//   d4c_emu_main <dir> <mode> <threshold>: reads <dir>/{x,t,f0}.f64 and <dir>/rows.i32 (48 kHz,
//   fft_size 2048), writes <dir>/ap.f64, exits with emu_d4c's return code.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
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
int main(int argc, char** argv) {
  if (argc < 4) return 92;
  const std::string dir = argv[1];
  const auto x = slurp<double>(dir + "/x.f64"), t = slurp<double>(dir + "/t.f64"), f0 = slurp<double>(dir + "/f0.f64");
  const auto rows = slurp<int>(dir + "/rows.i32");
  std::vector<double> ap(rows.size() * 1025), ap0(f0.size());
  const int rc = emu_d4c(x.data(), (int)x.size(), 48000, t.data(), f0.data(), (int)f0.size(), 2048, atof(argv[3]), atoi(argv[2]),
                         rows.data(), (int)rows.size(), ap.data(), ap0.data());
  FILE* f = fopen((dir + "/ap.f64").c_str(), "wb");
  fwrite(ap.data(), sizeof(double), ap.size(), f);
  fclose(f);
  return rc;
}
