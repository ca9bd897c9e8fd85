// world-b200: link-compatibility helpers of libworld.a (HOST code, no kernels).
//
// The reference's static library also exports the small helpers of world/common.h,
// world/matlabfunctions.h, world/fft.h and the band-aperiodicity codec of world/codec.h
// (`nm -g libworld.a`; SURVEY.md 8b: "keep them exported so third-party callers link").  None of
// them is on the analysis / synthesis path of this library: Dio, StoneMask, CheapTrick, D4C,
// Synthesis, Harvest and the spectral-envelope codec run in the CUDA kernels and do not call
// anything in this file, and nothing here is a fallback for them.  They operate on a few hundred
// to a few thousand doubles in caller-owned host memory, which is why they stay on the host.
//
// Reference behaviour restated here (file:line of /root/reference/externs/WORLD_v2/src):
//   fft.cpp:26-166 (plan API and its conventions), common.cpp:27-226, matlabfunctions.cpp:27-325,
//   codec.cpp:21-55,216-264.  The transform itself is an ordinary iterative radix-2 FFT, not the
//   reference's vendored split-radix code; results agree to rounding (tests/test_compat.py).
#include <math.h>
#include <stdint.h>
#include <algorithm>
#include <vector>
#include "../../include/world/codec.h"
#include "../../include/world/constantnumbers.h"
#include "../../include/world/matlabfunctions.h"
#include "wb_batch.h"

namespace {

// ---- plain complex FFT on interleaved doubles ---------------------------------------------------
// X[k] = sum_j a[j] exp(dir * 2 pi i j k / n), unnormalised, in place.  rev: bit-reversal table,
// tw: (cos, sin)(2 pi k / n) for k < n/2.
void fill_tables(int n, int* rev, double* tw) {
  int bits = 0;
  while ((1 << bits) < n) ++bits;
  for (int i = 0; i < n; ++i) {
    int r = 0;
    for (int b = 0; b < bits; ++b) r |= ((i >> b) & 1) << (bits - 1 - b);
    rev[i] = r;
  }
  const long double step = 2.0L * 3.14159265358979323846264338327950288L / n;
  for (int k = 0; k < n / 2; ++k) {
    tw[2 * k] = static_cast<double>(cosl(step * k));
    tw[2 * k + 1] = static_cast<double>(sinl(step * k));
  }
}

void transform(double* a, int n, int dir, const int* rev, const double* tw) {
  for (int i = 0; i < n; ++i) {
    const int r = rev[i];
    if (r > i) {
      std::swap(a[2 * i], a[2 * r]);
      std::swap(a[2 * i + 1], a[2 * r + 1]);
    }
  }
  for (int half = 1; half < n; half <<= 1) {
    const int stride = n / (2 * half);
    for (int base = 0; base < n; base += 2 * half) {
      for (int j = 0; j < half; ++j) {
        const double wr = tw[2 * j * stride], wi = dir * tw[2 * j * stride + 1];
        double* lo = a + 2 * (base + j);
        double* hi = a + 2 * (base + j + half);
        const double tr = hi[0] * wr - hi[1] * wi, ti = hi[0] * wi + hi[1] * wr;
        hi[0] = lo[0] - tr;
        hi[1] = lo[1] - ti;
        lo[0] += tr;
        lo[1] += ti;
      }
    }
  }
}

fft_plan make_plan(int n, int sign, unsigned int flags) {
  fft_plan p;
  p.n = n;
  p.sign = sign;
  p.flags = flags;
  p.c_in = nullptr;
  p.in = nullptr;
  p.c_out = nullptr;
  p.out = nullptr;
  p.input = new double[2 * (n > 0 ? n : 1)];
  p.ip = new int[n > 0 ? n : 1];
  p.w = new double[n > 4 ? n * 5 / 4 : 5];
  if (n > 0) fill_tables(n, p.ip, p.w);
  return p;
}

uint32_t g_rng[4] = {123456789u, 362436069u, 521288629u, 88675123u};

inline uint32_t rng_next() {     // xorshift128
  const uint32_t t = g_rng[0] ^ (g_rng[0] << 11);
  g_rng[0] = g_rng[1];
  g_rng[1] = g_rng[2];
  g_rng[2] = g_rng[3];
  g_rng[3] = (g_rng[3] ^ (g_rng[3] >> 19)) ^ (t ^ (t >> 8));
  return g_rng[3];
}

// one pass of the 3rd-order low-pass of decimate(), direct form II with the state w
void iir_pass(const double* x, int n, const double* a, const double* b, double* y) {
  double w0 = 0.0, w1 = 0.0, w2 = 0.0;
  for (int i = 0; i < n; ++i) {
    const double wt = x[i] + a[0] * w0 + a[1] * w1 + a[2] * w2;
    y[i] = b[0] * wt + b[1] * w0 + b[1] * w1 + b[0] * w2;
    w2 = w1;
    w1 = w0;
    w0 = wt;
  }
}

}  // namespace

extern "C" {

// ---- world/fft.h -------------------------------------------------------------------------------
fft_plan fft_plan_dft_1d(int n, fft_complex* in, fft_complex* out, int sign, unsigned int flags) {
  fft_plan p = make_plan(n, sign, flags);
  p.c_in = in;
  p.c_out = out;
  return p;
}

fft_plan fft_plan_dft_c2r_1d(int n, fft_complex* in, double* out, unsigned int flags) {
  fft_plan p = make_plan(n, FFT_BACKWARD, flags);
  p.c_in = in;
  p.out = out;
  return p;
}

fft_plan fft_plan_dft_r2c_1d(int n, double* in, fft_complex* out, unsigned int flags) {
  fft_plan p = make_plan(n, FFT_FORWARD, flags);
  p.in = in;
  p.c_out = out;
  return p;
}

void fft_execute(fft_plan p) {
  const int n = p.n;
  double* a = p.input;
  if (n <= 0) return;
  if (p.c_in != nullptr && p.c_out != nullptr) {
    // c2c: the reference hands the data to its exp(+-i) kernel and conjugates the result, which is
    // the DFT (forward) or n * IDFT (backward) of the CONJUGATED input (fft.cpp:36-45,61-71)
    for (int i = 0; i < n; ++i) { a[2 * i] = p.c_in[i][0]; a[2 * i + 1] = -p.c_in[i][1]; }
    transform(a, n, p.sign == FFT_FORWARD ? -1 : 1, p.ip, p.w);
    for (int i = 0; i < n; ++i) { p.c_out[i][0] = a[2 * i]; p.c_out[i][1] = a[2 * i + 1]; }
  } else if (p.sign == FFT_FORWARD) {       // r2c: bins 0 .. n/2 (fft.cpp:49-60)
    for (int i = 0; i < n; ++i) { a[2 * i] = p.in[i]; a[2 * i + 1] = 0.0; }
    transform(a, n, -1, p.ip, p.w);
    for (int k = 0; k <= n / 2; ++k) { p.c_out[k][0] = a[2 * k]; p.c_out[k][1] = a[2 * k + 1]; }
    p.c_out[0][1] = 0.0;
    p.c_out[n / 2][1] = 0.0;
  } else {                                  // c2r: Hermitian extension of bins 0 .. n/2 (fft.cpp:27-35)
    a[0] = p.c_in[0][0];
    a[1] = 0.0;
    for (int k = 1; k < n / 2; ++k) {
      a[2 * k] = p.c_in[k][0];
      a[2 * k + 1] = p.c_in[k][1];
      a[2 * (n - k)] = p.c_in[k][0];
      a[2 * (n - k) + 1] = -p.c_in[k][1];
    }
    if (n > 1) { a[n] = p.c_in[n / 2][0]; a[n + 1] = 0.0; }
    transform(a, n, 1, p.ip, p.w);
    for (int i = 0; i < n; ++i) p.out[i] = a[2 * i];
  }
}

void fft_destroy_plan(fft_plan p) {
  delete[] p.input;
  delete[] p.ip;
  delete[] p.w;
}

// ---- world/common.h ----------------------------------------------------------------------------
int GetSuitableFFTSize(int sample) {
  return static_cast<int>(pow(2.0, static_cast<int>(log(static_cast<double>(sample)) / world::kLog2) + 1.0));
}

void NuttallWindow(int y_length, double* y) {
  for (int i = 0; i < y_length; ++i) {
    const double t = i / (y_length - 1.0);
    y[i] = 0.355768 - 0.487396 * cos(2.0 * world::kPi * t) + 0.144232 * cos(4.0 * world::kPi * t) -
           0.012604 * cos(6.0 * world::kPi * t);
  }
}

// bins below f0 receive the spectrum mirrored about f0 / 2: output[i] = input[i] + input(f0 - i df),
// the second term read by uniform-grid interpolation on the axis that starts at f0 and steps by -df
// (common.cpp:56-75; only the first 1 + int(f0 fft_size / fs) bins are written)
void DCCorrection(const double* input, double f0, int fs, int fft_size, double* output) {
  const int upper_limit = 2 + static_cast<int>(f0 * fft_size / fs);
  std::vector<double> axis(upper_limit), replica(upper_limit);
  for (int i = 0; i < upper_limit; ++i) axis[i] = static_cast<double>(i) * fs / fft_size;
  interp1Q(f0 - axis[0], -static_cast<double>(fs) / fft_size, input, upper_limit + 1, axis.data(),
           upper_limit - 1, replica.data());
  for (int i = 0; i < upper_limit - 1; ++i) output[i] = input[i] + replica[i];
}

// rectangular smoothing of `width` Hz as a difference of the running integral of the spectrum,
// mirrored by `boundary` bins at both ends (common.cpp:27-46,77-111)
void LinearSmoothing(const double* input, double width, int fs, int fft_size, double* output) {
  const int half = fft_size / 2;
  const int boundary = static_cast<int>(width * fft_size / fs) + 1;
  const int n = half + 2 * boundary + 1;
  std::vector<double> integral(n), axis(half + 1), low(half + 1), high(half + 1);
  double run = 0.0;
  for (int i = 0; i < n; ++i) {
    const int k = i - boundary;                              // bin on the unmirrored axis
    const double v = input[k < 0 ? -k : (k > half ? 2 * half - k : k)];
    run = i == 0 ? v * fs / fft_size : v * fs / fft_size + run;
    integral[i] = run;
  }
  for (int i = 0; i <= half; ++i) axis[i] = static_cast<double>(i) / fft_size * fs - width / 2.0;
  const double origin = -(boundary - 0.5) * fs / fft_size;
  const double df = static_cast<double>(fs) / fft_size;
  interp1Q(origin, df, integral.data(), n, axis.data(), half + 1, low.data());
  for (int i = 0; i <= half; ++i) axis[i] += width;
  interp1Q(origin, df, integral.data(), n, axis.data(), half + 1, high.data());
  for (int i = 0; i <= half; ++i) output[i] = (high[i] - low[i]) / width;
}

void InitializeForwardRealFFT(int fft_size, ForwardRealFFT* f) {
  f->fft_size = fft_size;
  f->waveform = new double[fft_size];
  f->spectrum = new fft_complex[fft_size];
  f->forward_fft = fft_plan_dft_r2c_1d(fft_size, f->waveform, f->spectrum, FFT_ESTIMATE);
}

void DestroyForwardRealFFT(ForwardRealFFT* f) {
  fft_destroy_plan(f->forward_fft);
  delete[] f->spectrum;
  delete[] f->waveform;
}

void InitializeInverseRealFFT(int fft_size, InverseRealFFT* f) {
  f->fft_size = fft_size;
  f->waveform = new double[fft_size];
  f->spectrum = new fft_complex[fft_size];
  f->inverse_fft = fft_plan_dft_c2r_1d(fft_size, f->spectrum, f->waveform, FFT_ESTIMATE);
}

void DestroyInverseRealFFT(InverseRealFFT* f) {
  fft_destroy_plan(f->inverse_fft);
  delete[] f->spectrum;
  delete[] f->waveform;
}

void InitializeInverseComplexFFT(int fft_size, InverseComplexFFT* f) {
  f->fft_size = fft_size;
  f->input = new fft_complex[fft_size];
  f->output = new fft_complex[fft_size];
  f->inverse_fft = fft_plan_dft_1d(fft_size, f->input, f->output, FFT_BACKWARD, FFT_ESTIMATE);
}

void DestroyInverseComplexFFT(InverseComplexFFT* f) {
  fft_destroy_plan(f->inverse_fft);
  delete[] f->input;
  delete[] f->output;
}

void InitializeMinimumPhaseAnalysis(int fft_size, MinimumPhaseAnalysis* m) {
  m->fft_size = fft_size;
  m->log_spectrum = new double[fft_size];
  m->minimum_phase_spectrum = new fft_complex[fft_size];
  m->cepstrum = new fft_complex[fft_size];
  m->inverse_fft = fft_plan_dft_r2c_1d(fft_size, m->log_spectrum, m->cepstrum, FFT_ESTIMATE);
  m->forward_fft = fft_plan_dft_1d(fft_size, m->cepstrum, m->minimum_phase_spectrum, FFT_FORWARD, FFT_ESTIMATE);
}

void DestroyMinimumPhaseAnalysis(MinimumPhaseAnalysis* m) {
  fft_destroy_plan(m->forward_fft);
  fft_destroy_plan(m->inverse_fft);
  delete[] m->cepstrum;
  delete[] m->log_spectrum;
  delete[] m->minimum_phase_spectrum;
}

// log spectrum (even) -> cepstrum -> causal fold (x2 for 0 < i < n/2, 0 beyond n/2) -> spectrum ->
// exp; the conjugations mirror the plan conventions above (common.cpp:182-217)
void GetMinimumPhaseSpectrum(const MinimumPhaseAnalysis* m) {
  const int n = m->fft_size, h = n / 2;
  for (int i = 1; i < h; ++i) m->log_spectrum[n - i] = m->log_spectrum[i];
  fft_execute(m->inverse_fft);
  fft_complex* c = m->cepstrum;
  for (int i = 0; i < n; ++i) {
    const double g = (i == 0 || i == h) ? 1.0 : (i < h ? 2.0 : 0.0);
    c[i][0] = i > h ? 0.0 : c[i][0] * g;
    c[i][1] = i > h ? 0.0 : c[i][1] * -g;
  }
  fft_execute(m->forward_fft);
  fft_complex* s = m->minimum_phase_spectrum;
  for (int i = 0; i <= h; ++i) {
    const double mag = exp(s[i][0] / n), ph = s[i][1] / n;
    s[i][0] = mag * cos(ph);
    s[i][1] = mag * sin(ph);
  }
}

// ---- world/matlabfunctions.h -------------------------------------------------------------------
void fftshift(const double* x, int x_length, double* y) {
  const int h = x_length / 2;
  for (int i = 0; i < h; ++i) {
    y[i] = x[i + h];
    y[i + h] = x[i];
  }
}

// index[i] = number of knots at or below edges[i], clamped to [1, x_length - 1]; the cursor only
// moves forward, which is what the reference's merge loop does (matlabfunctions.cpp:136-155)
void histc(const double* x, int x_length, const double* edges, int edges_length, int* index) {
  int cursor = 1;
  for (int i = 0; i < edges_length; ++i) {
    while (cursor < x_length && !(edges[i] < x[cursor])) ++cursor;
    if (cursor >= x_length) {
      for (; i < edges_length; ++i) index[i] = x_length - 1;
      return;
    }
    index[i] = cursor;
  }
}

void interp1(const double* x, const double* y, int x_length, const double* xi, int xi_length, double* yi) {
  std::vector<int> seg(xi_length > 0 ? xi_length : 1, 0);
  histc(x, x_length, xi, xi_length, seg.data());
  for (int i = 0; i < xi_length; ++i) {
    const int k = seg[i];
    const double s = (xi[i] - x[k - 1]) / (x[k] - x[k - 1]);
    yi[i] = y[k - 1] + s * (y[k] - y[k - 1]);
  }
}

// zero-phase decimation: 9-sample odd reflection at both ends, the low-pass forward and
// backward, every r-th sample (matlabfunctions.cpp:184-210)
void decimate(const double* x, int x_length, int r, double* y) {
  const int pad = 9, n = x_length + 2 * pad;
  double a[3] = {0.0, 0.0, 0.0}, b[2] = {0.0, 0.0};
  wb::decimate_filter_coefficients(r, a, b);          // zeros for r outside 2..12, as in the reference
  std::vector<double> u(n), v(n);
  for (int i = 0; i < n; ++i) {
    const int j = i - pad;
    u[i] = j < 0 ? 2 * x[0] - x[-j] : (j < x_length ? x[j] : 2 * x[x_length - 1] - x[2 * (x_length - 1) - j]);
  }
  iir_pass(u.data(), n, a, b, v.data());
  for (int i = 0; i < n; ++i) u[i] = v[n - 1 - i];
  iir_pass(u.data(), n, a, b, v.data());
  const int nout = (x_length - 1) / r + 1;
  const int nbeg = r - r * nout + x_length;
  int count = 0;
  for (int i = nbeg; i < x_length + pad; i += r) y[count++] = v[n - 1 - (i + pad - 1)];
}

int matlab_round(double x) { return x > 0 ? static_cast<int>(x + 0.5) : static_cast<int>(x - 0.5); }

void diff(const double* x, int x_length, double* y) {
  for (int i = 0; i + 1 < x_length; ++i) y[i] = x[i + 1] - x[i];
}

// knots at x + k shift; base index by truncation, y treated as constant beyond its last knot
void interp1Q(double x, double shift, const double* y, int x_length, const double* xi, int xi_length, double* yi) {
  for (int i = 0; i < xi_length; ++i) {
    const double pos = (xi[i] - x) / shift;
    const int base = static_cast<int>(pos);
    const double frac = pos - base;
    const double dy = base < x_length - 1 ? y[base + 1] - y[base] : 0.0;
    yi[i] = y[base] + dy * frac;
  }
}

void randn_reseed(void) {
  g_rng[0] = 123456789u;
  g_rng[1] = 362436069u;
  g_rng[2] = 521288629u;
  g_rng[3] = 88675123u;
}

double randn(void) {      // sum of 12 uniform 28-bit draws, centred
  uint32_t acc = 0;
  for (int i = 0; i < 12; ++i) acc += rng_next() >> 4;
  return acc / 268435456.0 - 6.0;
}

// y[0 .. fft_size) = circular convolution of x and h, both scaled by 1 / fft_size before their
// transforms (matlabfunctions.cpp:279-313)
void fast_fftfilt(const double* x, int x_length, const double* h, int h_length, int fft_size,
                  const ForwardRealFFT* fwd, const InverseRealFFT* inv, double* y) {
  std::vector<double> xs(2 * (fft_size / 2 + 1));
  for (int i = 0; i < fft_size; ++i) fwd->waveform[i] = i < x_length ? x[i] / fft_size : 0.0;
  fft_execute(fwd->forward_fft);
  for (int k = 0; k <= fft_size / 2; ++k) { xs[2 * k] = fwd->spectrum[k][0]; xs[2 * k + 1] = fwd->spectrum[k][1]; }
  for (int i = 0; i < fft_size; ++i) fwd->waveform[i] = i < h_length ? h[i] / fft_size : 0.0;
  fft_execute(fwd->forward_fft);
  for (int k = 0; k <= fft_size / 2; ++k) {
    const double hr = fwd->spectrum[k][0], hi = fwd->spectrum[k][1];
    inv->spectrum[k][0] = xs[2 * k] * hr - xs[2 * k + 1] * hi;
    inv->spectrum[k][1] = xs[2 * k] * hi + xs[2 * k + 1] * hr;
  }
  fft_execute(inv->inverse_fft);
  for (int i = 0; i < fft_size; ++i) y[i] = inv->waveform[i];
}

double matlab_std(const double* x, int x_length) {
  double mean = 0.0;
  for (int i = 0; i < x_length; ++i) mean += x[i];
  mean /= x_length;
  double s = 0.0;
  for (int i = 0; i < x_length; ++i) s += pow(x[i] - mean, 2.0);
  return sqrt(s / (x_length - 1));
}

// ---- world/codec.h: band aperiodicity ---------------------------------------------------------------
void CodeAperiodicity(const double* const* aperiodicity, int f0_length, int fs, int fft_size,
                      int number_of_aperiodicities, double** coded_aperiodicity) {
  const int rows = fft_size / 2 + 1;
  std::vector<double> centres(number_of_aperiodicities > 0 ? number_of_aperiodicities : 1), db(rows);
  for (int b = 0; b < number_of_aperiodicities; ++b) centres[b] = world::kFrequencyInterval * (b + 1.0);
  for (int f = 0; f < f0_length; ++f) {
    for (int k = 0; k < rows; ++k) db[k] = 20 * log10(aperiodicity[f][k]);
    interp1Q(0, static_cast<double>(fs) / fft_size, db.data(), rows, centres.data(), number_of_aperiodicities,
             coded_aperiodicity[f]);
  }
}

// NB parameter order: (fs, number_of_aperiodicities, fft_size), see include/world/codec.h
void DecodeAperiodicity(const double* const* coded_aperiodicity, int f0_length, int fs,
                        int number_of_aperiodicities, int fft_size, double** aperiodicity) {
  const int rows = fft_size / 2 + 1, nb = number_of_aperiodicities;
  std::vector<double> axis(rows), knots_x(nb + 2), knots_y(nb + 2);
  for (int k = 0; k < rows; ++k) axis[k] = static_cast<double>(fs) / fft_size * k;
  for (int b = 0; b <= nb; ++b) knots_x[b] = b * world::kFrequencyInterval;
  knots_x[nb + 1] = fs / 2.0;
  knots_y[0] = -60.0;
  knots_y[nb + 1] = -world::kMySafeGuardMinimum;
  for (int f = 0; f < f0_length; ++f) {
    double mean = 0.0;
    for (int b = 0; b < nb; ++b) {
      mean += coded_aperiodicity[f][b];
      knots_y[b + 1] = coded_aperiodicity[f][b];
    }
    mean /= nb;
    if (mean > -0.5) {                                  // unvoiced frame (codec.cpp:33-43): fully aperiodic
      for (int k = 0; k < rows; ++k) aperiodicity[f][k] = 1.0 - world::kMySafeGuardMinimum;
      continue;
    }
    interp1(knots_x.data(), knots_y.data(), nb + 2, axis.data(), rows, aperiodicity[f]);
    for (int k = 0; k < rows; ++k) aperiodicity[f][k] = pow(10.0, aperiodicity[f][k] / 20.0);
  }
}

}  // extern "C"
