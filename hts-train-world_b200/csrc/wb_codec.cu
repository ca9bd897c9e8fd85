// world-b200: mel-DCT coding / decoding of spectral envelopes (the analysis tool's codec tail).
//
// Reference: W/src/codec.cpp — CodeSpectralEnvelope :266-295, CodeOneFrame :122-134, DCTForCodec
// :72-88, GetParametersForCoding :161-179, DecodeSpectralEnvelope :297-324, DecodeOneFrame
// :139-156, IDCTForCodec :93-117, GetParametersForDecoding :184-209; interp1
// W/src/matlabfunctions.cpp:157-182.  The tool-level scalings (sp x 1e4, zero -> 1e-4, c0 + 12;
// ap x 1e4, c0 - 9.21034; lf0 = log f0, 0 stays 0; float32 outputs) are W/test/analysis.cpp
// :293-390.
//
// One CTA per frame: the row is read once from HBM, log / interp1 onto the mel grid / DCT (one
// real FFT of fft_size/2 points) happen in shared memory, and only number_of_dimensions values
// go back -- 20x less device-to-host traffic than the raw double rows.  The interp1 segment
// index and weight of every mel point depend only on (fs, fft_size) and come from a table
// computed once on the host with the reference's histc semantics.
#include <math.h>
#include <algorithm>
#include <map>
#include <vector>
#include "wb_batch.h"
#include "wb_fft.cuh"
#include "wb_zerocross.cuh"      // host_fft

namespace wb {
namespace {

constexpr double kM0 = 1127.01048, kF0mel = 700.0, kFloorFrequency = 40.0, kCeilFrequency = 20000.0;
inline double freq_to_mel(double f) { return kM0 * log(f / kF0mel + 1.0); }
inline double mel_to_freq(double m) { return kF0mel * (exp(m / kM0) - 1.0); }

struct CodecTables {
  int fs = 0, fft_size = 0, ndim = 0, max_dim = 0, log2max = 0;
  DevBuf<int> enc_idx;        // [max_dim] lower knot (0-based) of every mel point
  DevBuf<double> enc_s;       // [max_dim] interpolation weight
  DevBuf<double2> enc_w;      // [ndim] DCT weights (:169-174)
  DevBuf<int> dec_idx;        // [fft_size/2 + 1] lower knot in the (max_dim + 2)-point mel spectrum
  DevBuf<double> dec_s;
  DevBuf<double2> dec_w;      // [ndim] IDCT weights (:193-197)
};

// interp1 segment of one query: k = clamp(upper_bound(x, xi), 1, n - 1),
// s = (xi - x[k-1]) / (x[k] - x[k-1]).  Returns k - 1.
int interp_segment(const std::vector<double>& x, double xi, double* s) {
  const int n = (int)x.size();
  int k = (int)(std::upper_bound(x.begin(), x.end(), xi) - x.begin());
  k = std::max(1, std::min(n - 1, k));
  *s = (xi - x[k - 1]) / (x[k] - x[k - 1]);
  return k - 1;
}

std::map<std::vector<int>, CodecTables*> g_codec;

CodecTables* codec_tables(int fs, int fft_size, int ndim) {
  const std::vector<int> key = {fs, fft_size, ndim};
  auto it = g_codec.find(key);
  if (it != g_codec.end()) return it->second;
  const int max_dim = fft_size / 2;
  int log2max = 0;
  while ((1 << log2max) < max_dim) ++log2max;
  if ((1 << log2max) != max_dim || log2max < 4 || log2max > 13 || ndim < 1 || ndim > max_dim) {
    set_error("codec: unsupported fft_size %d / dimensions %d", fft_size, ndim);
    return nullptr;
  }
  CodecTables* t = new CodecTables();
  t->fs = fs; t->fft_size = fft_size; t->ndim = ndim; t->max_dim = max_dim; t->log2max = log2max;
  const double floor_mel = freq_to_mel(kFloorFrequency);
  const double ceil_mel = freq_to_mel(std::min(fs / 2.0, kCeilFrequency));
  // ---- coding: knots = mel(bin frequencies); the reference leaves the last of its fft_size/2+1
  // knots uninitialised, it is never reached by the mel grid -> +inf here
  std::vector<double> knots(max_dim + 1);
  for (int i = 0; i < max_dim; ++i) knots[i] = freq_to_mel(static_cast<double>(i) * fs / fft_size);
  knots[max_dim] = HUGE_VAL;
  std::vector<int> idx(max_dim);
  std::vector<double> s(max_dim);
  for (int i = 0; i < max_dim; ++i) {
    const double mel = (ceil_mel - floor_mel) * i / max_dim + floor_mel;
    idx[i] = interp_segment(knots, mel, &s[i]);
    if (idx[i] + 1 >= max_dim) { idx[i] = max_dim - 2; s[i] = (mel - knots[max_dim - 2]) / (knots[max_dim - 1] - knots[max_dim - 2]); }
  }
  std::vector<double2> w(ndim), wd(ndim);
  for (int i = 0; i < ndim; ++i) {
    w[i] = make_double2(2.0 * cos(i * kPi / fft_size) / sqrt((double)fft_size), 2.0 * sin(i * kPi / fft_size) / sqrt((double)fft_size));
    wd[i] = make_double2(cos(i * kPi / fft_size) * sqrt((double)fft_size), sin(i * kPi / fft_size) * sqrt((double)fft_size));
  }
  w[0].x /= sqrt(2.0);
  wd[0].x /= sqrt(2.0);
  // ---- decoding: knots in Hz = {0, mel_to_freq(grid), fs/2}, queries = bin frequencies
  std::vector<double> dk(max_dim + 2);
  dk[0] = 0.0;
  for (int i = 0; i < max_dim; ++i) dk[i + 1] = mel_to_freq((ceil_mel - floor_mel) * i / max_dim + floor_mel);
  dk[max_dim + 1] = fs / 2.0;
  std::vector<int> didx(fft_size / 2 + 1);
  std::vector<double> ds(fft_size / 2 + 1);
  for (int i = 0; i <= fft_size / 2; ++i) didx[i] = interp_segment(dk, static_cast<double>(i) * fs / fft_size, &ds[i]);
  bool ok = t->enc_idx.alloc(max_dim) && t->enc_s.alloc(max_dim) && t->enc_w.alloc(ndim) &&
            t->dec_idx.alloc(didx.size()) && t->dec_s.alloc(ds.size()) && t->dec_w.alloc(ndim);
  ok = ok && WB_CUDA(cudaMemcpy(t->enc_idx.p, idx.data(), max_dim * sizeof(int), cudaMemcpyHostToDevice)) &&
       WB_CUDA(cudaMemcpy(t->enc_s.p, s.data(), max_dim * sizeof(double), cudaMemcpyHostToDevice)) &&
       WB_CUDA(cudaMemcpy(t->enc_w.p, w.data(), ndim * sizeof(double2), cudaMemcpyHostToDevice)) &&
       WB_CUDA(cudaMemcpy(t->dec_idx.p, didx.data(), didx.size() * sizeof(int), cudaMemcpyHostToDevice)) &&
       WB_CUDA(cudaMemcpy(t->dec_s.p, ds.data(), ds.size() * sizeof(double), cudaMemcpyHostToDevice)) &&
       WB_CUDA(cudaMemcpy(t->dec_w.p, wd.data(), ndim * sizeof(double2), cudaMemcpyHostToDevice));
  if (!ok) { delete t; return nullptr; }
  g_codec[key] = t;
  return t;
}

// dynamic shared memory: [ buf: cpad_size(max_dim/2) double2 | logsp: max_dim + 2 doubles ]
// out[f][d] = DCT coefficient d of frame f (+ c0_add on d = 0).  scale multiplies the row first;
// a scaled value of exactly 0 becomes zero_floor (W/test/analysis.cpp:297-301); pass 0 to disable.
// F32LOG: the logarithm of every bin is taken in FP32 (1 ulp, ~1e-6 absolute on log spectra of
// magnitude <= 20) -- the batch path, whose outputs are the tool's float32 files (the coefficients
// move by < 1e-6, their own float32 spacing is 1e-6 at c0 ~ 14); the drop-in CodeSpectralEnvelope,
// which returns doubles, keeps the FP64 logarithm.
template <int LOG2MAX, bool F32LOG = false>     // LOG2MAX = log2(fft_size / 2); 0: given at run time
__global__ void __launch_bounds__(128)
codec_encode_kernel(const double* __restrict__ rows, int half, int log2max_rt, const int* __restrict__ idx,
                    const double* __restrict__ s, const double2* __restrict__ weight, int ndim, double scale,
                    double zero_floor, double c0_add, const double2* __restrict__ tw, double* __restrict__ out) {
  extern __shared__ double2 smem2[];
  const int log2max = LOG2MAX > 0 ? LOG2MAX : log2max_rt;
  constexpr int LM = LOG2MAX > 0 ? LOG2MAX - 1 : 0;
  constexpr int TWL = LOG2MAX > 0 ? LOG2MAX : kTwLog2;   // compact twiddle table of this size, or the master table
  const int max_dim = 1 << log2max, M = max_dim >> 1, log2m = log2max - 1;
  double2* buf = smem2;
  double* bufd = reinterpret_cast<double*>(buf);
  double* logsp = reinterpret_cast<double*>(buf + cpad_size(M));
  const int tid = threadIdx.x, T = 128;
  const size_t f = blockIdx.x;
  const double* __restrict__ row = rows + f * (half + 1);
  auto log_of = [&](double raw) {
    double v = raw * scale;
    if (zero_floor != 0.0 && v == 0.0) v = zero_floor;
    return F32LOG ? static_cast<double>(logf(static_cast<float>(v))) : log(v);
  };
  if constexpr (LOG2MAX > 0) {
    // the whole row is requested before the first logarithm: one HBM round trip per frame instead of
    // one per loop iteration (the kernel was bound by the latency of these loads)
    constexpr int kPer = (1 << LOG2MAX) / 128;
    double raw[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) raw[j] = row[tid + j * T];
#pragma unroll
    for (int j = 0; j < kPer; ++j) logsp[tid + j * T] = log_of(raw[j]);
  } else {
    for (int k = tid; k < max_dim; k += T) logsp[k] = log_of(row[k]);   // bins 0 .. max_dim-1 are all interp1 ever reads
  }
  __syncthreads();
  // mel spectrum j -> DCT input position (:75-79): even j -> j/2, odd j -> max_dim - 1 - (j-1)/2
  for (int j = tid; j < max_dim; j += T) {
    const int k = idx[j];
    const double y0 = logsp[k];
    const double v = add_rn(y0, mul_rn(s[j], add_rn(logsp[k + 1], -y0)));
    const int pos = (j & 1) ? (max_dim - 1 - (j >> 1)) : (j >> 1);
    bufd[rfft_in_slot(pos, log2m)] = v;
  }
  fft_dit<LM, false, 128, 3, TWL>(buf, log2m, tw);
  const double normalization = sqrt((double)max_dim);
  for (int d = tid; d < ndim; d += T) {
    const double2 X = rfft_bin<TWL>(buf, log2m, d, tw);
    const double2 w = weight[d];
    double v = (X.x * w.x - X.y * w.y) / normalization;
    if (d == 0) v += c0_add;
    out[f * ndim + d] = v;
  }
}

// dynamic shared memory: [ buf: cpad_size(max_dim) double2 | mel: max_dim + 2 doubles ]
__global__ void __launch_bounds__(128)
codec_decode_kernel(const double* __restrict__ coded, int half, int log2max, const int* __restrict__ idx,
                    const double* __restrict__ s, const double2* __restrict__ weight, int ndim,
                    const double2* __restrict__ tw, double* __restrict__ rows) {
  extern __shared__ double2 smem2[];
  const int max_dim = 1 << log2max;
  double2* buf = smem2;
  double* mel = reinterpret_cast<double*>(buf + cpad_size(max_dim));
  const int tid = threadIdx.x, T = 128;
  const size_t f = blockIdx.x;
  const double normalization = sqrt((double)max_dim);
  for (int i = tid; i < max_dim; i += T) {        // IDCTForCodec (:93-106)
    double2 z = make_double2(0.0, 0.0);
    if (i < ndim) {
      const double c = coded[f * ndim + i];
      z = make_double2(c * weight[i].x * normalization, -c * weight[i].y * normalization);
    }
    buf[cpad(brev(i, log2max))] = z;
  }
  // the reference's backward c2c wrapper returns conj(sum_j x_j e^{-2 pi i j k / n}) (W/src/fft.cpp
  // :36-46); only the real part is used, which equals the real part of the forward transform
  fft_dit<0, false, 128, 3>(buf, log2max, tw);
  for (int i = tid; i < max_dim / 2; i += T) {    // (:110-116)
    mel[1 + i * 2] = buf[cpad(i)].x;
    mel[1 + i * 2 + 1] = buf[cpad(max_dim - i - 1)].x;
  }
  __syncthreads();
  if (tid == 0) { mel[0] = mel[1]; mel[max_dim + 1] = mel[max_dim]; }
  __syncthreads();
  for (int k = tid; k <= half; k += T) {          // DecodeOneFrame (:150-154)
    const int q = idx[k];
    const double y0 = mel[q];
    const double v = add_rn(y0, mul_rn(s[k], add_rn(mel[q + 1], -y0)));
    rows[f * (half + 1) + k] = exp(v / max_dim);
  }
}

__global__ void lf0_kernel(const double* __restrict__ f0, int n, float* __restrict__ lf0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) lf0[i] = f0[i] != 0.0 ? static_cast<float>(log(f0[i])) : 0.f;     // ToLF0, analysis.cpp:216-224
}

__global__ void coded_to_float_kernel(const double* __restrict__ in, long long n, int ndim, int clamp_c0,
                                      float* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = in[i];
  if (clamp_c0 && (i % ndim) == 0 && v > 0 && v < 1e-4) v = 0;                  // analysis.cpp:349-351
  out[i] = static_cast<float>(v);
}

// per-dimension {count, sum, sum of squares} of a [frames][ndim] float matrix.  The matrix is read
// as ONE contiguous stream (coalesced, every sector used once): the CTA has R * ndim threads and
// the grid stride is a multiple of ndim, so a thread always meets the same column and keeps its
// partials in registers; the R threads of a column are joined in shared memory.  (One CTA column
// per dimension read 4 bytes of every 200-byte row: 8x the traffic, and a copy running beside it
// -- the waveform download of the end-to-end pipeline -- stretched it from 1.2 to 3.9 ms.)
__global__ void __launch_bounds__(1024)
feature_stats_kernel(const float* __restrict__ m, long long n_elems, int ndim, double* __restrict__ out3) {
  extern __shared__ double sh[];                    // [3][blockDim.x]
  const int T = blockDim.x, tid = threadIdx.x;      // T % ndim == 0
  double c = 0.0, s = 0.0, q = 0.0;
  for (long long i = blockIdx.x * (long long)T + tid; i < n_elems; i += (long long)gridDim.x * T) {
    const double x = m[i];
    c += 1.0; s += x; q += x * x;
  }
  sh[tid] = c; sh[T + tid] = s; sh[2 * T + tid] = q;
  __syncthreads();
  if (tid < ndim) {                                 // column tid: threads tid, tid + ndim, ...
    double v[3] = {0.0, 0.0, 0.0};
    for (int t = tid; t < T; t += ndim) { v[0] += sh[t]; v[1] += sh[T + t]; v[2] += sh[2 * T + t]; }
    atomicAdd(&out3[tid * 3], v[0]); atomicAdd(&out3[tid * 3 + 1], v[1]); atomicAdd(&out3[tid * 3 + 2], v[2]);
  }
}


// ---- global-variance statistics (scripts/Training.pl make_data_gv :1402-1456, data/Makefile.in:447-458) ---
// For every utterance and every static stream the reference runs `vstat -d -o 2` over the frames of the
// utterance: per-dimension sum and sum of squares in double, in frame order, variance = E[x^2] - mean^2
// over the k frames it read; lf0 first loses its unvoiced frames (the "1e+10" filter of :1437).  One CTA per
// utterance; thread c owns column c of [mgc | lf0 | bap] and walks the frames in order, i.e. the additions
// happen in vstat's order (neighbouring threads read neighbouring floats of a row: coalesced).
// out[u][c] = variance (NaN when the stream has no frame: vstat prints nothing, the NaN check of
// Training.pl:1451 then drops the utterance).
__global__ void __launch_bounds__(256)
gv_utterance_kernel(const float* __restrict__ mgc, const float* __restrict__ lf0, const float* __restrict__ bap,
                    int mgc_dim, int bap_dim, const int* __restrict__ f_off, const int* __restrict__ f_len,
                    double* __restrict__ out) {
  const int u = blockIdx.x, ncol = mgc_dim + 1 + bap_dim;
  const int f0 = f_off[u], n = f_len[u];
  for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
    const float* __restrict__ base;
    int stride;
    bool msd = false;
    if (c < mgc_dim) { base = mgc + (size_t)f0 * mgc_dim + c; stride = mgc_dim; }
    else if (c == mgc_dim) { base = lf0 + f0; stride = 1; msd = true; }
    else { base = bap + (size_t)f0 * bap_dim + (c - mgc_dim - 1); stride = bap_dim; }
    double sum = 0.0, sq = 0.0;
    long long k = 0;
    for (int i = 0; i < n; ++i) {
      const float v = base[(size_t)i * stride];
      if (msd && v == 0.0f) continue;                 // unvoiced: lf0 = 0 here, -1e10 in the reference's files
      const double x = v;
      sum = add_rn(sum, x);
      sq = add_rn(sq, mul_rn(x, x));
      ++k;
    }
    double var = __longlong_as_double(0x7ff8000000000000LL);
    if (k > 0) { const double mean = sum / k; var = add_rn(sq / k, -mul_rn(mean, mean)); }
    out[(size_t)u * ncol + c] = var;
  }
}


// ---- the synth tool's decode of CODED aperiodicity (W/test/synth.cpp:221-247) ------------------------------
// c[0] += 9.210340;  mgc2sp(c, m, 0.55, 0, x, y, fft_size)  =  freqt(c, m, c2, fft_size / 2, -0.55) followed by
// x = Re FFT_fft_size(c2) (W/test/sptkfunctions.cpp:186-275; the gnorm / gc2gc / ignorm steps are the identity
// for gamma = 0);  ap[j] = exp(x[j]) / 1e4.  The whole chain before the exponential is linear in c, so it is one
// (fft_size / 2 + 1) x (m + 1) matrix B, built once on the host by running the reference's recursion on the unit
// vectors and transforming the result: x = B c, 25 multiply-adds per bin instead of a 25 x 1024-step
// sequential recursion and a transform per frame.  (Summation order differs from the recursion: 1e-15.)
// The reference fills only the first m bins of a row (the rest is uninitialised) and, for an even
// ap_dimension, reads one coefficient past the ones it loaded; here every bin gets exp(x[j]) / 1e4 and the
// missing coefficient is 0.
struct BapTables { int fft_size = 0, m = 0; DevBuf<double> B; };      // B[j][i], j <= fft_size / 2, i <= m
std::map<std::pair<int, int>, BapTables*> g_bap;

void host_freqt(const std::vector<double>& c1, int m2, double a, std::vector<double>* out) {   // sptkfunctions.cpp freqt
  const int m1 = (int)c1.size() - 1;
  const double b = 1 - a * a;
  std::vector<double> g(m2 + 1, 0.0), d(m2 + 1, 0.0);
  for (int i = -m1; i <= 0; ++i) {
    g[0] = c1[-i] + a * (d[0] = g[0]);
    if (1 <= m2) g[1] = b * d[0] + a * (d[1] = g[1]);
    for (int j = 2; j <= m2; ++j) g[j] = d[j - 1] + a * ((d[j] = g[j]) - g[j - 1]);
  }
  *out = g;
}

BapTables* bap_tables(int fft_size, int m) {
  auto it = g_bap.find({fft_size, m});
  if (it != g_bap.end()) return it->second;
  if (fft_size < 32 || (fft_size & (fft_size - 1)) || m < 1 || m > 255) { set_error("bap decode: fft_size %d / order %d", fft_size, m); return nullptr; }
  const int H = fft_size / 2 + 1;
  std::vector<double> B((size_t)H * (m + 1));
  for (int i = 0; i <= m; ++i) {
    std::vector<double> e(m + 1, 0.0), c2;
    e[i] = 1.0;
    host_freqt(e, fft_size / 2, -0.55, &c2);
    std::vector<double> re(fft_size, 0.0), im(fft_size, 0.0);
    for (int k = 0; k <= fft_size / 2; ++k) re[k] = c2[k];
    host_fft(re, im);
    for (int j = 0; j < H; ++j) B[(size_t)j * (m + 1) + i] = re[j];
  }
  BapTables* t = new BapTables();
  t->fft_size = fft_size; t->m = m;
  if (!t->B.alloc(B.size()) || !WB_CUDA(cudaMemcpy(t->B.p, B.data(), B.size() * sizeof(double), cudaMemcpyHostToDevice))) { delete t; return nullptr; }
  g_bap[{fft_size, m}] = t;
  return t;
}

constexpr int kBapFrames = 8;            // frames per CTA: a thread keeps its bin's row of B in registers
template <int MAXM>
__global__ void __launch_bounds__(256)
bap_decode_kernel(const float* __restrict__ bap, int bap_dim, int m, int n_coef, const double* __restrict__ B, int H,
                  int n_frames, double* __restrict__ ap_rows) {
  __shared__ double c[kBapFrames][MAXM + 1];
  const int f0 = blockIdx.x * kBapFrames;
  for (int t = threadIdx.x; t < kBapFrames * (m + 1); t += blockDim.x) {
    const int fr = t / (m + 1), i = t - fr * (m + 1);
    double v = 0.0;
    if (f0 + fr < n_frames && i < n_coef) v = static_cast<double>(bap[(size_t)(f0 + fr) * bap_dim + i]);   // ToDouble
    if (i == 0) v += 9.210340;                                                                                 // :241
    c[fr][i] = v;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    double row[MAXM + 1];
#pragma unroll
    for (int i = 0; i <= MAXM; ++i) row[i] = i <= m ? B[(size_t)j * (m + 1) + i] : 0.0;
    for (int fr = 0; fr < kBapFrames && f0 + fr < n_frames; ++fr) {
      double x = 0.0;
#pragma unroll
      for (int i = 0; i <= MAXM; ++i) x += row[i] * c[fr][i];
      ap_rows[(size_t)(f0 + fr) * H + j] = exp(x) / 1e4;                                                     // :244
    }
  }
}

__global__ void lf0_to_f0_kernel(const float* __restrict__ lf0, int n, double* __restrict__ f0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const double v = lf0[i]; f0[i] = v != 0.0 ? exp(v) : 0.0; }          // ToF0, W/test/synth.cpp:81-89
}

}  // namespace

bool codec_encode_run(const double* d_rows, int n_frames, int fs, int fft_size, int ndim, double scale,
                      double zero_floor, double c0_add, double* d_out, bool f32log) {
  Context* c = ctx();
  if (!c) return false;
  if (n_frames <= 0) return true;
  CodecTables* t = codec_tables(fs, fft_size, ndim);
  if (!t) return false;
  const size_t smem = cpad_size(t->max_dim / 2) * sizeof(double2) + (t->max_dim + 2) * sizeof(double);
  KernelTimer kt("codec_encode_kernel");
#define WB_ENC_LAUNCH(L, F)                                                                                         \
  do {                                                                                                              \
    WB_CUDA_OR_RETURN(cudaFuncSetAttribute(codec_encode_kernel<L, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), false); \
    codec_encode_kernel<L, F><<<n_frames, 128, smem, c->stream>>>(d_rows, fft_size / 2, t->log2max, t->enc_idx.p, t->enc_s.p, t->enc_w.p, ndim, \
                                                              scale, zero_floor, c0_add, L > 0 ? c->tw_c(L > 0 ? L : 4) : c->d_twiddle, d_out);     \
  } while (0)
  switch (t->log2max) {
    case 9: if (f32log) WB_ENC_LAUNCH(9, true); else WB_ENC_LAUNCH(9, false); break;
    case 10: if (f32log) WB_ENC_LAUNCH(10, true); else WB_ENC_LAUNCH(10, false); break;
    default: WB_ENC_LAUNCH(0, false); break;
  }
#undef WB_ENC_LAUNCH
  WB_LAUNCH_CHECK(); kt.stop();
  return true;
}

bool codec_decode_run(const double* d_coded, int n_frames, int fs, int fft_size, int ndim, double* d_rows) {
  Context* c = ctx();
  if (!c) return false;
  if (n_frames <= 0) return true;
  CodecTables* t = codec_tables(fs, fft_size, ndim);
  if (!t) return false;
  const size_t smem = cpad_size(t->max_dim) * sizeof(double2) + (t->max_dim + 2) * sizeof(double);
  WB_CUDA_OR_RETURN(cudaFuncSetAttribute(codec_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), false);
  KernelTimer kt("codec_decode_kernel");
  codec_decode_kernel<<<n_frames, 128, smem, c->stream>>>(d_coded, fft_size / 2, t->log2max, t->dec_idx.p, t->dec_s.p, t->dec_w.p, ndim,
                                                         c->d_twiddle, d_rows);
  WB_LAUNCH_CHECK(); kt.stop();
  return true;
}

// the analysis tool's coded outputs (W/test/analysis.cpp:293-390) for a whole batch
bool batch_code_features(Batch* b, int mgc_dim, int bap_dim) {
  Context* c = ctx();
  if (!c) return false;
  if (!b->sp.p || !b->ap.p || b->fft_size <= 0) { set_error("code: CheapTrick and D4C have not been run"); return false; }
  const int F = b->total_frames;
  DevBuf<double> tmp;
  if (!tmp.alloc((size_t)F * std::max(mgc_dim, bap_dim)) || !b->lf0.alloc(F) || !b->mgc.alloc((size_t)F * mgc_dim) ||
      !b->bap.alloc((size_t)F * bap_dim))
    return false;
  b->mgc_dim = mgc_dim; b->bap_dim = bap_dim;
  if (F == 0) return true;
  cudaStream_t st = c->stream;
  lf0_kernel<<<(F + 255) / 256, 256, 0, st>>>(b->f0.p, F, b->lf0.p);
  WB_LAUNCH_CHECK();
  if (!codec_encode_run(b->sp.p, F, b->fs, b->fft_size, mgc_dim, 1e4, 0.0001, 12.0, tmp.p, true)) return false;
  long long n = (long long)F * mgc_dim;
  coded_to_float_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tmp.p, n, mgc_dim, 0, b->mgc.p);
  WB_LAUNCH_CHECK();
  if (!codec_encode_run(b->ap.p, F, b->fs, b->fft_size, bap_dim, 1e4, 0.0, -9.210340, tmp.p, true)) return false;
  n = (long long)F * bap_dim;
  coded_to_float_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tmp.p, n, bap_dim, 1, b->bap.p);
  WB_LAUNCH_CHECK();
  WB_CUDA_OR_RETURN(cudaStreamSynchronize(st), false);
  return true;
}

bool batch_feature_stats(Batch* b, double* h_out) {   // [(1 + mgc_dim)][3]: lf0 (voiced frames), then mgc
  Context* c = ctx();
  if (!c) return false;
  if (!b->mgc.p) { set_error("stats: features have not been coded"); return false; }
  const int F = b->total_frames, nd = b->mgc_dim;
  DevBuf<double> d;
  if (!d.alloc((size_t)(nd + 1) * 3)) return false;
  cudaStream_t st = c->stream;
  if (!dev_fill(d.p, 0, (size_t)(nd + 1) * 3 * sizeof(double))) return false;
  if (F > 0) {
    if (nd > 1024) { set_error("stats: %d dimensions (<= 1024 supported)", nd); return false; }
    const int threads = nd * std::max(1, 256 / nd);
    feature_stats_kernel<<<c->sm_count * 4, threads, 3 * threads * sizeof(double), st>>>(b->mgc.p, (long long)F * nd, nd, d.p + 3);
    WB_LAUNCH_CHECK();
  }
  if (!read_back(h_out, d.p, (size_t)(nd + 1) * 3 * sizeof(double))) return false;
  return true;
}

// per-utterance variances [n_utt][mgc_dim + 1 + bap_dim] (host, doubles) and, per column, the partials
// {count, sum, sum of squares} over the utterances of this batch of the variances ROUNDED TO FLOAT32 (the
// reference's tmp.var1 is a float file; data/Makefile.in:451-453 runs vstat over it): the all-reduce of
// these rows gives stats/gv.var on any number of GPUs.  Utterances whose variance is NaN are not counted.
bool batch_gv_stats(Batch* b, double* h_per_utt, double* h_partials) {
  Context* c = ctx();
  if (!c) return false;
  if (!b->mgc.p || !b->bap.p || !b->lf0.p) { set_error("gv: features have not been coded (wb200_batch_code)"); return false; }
  const int ncol = b->mgc_dim + 1 + b->bap_dim;
  std::vector<double> local((size_t)std::max(1, b->n_utt) * ncol);
  if (b->n_utt > 0) {
    DevBuf<double> d;
    if (!d.alloc((size_t)b->n_utt * ncol)) return false;
    gv_utterance_kernel<<<b->n_utt, 256, 0, c->stream>>>(b->mgc.p, b->lf0.p, b->bap.p, b->mgc_dim, b->bap_dim, b->f_off.p, b->f_len.p, d.p);
    WB_LAUNCH_CHECK();
    if (!WB_CUDA(cudaMemcpyAsync(local.data(), d.p, (size_t)b->n_utt * ncol * sizeof(double), cudaMemcpyDeviceToHost, c->stream)) ||
        !WB_CUDA(cudaStreamSynchronize(c->stream)))
      return false;
  }
  if (h_per_utt) memcpy(h_per_utt, local.data(), (size_t)b->n_utt * ncol * sizeof(double));
  if (h_partials) {
    for (int k = 0; k < ncol; ++k) {
      double cnt = 0.0, sum = 0.0, sq = 0.0;
      for (int u = 0; u < b->n_utt; ++u) {              // utterance order = the order of the reference's loop over files
        const double v = local[(size_t)u * ncol + k];
        if (v != v) continue;
        const double x = static_cast<double>(static_cast<float>(v));
        cnt += 1.0; sum += x; sq += x * x;
      }
      h_partials[3 * k] = cnt; h_partials[3 * k + 1] = sum; h_partials[3 * k + 2] = sq;
    }
  }
  return true;
}

// coded aperiodicity [n_frames][bap_dim] float32 (device) -> aperiodicity rows [n_frames][fft_size/2+1]
bool bap_decode_run(const float* d_bap, int n_frames, int fft_size, int bap_dim, double* d_rows) {
  Context* c = ctx();
  if (!c) return false;
  if (n_frames <= 0) return true;
  const int m = (bap_dim % 2 == 1) ? bap_dim - 1 : bap_dim;          // W/test/synth.cpp:221-223, 242
  BapTables* t = bap_tables(fft_size, m);
  if (!t) return false;
  const int H = fft_size / 2 + 1, n_coef = std::min(bap_dim, m + 1);
  const int grid = (n_frames + kBapFrames - 1) / kBapFrames;
  KernelTimer kt("bap_decode_kernel");
  if (m <= 32) bap_decode_kernel<32><<<grid, 256, 0, c->stream>>>(d_bap, bap_dim, m, n_coef, t->B.p, H, n_frames, d_rows);
  else if (m <= 64) bap_decode_kernel<64><<<grid, 256, 0, c->stream>>>(d_bap, bap_dim, m, n_coef, t->B.p, H, n_frames, d_rows);
  else { set_error("bap decode: order %d (<= 64 supported)", m); return false; }
  WB_LAUNCH_CHECK(); kt.stop();
  return true;
}
bool lf0_to_f0_run(const float* d_lf0, int n, double* d_f0) {
  Context* c = ctx();
  if (!c) return false;
  if (n <= 0) return true;
  lf0_to_f0_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d_lf0, n, d_f0);
  WB_LAUNCH_CHECK();
  return true;
}

}  // namespace wb
