// world-b200: Dio F0 estimation over a batch of utterances.
//
// Reference: W/src/dio.cpp — Dio :642-647, DioGeneralBody :578-634, GetSpectrumForEstimation
// :60-106, DesignLowCutFilter :40-53, GetFilteredSignal :296-343, ZeroCrossingEngine :357-393,
// GetFourZeroCrossingIntervals :402-435, GetF0CandidateContour(+Sub) :441-508,
// GetF0CandidatesAndScores :549-572, GetBestF0Contour :112-126, FixF0Contour :259-289
// (FixStep1-4 :132-253, SelectBestF0 :190-209); W/src/matlabfunctions.cpp interp1 :157-182.
//
// The reference filters the whole utterance with FFTs of 2^17..2^19 points (one forward, then
// per band one forward of the filter and one inverse: 16 big transforms per utterance).  The
// filters are short (low-cut (+-fs/50) convolved with a Nuttall low-pass of <= 4*fs/(2*71 Hz)
// taps), so here the same circular convolution is evaluated by overlap-save with FFT blocks
// that live entirely in shared memory: one forward transform of an 8192-sample block, then
// for each of the 7 bands a multiply by the precomputed combined filter spectrum and one
// inverse transform.  Indexing the input modulo the reference's FFT size reproduces its
// wrap-around at the utterance edges.  Zero crossings are then
// compacted per (utterance, band, event type) with ballot scans and interpolated onto the
// frame times by binary search.
#include <math.h>
#include <algorithm>
#include <map>
#include <vector>
#include "wb_batch.h"
#include "wb_fft.cuh"

namespace wb {
namespace {

constexpr int kMaxDioBands = 16;
constexpr int kZcChunk = 2048;       // sample pairs per CTA in the zero-crossing kernels

struct DioFilterBank {
  int nb = 0;
  int bn = 0, log2bn = 0;            // overlap-save block size (real points)
  int hN = 0;                        // low-cut half length
  int hal[kMaxDioBands];             // Nuttall half_average_length per band
  int D = 0, V = 0;                  // block pre-roll and valid outputs per block
  DevBuf<double2> G;                 // [nb][bn/2 + 1] spectra of the causal combined filters
  double boundary_f0[kMaxDioBands];
};

// ---- host: build the combined filters and their spectra ------------------------------------------
void host_fft(std::vector<double>& re, std::vector<double>& im) {   // in-place radix-2, forward
  const size_t n = re.size();
  for (size_t i = 1, j = 0; i < n; ++i) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    const long double ang = -2.0L * 3.14159265358979323846264338327950288L / len;
    for (size_t i = 0; i < n; i += len)
      for (size_t k = 0; k < len / 2; ++k) {
        const double wr = (double)cosl(ang * k), wi = (double)sinl(ang * k);
        const size_t a = i + k, b = i + k + len / 2;
        const double xr = re[b] * wr - im[b] * wi, xi = re[b] * wi + im[b] * wr;
        re[b] = re[a] - xr; im[b] = im[a] - xi;
        re[a] += xr; im[a] += xi;
      }
  }
}

bool build_filter_bank(double actual_fs, const DioParams& p, DioFilterBank* fb) {
  fb->nb = 1 + static_cast<int>(log(p.f0_ceil / p.f0_floor) / kLog2 * p.channels_in_octave);   // :582-583
  if (fb->nb < 1 || fb->nb > kMaxDioBands) { set_error("Dio: unsupported number of bands %d", fb->nb); return false; }
  for (int i = 0; i < fb->nb; ++i) {
    fb->boundary_f0[i] = p.f0_floor * pow(2.0, (i + 1) / p.channels_in_octave);                // :585-586
    fb->hal[i] = matlab_round(actual_fs / fb->boundary_f0[i] / 2.0);                           // :532
  }
  // low-cut filter (:40-53, :86-87): zero-phase kernel lc[k], k in [-hN, hN]
  const int cutoff = matlab_round(actual_fs / 50.0);
  const int N = cutoff * 2 + 1;
  fb->hN = (N - 1) / 2;
  std::vector<double> lc(N);
  double sum = 0.0;
  for (int i = 1; i <= N; ++i) lc[i - 1] = 0.5 - 0.5 * cos(i * 2.0 * kPi / (N + 1));
  for (int i = 0; i < N; ++i) sum += lc[i];
  for (int i = 0; i < N; ++i) lc[i] = -lc[i] / sum;
  lc[fb->hN] += 1.0;                                   // the "+1" lands on k = 0
  int lmax = 0;
  for (int b = 0; b < fb->nb; ++b) lmax = std::max(lmax, 4 * fb->hal[b] + 2 * fb->hN);
  int bn = 1024;
  while (bn < 3 * lmax && bn < 8192) bn <<= 1;
  if (bn < lmax + 64) { set_error("Dio: filters of %d taps do not fit the block FFT (fs too high)", lmax); return false; }
  fb->bn = bn;
  fb->log2bn = 0; while ((1 << fb->log2bn) < bn) ++fb->log2bn;
  fb->D = 2 * fb->hal[0] + fb->hN - 1;
  fb->V = bn - (4 * fb->hal[0] + 2 * fb->hN - 1);
  std::vector<double2> G((size_t)fb->nb * (bn / 2 + 1));
  for (int b = 0; b < fb->nb; ++b) {
    const int ln = 4 * fb->hal[b];
    std::vector<double> nut(ln);
    for (int i = 0; i < ln; ++i) {                      // NuttallWindow, common.cpp:113-121
      const double tmp = i / (ln - 1.0);
      nut[i] = 0.355768 - 0.487396 * cos(2.0 * kPi * tmp) + 0.144232 * cos(4.0 * kPi * tmp) -
               0.012604 * cos(6.0 * kPi * tmp);
    }
    // causal combined filter g'[k'] = sum_j lc[j] * nut[k' - j], k' in [0, ln + N - 1)
    std::vector<double> re(bn, 0.0), im(bn, 0.0);
    for (int j = 0; j < N; ++j) {
      const double c = lc[j];
      for (int i = 0; i < ln; ++i) re[i + j] += c * nut[i];
    }
    host_fft(re, im);
    for (int k = 0; k <= bn / 2; ++k) G[(size_t)b * (bn / 2 + 1) + k] = make_double2(re[k], im[k]);
  }
  if (!fb->G.alloc(G.size())) return false;
  return WB_CUDA(cudaMemcpy(fb->G.p, G.data(), G.size() * sizeof(double2), cudaMemcpyHostToDevice));
}

struct DioConst {
  int nb, bn, log2bn, hN, D, V;
  int hal[kMaxDioBands];
  double boundary_f0[kMaxDioBands];
  double actual_fs, f0_floor, f0_ceil, allowed_range;
  int voice_range_minimum;
};

// ---- mean of y over y_length (= x_length + 1, the extra sample is 0) (:70-73) ----------------------
__global__ void dio_mean_kernel(const double* __restrict__ x_all, const long long* __restrict__ x_off,
                                const int* __restrict__ x_len, const int* __restrict__ y_len,
                                double* __restrict__ mean) {
  __shared__ double red[96];
  const int u = blockIdx.x;
  const double* __restrict__ x = x_all + x_off[u];
  const int n = x_len[u];
  double v[1] = {0.0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) v[0] += x[i];
  block_sum<1>(v, red);
  if (threadIdx.x == 0) mean[u] = v[0] / y_len[u];
}

// ---- overlap-save filtering: one CTA per (block, utterance) ---------------------------------------
// dynamic shared memory: [ xs: cpad_size(bn/2) double2 | ws: cpad_size(bn/2) double2 ]
template <int LOG2BN>     // 0: block size given at run time (c.log2bn)
__global__ void __launch_bounds__(256)
dio_filter_kernel(const double* __restrict__ x_all, const long long* __restrict__ x_off,
                  const int* __restrict__ x_len, const int* __restrict__ y_len_all,
                  const int* __restrict__ fft_mask_all, const double* __restrict__ mean_all,
                  const long long* __restrict__ F_off, const double2* __restrict__ G,
                  const double2* __restrict__ tw, DioConst c, int utt0, double* __restrict__ F) {
  extern __shared__ double2 smem2[];
  const int u = utt0 + blockIdx.y;
  const int y_len = y_len_all[u];
  const int n0 = blockIdx.x * c.V;
  if (n0 >= y_len) return;
  constexpr int LM = LOG2BN > 0 ? LOG2BN - 1 : 0;
  const int log2m = LOG2BN > 0 ? LOG2BN - 1 : c.log2bn - 1, M = 1 << log2m;
  double2* xs = smem2;
  double2* ws = smem2 + cpad_size(M);
  double* xsd = reinterpret_cast<double*>(xs);
  double* wsd = reinterpret_cast<double*>(ws);
  const int tid = threadIdx.x, T = blockDim.x;
  const double* __restrict__ x = x_all + x_off[u];
  const int xl = x_len[u];
  const int mask = fft_mask_all[u];
  const double mean = mean_all[u];
  // input block: ypad[(n0 - D + i) mod FS]
  for (int i = tid; i < c.bn; i += T) {
    const int m = (n0 - c.D + i) & mask;
    double v = 0.0;
    if (m < y_len) v = (m < xl ? x[m] : 0.0) - mean;
    xsd[rfft_in_slot(i, log2m)] = v;
  }
  fft_dit<LM, false, 256>(xs, log2m, tw);
  // half spectrum in place: slot k = X[k] (k < M), slot 0 = (X[0], X[M])
  for (int k = tid; k <= M / 2; k += T) {
    if (k == 0) {
      const double2 z0 = xs[0];
      xs[0] = make_double2(z0.x + z0.y, z0.x - z0.y);
    } else {
      const double2 a = rfft_bin(xs, log2m, k, tw);
      const double2 b = rfft_bin(xs, log2m, M - k, tw);
      xs[cpad(k)] = a;
      xs[cpad(M - k)] = b;
    }
  }
  __syncthreads();
  const int n_out = min(c.V, y_len - n0);
  double* __restrict__ Fu = F + F_off[u - utt0];
  for (int b = 0; b < c.nb; ++b) {
    const double2* __restrict__ Gb = G + (size_t)b * (M + 1);
    for (int k = tid; k <= M / 2; k += T) {
      if (k == 0) {
        const double2 x0 = xs[0];
        const double2 y0 = make_double2(x0.x * Gb[0].x, 0.0), yM = make_double2(x0.y * Gb[M].x, 0.0);
        ws[cpad(brev(0, log2m))] = c2r_pack(y0, yM, 0, log2m, tw);
      } else {
        const double2 yk = cmul(xs[cpad(k)], Gb[k]);
        const double2 ym = cmul(xs[cpad(M - k)], Gb[M - k]);
        ws[cpad(brev(k, log2m))] = c2r_pack(yk, ym, k, log2m, tw);
        if (k != M - k) ws[cpad(brev(M - k, log2m))] = c2r_pack(ym, yk, M - k, log2m, tw);
      }
    }
    fft_dit<LM, true, 256>(ws, log2m, tw);
    const int shift = c.D + 2 * c.hal[b] + c.hN;       // filtered_b[n] = conv[n - n0 + shift]
    double* __restrict__ dst = Fu + (size_t)b * y_len + n0;
    for (int i = tid; i < n_out; i += T) dst[i] = wsd[rfft_out_slot(i + shift)];
    __syncthreads();
  }
}

// ---- zero crossings ------------------------------------------------------------------------------
// Event types on the filtered signal s (GetFourZeroCrossingIntervals :402-435):
//   0 negative-going  s[i] > 0 >= s[i+1]            1 positive-going  -s
//   2 peaks           d[i] = s[i+1] - s[i], d[i] > 0 >= d[i+1]      3 dips  -d
__device__ __forceinline__ bool zc_event(double a, double b) { return 0.0 < a && b <= 0.0; }
__device__ __forceinline__ double zc_fine(int e, double a, double b) {   // :376-379
  return add_rn((double)e, -div_rn(a, add_rn(b, -a)));
}

template <bool WRITE>
__global__ void __launch_bounds__(256)
dio_zc_kernel(const double* __restrict__ F, const long long* __restrict__ F_off,
              const int* __restrict__ y_len_all, int nb, int utt0, int n_chunks_max,
              int* __restrict__ counts,              // [lists][n_chunks_max] (count pass: out; write: exclusive offsets)
              const long long* __restrict__ list_off, double* __restrict__ edges) {
  __shared__ int wcnt[4][8];
  __shared__ int run[4];
  const int ub = blockIdx.y;                         // local utterance * nb + band
  const int u_local = ub / nb, b = ub % nb;
  const int y_len = y_len_all[utt0 + u_local];
  const int chunk = blockIdx.x;
  const int i0 = chunk * kZcChunk;
  if (i0 >= y_len - 1) {
    if (!WRITE && threadIdx.x < 4) counts[((size_t)ub * 4 + threadIdx.x) * n_chunks_max + chunk] = 0;
    return;
  }
  const double* __restrict__ s = F + F_off[u_local] + (size_t)b * y_len;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid < 4) run[tid] = WRITE ? counts[((size_t)ub * 4 + tid) * n_chunks_max + chunk] : 0;
  __syncthreads();
  for (int it = 0; it < kZcChunk / 256; ++it) {
    const int i = i0 + it * 256 + tid;
    bool ev[4] = {false, false, false, false};
    double fine[4] = {0.0, 0.0, 0.0, 0.0};
    if (i < y_len - 1) {
      const double s0 = s[i], s1 = s[i + 1];
      ev[0] = zc_event(s0, s1);
      ev[1] = zc_event(-s0, -s1);
      if (WRITE && ev[0]) fine[0] = zc_fine(i + 1, s0, s1);
      if (WRITE && ev[1]) fine[1] = zc_fine(i + 1, -s0, -s1);
      if (i < y_len - 2) {
        const double s2 = s[i + 2];
        const double d0 = add_rn(s1, -s0), d1 = add_rn(s2, -s1);
        ev[2] = zc_event(d0, d1);
        ev[3] = zc_event(-d0, -d1);
        if (WRITE && ev[2]) fine[2] = zc_fine(i + 1, d0, d1);
        if (WRITE && ev[3]) fine[3] = zc_fine(i + 1, -d0, -d1);
      }
    }
    unsigned bal[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      bal[t] = __ballot_sync(0xffffffffu, ev[t]);
      if (lane == 0) wcnt[t][wid] = __popc(bal[t]);
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (WRITE && ev[t]) {
        int pos = run[t];
        for (int w = 0; w < wid; ++w) pos += wcnt[t][w];
        pos += __popc(bal[t] & ((1u << lane) - 1u));
        edges[list_off[(size_t)ub * 4 + t] + pos] = fine[t];
      }
    }
    __syncthreads();
    if (tid < 4) {
      int tot = 0;
      for (int w = 0; w < 8; ++w) tot += wcnt[tid][w];
      run[tid] += tot;
    }
    __syncthreads();
  }
  if (!WRITE && tid < 4) counts[((size_t)ub * 4 + tid) * n_chunks_max + chunk] = run[tid];
}

// one thread per list: exclusive scan over chunks (in place), list totals out
__global__ void dio_zc_scan_kernel(int* __restrict__ counts, int n_lists, int n_chunks_max,
                                   int* __restrict__ totals) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_lists) return;
  int* c = counts + (size_t)l * n_chunks_max;
  int acc = 0;
  for (int i = 0; i < n_chunks_max; ++i) { const int v = c[i]; c[i] = acc; acc += v; }
  totals[l] = acc;
}

// interp1 (matlabfunctions.cpp:157-182) of interval-F0 at time t; knots are the mid-points of
// consecutive edges: loc_i = (e_i + e_{i+1}) / 2 / fs, val_i = fs / (e_{i+1} - e_i), i < n_int.
__device__ __forceinline__ double zc_loc(const double* __restrict__ e, int i, double fs) {
  return div_rn(div_rn(add_rn(e[i], e[i + 1]), 2.0), fs);
}
__device__ __forceinline__ double zc_val(const double* __restrict__ e, int i, double fs) {
  return div_rn(fs, add_rn(e[i + 1], -e[i]));
}
__device__ double zc_interp(const double* __restrict__ e, int n_int, double fs, double t) {
  int lo = 0, hi = n_int;                 // upper_bound: first knot with loc > t
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (zc_loc(e, mid, fs) <= t) lo = mid + 1; else hi = mid;
  }
  const int k = max(1, min(n_int - 1, lo));
  const double x0 = zc_loc(e, k - 1, fs), x1 = zc_loc(e, k, fs);
  const double y0 = zc_val(e, k - 1, fs), y1 = zc_val(e, k, fs);
  const double sfrac = div_rn(add_rn(t, -x0), add_rn(x1, -x0));
  return add_rn(y0, mul_rn(sfrac, add_rn(y1, -y0)));
}

// candidates and scores per (utterance, band, frame)  (:441-508, :549-572)
__global__ void dio_candidates_kernel(const double* __restrict__ edges, const long long* __restrict__ list_off,
                                      const int* __restrict__ list_cnt, const int* __restrict__ f_off,
                                      const int* __restrict__ f_len, const double* __restrict__ frame_t,
                                      DioConst c, int utt0, int total_frames,
                                      double* __restrict__ cand, double* __restrict__ score) {
  const int u_local = blockIdx.y / c.nb, b = blockIdx.y % c.nb;
  const int u = utt0 + u_local;
  const int n_fr = f_len[u];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_fr) return;
  const int fidx = f_off[u] + i;
  const size_t l0 = ((size_t)u_local * c.nb + b) * 4;
  int n_int[4];
  bool ok = true;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int cnt = list_cnt[l0 + t];
    n_int[t] = cnt < 2 ? 0 : cnt - 1;          // ZeroCrossingEngine's return value (:369-373, :392)
    ok = ok && (n_int[t] - 2 > 0);             // CheckEvent (:484-487)
  }
  double cd = 0.0, sc = kMaximumValue;
  if (ok) {
    const double t = frame_t[fidx];
    double v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = zc_interp(edges + list_off[l0 + q], n_int[q], c.actual_fs, t);
    cd = div_rn(add_rn(add_rn(add_rn(v[0], v[1]), v[2]), v[3]), 4.0);
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) { const double d = add_rn(v[q], -cd); acc = add_rn(acc, mul_rn(d, d)); }
    sc = sqrt(div_rn(acc, 3.0));
    const double bf = c.boundary_f0[b];
    if (cd > bf || cd < bf / 2.0 || cd > c.f0_ceil || cd < c.f0_floor) { cd = 0.0; sc = kMaximumValue; }
  }
  cand[(size_t)b * total_frames + fidx] = cd;
  score[(size_t)b * total_frames + fidx] = div_rn(sc, add_rn(cd, kMySafeGuardMinimum));   // :566-567
}

// SelectBestF0 (:190-209)
__device__ double dio_select_best(double current_f0, double past_f0, const double* __restrict__ cand,
                                  int total_frames, int nb, int target, double allowed_range) {
  const double reference_f0 = div_rn(add_rn(mul_rn(current_f0, 3.0), -past_f0), 2.0);
  double best = cand[target];
  double minimum_error = fabs(add_rn(reference_f0, -best));
  for (int i = 1; i < nb; ++i) {
    const double cv = cand[(size_t)i * total_frames + target];
    const double e = fabs(add_rn(reference_f0, -cv));
    if (e < minimum_error) { minimum_error = e; best = cv; }
  }
  if (fabs(add_rn(1.0, -div_rn(best, reference_f0))) > allowed_range) return 0.0;
  return best;
}

// best contour + FixF0Contour, one CTA per utterance (:112-126, :259-289)
__global__ void __launch_bounds__(256)
dio_fix_kernel(const double* __restrict__ cand, const double* __restrict__ score,
               const int* __restrict__ f_off, const int* __restrict__ f_len, DioConst c, int utt0,
               int total_frames, double* __restrict__ tmp1, double* __restrict__ tmp2,
               int* __restrict__ pos_idx, int* __restrict__ neg_idx, double* __restrict__ f0_out) {
  const int u = utt0 + blockIdx.x;
  const int off = f_off[u], F = f_len[u];
  const int tid = threadIdx.x, T = blockDim.x;
  const int vrm = c.voice_range_minimum;
  if (F <= vrm) return;                               // :265 (f0 is left untouched)
  double* best = tmp2 + off;       // best contour, later step2
  double* s1 = tmp1 + off;         // step1, later step3
  double* out = f0_out + off;
  const double* cd = cand + off;
  const double* sc = score + off;
  // GetBestF0Contour: first minimum wins
  for (int i = tid; i < F; i += T) {
    double m = sc[i], bv = cd[i];
    for (int j = 1; j < c.nb; ++j) {
      const double sj = sc[(size_t)j * total_frames + i];
      if (m > sj) { m = sj; bv = cd[(size_t)j * total_frames + i]; }
    }
    out[i] = bv;                   // park the best contour in the output buffer
  }
  __syncthreads();
  // FixStep1 (:132-150)
  for (int i = tid; i < F; i += T) {
    double v = 0.0;
    if (i >= vrm) {
      const double bi = (i < F - vrm) ? out[i] : 0.0;
      const double bp = (i - 1 >= vrm && i - 1 < F - vrm) ? out[i - 1] : 0.0;
      v = fabs(div_rn(add_rn(bi, -bp), add_rn(kMySafeGuardMinimum, bi))) < c.allowed_range ? bi : 0.0;
    }
    s1[i] = v;
  }
  __syncthreads();
  // FixStep2 (:156-169)
  const int center = (vrm - 1) / 2;
  for (int i = tid; i < F; i += T) {
    double v = s1[i];
    if (i >= center && i < F - center)
      for (int j = -center; j <= center; ++j)
        if (s1[i + j] == 0) { v = 0.0; break; }
    best[i] = v;                   // step2
  }
  __syncthreads();
  if (tid == 0) {
    const double* s2 = best;
    int pc = 0, nc = 0;            // GetNumberOfVoicedSections (:174-184)
    int* pidx = pos_idx + off;
    int* nidx = neg_idx + off;
    for (int i = 1; i < F; ++i) {
      if (s2[i] == 0 && s2[i - 1] != 0) nidx[nc++] = i - 1;
      else if (s2[i - 1] == 0 && s2[i] != 0) pidx[pc++] = i;
    }
    // FixStep3 (:215-231): forward extension
    double* s3 = s1;
    for (int i = 0; i < F; ++i) s3[i] = s2[i];
    for (int i = 0; i < nc; ++i) {
      const int limit = i == nc - 1 ? F - 1 : nidx[i + 1];
      for (int j = nidx[i]; j < limit; ++j) {
        s3[j + 1] = dio_select_best(s3[j], s3[j - 1], cd + (j + 1), total_frames, c.nb, 0, c.allowed_range);
        if (s3[j + 1] == 0) break;
      }
    }
    // FixStep4 (:237-253): backward extension
    for (int i = 0; i < F; ++i) out[i] = s3[i];
    for (int i = pc - 1; i >= 0; --i) {
      const int limit = i == 0 ? 1 : pidx[i - 1];
      for (int j = pidx[i]; j > limit; --j) {
        out[j - 1] = dio_select_best(out[j], out[j + 1], cd + (j - 1), total_frames, c.nb, 0, c.allowed_range);
        if (out[j - 1] == 0) break;
      }
    }
  }
}

std::map<std::vector<double>, DioFilterBank*> g_banks;   // keyed by (actual_fs, floor, ceil, channels)

}  // namespace

bool dio_run(Batch* b, const DioParams& p, double* d_f0_out) {
  Context* ctxp = ctx();
  if (!ctxp) return false;
  cudaStream_t st = ctxp->stream;
  const int n_utt = b->n_utt;
  if (n_utt == 0) return true;
  const int ratio = std::max(std::min(p.speed, 12), 1);
  if (ratio != 1) { set_error("Dio: option.speed = %d (decimation) is not implemented yet; the analysis tool uses 1", p.speed); return false; }
  const double actual_fs = (double)b->fs / ratio;
  const std::vector<double> key = {actual_fs, p.f0_floor, p.f0_ceil, p.channels_in_octave};
  DioFilterBank* fb = nullptr;
  auto it = g_banks.find(key);
  if (it == g_banks.end()) {
    fb = new DioFilterBank();
    if (!build_filter_bank(actual_fs, p, fb)) { delete fb; return false; }
    g_banks[key] = fb;
  } else {
    fb = it->second;
  }
  DioConst c;
  c.nb = fb->nb; c.bn = fb->bn; c.log2bn = fb->log2bn; c.hN = fb->hN; c.D = fb->D; c.V = fb->V;
  for (int i = 0; i < fb->nb; ++i) { c.hal[i] = fb->hal[i]; c.boundary_f0[i] = fb->boundary_f0[i]; }
  c.actual_fs = actual_fs; c.f0_floor = p.f0_floor; c.f0_ceil = p.f0_ceil; c.allowed_range = p.allowed_range;
  c.voice_range_minimum = static_cast<int>(0.5 + 1000.0 / p.frame_period / p.f0_floor) * 2 + 1;   // :263-264

  // per-utterance sizes
  std::vector<int> h_ylen(n_utt), h_mask(n_utt);
  for (int u = 0; u < n_utt; ++u) {
    h_ylen[u] = 1 + b->h_x_len[u] / ratio;                                                         // :589
    const int sample = h_ylen[u] + 4 * static_cast<int>(1.0 + actual_fs / fb->boundary_f0[0] / 2.0);  // :592-593
    const int fft_size = static_cast<int>(pow(2.0, static_cast<int>(log(static_cast<double>(sample)) / kLog2) + 1.0));
    h_mask[u] = fft_size - 1;
  }
  DevBuf<int> d_ylen, d_mask;
  DevBuf<double> d_mean, d_cand, d_score, d_tmp1, d_tmp2;
  DevBuf<int> d_pos, d_neg;
  const int TF = b->total_frames;
  if (!d_ylen.alloc(n_utt) || !d_mask.alloc(n_utt) || !d_mean.alloc(n_utt) ||
      !d_cand.alloc((size_t)c.nb * TF) || !d_score.alloc((size_t)c.nb * TF) || !d_tmp1.alloc(TF) ||
      !d_tmp2.alloc(TF) || !d_pos.alloc(TF) || !d_neg.alloc(TF))
    return false;
  WB_CUDA_OR_RETURN(cudaMemcpyAsync(d_ylen.p, h_ylen.data(), n_utt * sizeof(int), cudaMemcpyHostToDevice, st), false);
  WB_CUDA_OR_RETURN(cudaMemcpyAsync(d_mask.p, h_mask.data(), n_utt * sizeof(int), cudaMemcpyHostToDevice, st), false);
  WB_CUDA_OR_RETURN(cudaMemsetAsync(d_f0_out, 0, (size_t)TF * sizeof(double), st), false);
  dio_mean_kernel<<<n_utt, 256, 0, st>>>(b->x.p, b->x_off.p, b->x_len.p, d_ylen.p, d_mean.p);
  WB_LAUNCH_CHECK();

  const size_t smem = 2 * cpad_size(c.bn / 2) * sizeof(double2);
  WB_CUDA_OR_RETURN(cudaFuncSetAttribute(dio_filter_kernel<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), false);
  WB_CUDA_OR_RETURN(cudaFuncSetAttribute(dio_filter_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), false);

  // sub-batches bounded by the size of the filtered-signal scratch (nb * y_len doubles per utterance)
  const size_t kMaxScratchDoubles = (size_t)2 << 30;           // 16 GiB
  int u0 = 0;
  while (u0 < n_utt) {
    int u1 = u0;
    size_t tot = 0;
    int max_y = 0;
    std::vector<long long> h_foff;
    while (u1 < n_utt) {
      const size_t need = (size_t)c.nb * h_ylen[u1] + 8;
      if (u1 > u0 && tot + need > kMaxScratchDoubles) break;
      h_foff.push_back((long long)tot);
      tot += need;
      max_y = std::max(max_y, h_ylen[u1]);
      ++u1;
    }
    const int nu = u1 - u0;
    DevBuf<double> d_F, d_edges;
    DevBuf<long long> d_foff, d_loff;
    DevBuf<int> d_counts, d_ltot;
    const int n_blocks = (max_y + c.V - 1) / c.V;
    const int n_chunks = (max_y + kZcChunk - 1) / kZcChunk;
    const int n_lists = nu * c.nb * 4;
    if (!d_F.alloc(tot) || !d_foff.alloc(nu) || !d_counts.alloc((size_t)n_lists * n_chunks) ||
        !d_ltot.alloc(n_lists) || !d_loff.alloc(n_lists))
      return false;
    WB_CUDA_OR_RETURN(cudaMemcpyAsync(d_foff.p, h_foff.data(), nu * sizeof(long long), cudaMemcpyHostToDevice, st), false);
    KernelTimer kt1("dio_filter_kernel");
    if (c.log2bn == 13)
      dio_filter_kernel<13><<<dim3(n_blocks, nu), 256, smem, st>>>(b->x.p, b->x_off.p, b->x_len.p, d_ylen.p, d_mask.p, d_mean.p,
                                                            d_foff.p, fb->G.p, ctxp->d_twiddle, c, u0, d_F.p);
    else
      dio_filter_kernel<0><<<dim3(n_blocks, nu), 256, smem, st>>>(b->x.p, b->x_off.p, b->x_len.p, d_ylen.p, d_mask.p, d_mean.p,
                                                            d_foff.p, fb->G.p, ctxp->d_twiddle, c, u0, d_F.p);
    WB_LAUNCH_CHECK(); kt1.stop();
    KernelTimer kt2("dio_zc_kernel");
    dio_zc_kernel<false><<<dim3(n_chunks, nu * c.nb), 256, 0, st>>>(d_F.p, d_foff.p, d_ylen.p, c.nb, u0, n_chunks, d_counts.p, nullptr, nullptr);
    WB_LAUNCH_CHECK(); kt2.stop();
    dio_zc_scan_kernel<<<(n_lists + 127) / 128, 128, 0, st>>>(d_counts.p, n_lists, n_chunks, d_ltot.p);
    WB_LAUNCH_CHECK();
    std::vector<int> h_ltot(n_lists);
    WB_CUDA_OR_RETURN(cudaMemcpyAsync(h_ltot.data(), d_ltot.p, n_lists * sizeof(int), cudaMemcpyDeviceToHost, st), false);
    WB_CUDA_OR_RETURN(cudaStreamSynchronize(st), false);
    std::vector<long long> h_loff(n_lists);
    long long etot = 0;
    for (int l = 0; l < n_lists; ++l) { h_loff[l] = etot; etot += h_ltot[l]; }
    if (!d_edges.alloc((size_t)etot + 2)) return false;
    WB_CUDA_OR_RETURN(cudaMemcpyAsync(d_loff.p, h_loff.data(), n_lists * sizeof(long long), cudaMemcpyHostToDevice, st), false);
    KernelTimer kt3("dio_zc_kernel");
    dio_zc_kernel<true><<<dim3(n_chunks, nu * c.nb), 256, 0, st>>>(d_F.p, d_foff.p, d_ylen.p, c.nb, u0, n_chunks, d_counts.p, d_loff.p, d_edges.p);
    WB_LAUNCH_CHECK(); kt3.stop();
    int max_f = 0;
    for (int u = u0; u < u1; ++u) max_f = std::max(max_f, b->h_f_len[u]);
    if (max_f > 0) {
      KernelTimer kt4("dio_candidates_kernel");
      dio_candidates_kernel<<<dim3((max_f + 127) / 128, nu * c.nb), 128, 0, st>>>(d_edges.p, d_loff.p, d_ltot.p, b->f_off.p, b->f_len.p,
                                                                               b->frame_t.p, c, u0, TF, d_cand.p, d_score.p);
      WB_LAUNCH_CHECK(); kt4.stop();
    }
    KernelTimer kt5("dio_fix_kernel");
    dio_fix_kernel<<<nu, 256, 0, st>>>(d_cand.p, d_score.p, b->f_off.p, b->f_len.p, c, u0, TF, d_tmp1.p, d_tmp2.p, d_pos.p, d_neg.p, d_f0_out);
    WB_LAUNCH_CHECK(); kt5.stop();
    WB_CUDA_OR_RETURN(cudaStreamSynchronize(st), false);      // scratch buffers die with this scope
    u0 = u1;
  }
  return true;
}

}  // namespace wb
