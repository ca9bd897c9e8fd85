// world-b200: Harvest F0 estimation over a batch of utterances.
//
// Reference: W/src/harvest.cpp — Harvest :1223-1255, HarvestGeneralBody :1145-1215,
// GetWaveformAndSpectrum(+Sub) :43-93, GetFilteredSignal :99-148, zero-crossing engine :162-238,
// GetF0CandidateContour(+Sub) :240-293, DetectOfficialF0Candidates :348-412, OverlapF0Candidates
// :417-429, GetRefinedF0 / GetMeanF0 / FixF0 :433-631, RemoveUnreliableCandidates :652-688,
// FixF0Contour :1027-1044 (SearchF0Base :693-706, FixStep1 :711-724, GetBoundaryList :729-745,
// FixStep2 :750-764, GetMultiChannelF0 :769-781, ExtendF0 :794-823, ExtendSub :845-862, Extend
// :867-883, MakeSortedOrder :888-901, SearchScore :906-912, MergeF0Sub :917-939, MergeF0 :944-971,
// FixStep3 :976-1004, FixStep4 :1009-1032), SmoothF0Contour / FilteringF0 :1049-1113;
// W/src/matlabfunctions.cpp decimate :184-210, FilterForDecimate :27-125.
//
// Pipeline (all on the device, batched over utterances):
//   1. decimation to ~8 kHz: the reference's zero-phase 3rd-order IIR (forward, then backward)
//      runs as chunks with a 512-sample warm-up (the slowest pole has radius 0.89, so the
//      warm-up transient is below 1e-25 relative), one thread per chunk;
//   2. ~150 band-pass channels by overlap-save (wb_zerocross.cuh), zero-crossing compaction and
//      interp1 onto the 1 ms frame grid exactly as in Dio;
//   3. per 1 ms frame: runs of >= 10 adjacent channels -> base candidates; +-3 frame overlap;
//   4. candidate refinement (instantaneous frequency at <= 6 harmonics): ONE WARP per
//      (frame, candidate) evaluates the <= 6 needed bins of the two windowed spectra directly
//      (windows are <= ~600 samples at 8 kHz, so a direct DFT of 6 bins with phasor recurrences
//      costs less than the reference's two FFTs of 512..2048 points and needs no shared memory);
//   5. contour logic (FixStep1-4, smoothing): data-dependent sequential recurrences along the
//      frames; one warp per utterance (lane 0 walks, the loops over candidates are short).
// The reference reads uninitialised heap memory in two places (RemoveUnreliableCandidates rows
// 0 and F-1, FixStep1 for frames whose base f0 is 0); both are treated as zeros here.
#include <math.h>
#include <algorithm>
#include <map>
#include <vector>
#include <chrono>
#include "wb_batch.h"
#include "wb_fft.cuh"
#include "wb_zerocross.cuh"

namespace wb {
namespace {

constexpr int kIirChunk = 512, kIirWarm = 512;
constexpr int kOverlap = 7;           // overlap_parameter (:1185)

struct HarvestConst {
  int fs, r, nch, lag;
  double actual_fs, f0_floor, f0_ceil;
  double a[3], b[2];
};

struct HarvestBank {
  int nch = 0, bn = 0, log2bn = 0, D = 0, V = 0;
  std::vector<double> boundary;
  DevBuf<double2> G;          // [nch][bn/2 + 1]
  DevBuf<int> shift;          // [nch]
  DevBuf<double> d_boundary;  // [nch]
};

bool decimate_coefficients(int r, double* a, double* b) {   // W/src/matlabfunctions.cpp:29-112
  static const double tab[13][5] = {
      {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0},
      {0.041156734567757189, -0.42599112459189636, 0.041037215479961225, 0.16797464681802227, 0.50392394045406674},
      {0.95039378983237421, -0.67429146741526791, 0.15412211621346475, 0.071221945171178636, 0.21366583551353591},
      {1.4499664446880227, -0.98943497080950582, 0.24578252340690215, 0.036710750339322612, 0.11013225101796784},
      {1.7610939654280557, -1.2554914843859768, 0.3237186507788215, 0.021334858522387423, 0.06400457556716227},
      {1.9715352749512141, -1.4686795689225347, 0.3893908434965701, 0.013469181309343825, 0.040407543928031475},
      {2.1225239019534703, -1.6395144861046302, 0.44469707800587366, 0.0090366882681608418, 0.027110064804482525},
      {2.2357462340187593, -1.7780899984041358, 0.49152555365968692, 0.0063522763407111993, 0.019056829022133598},
      {2.3236003491759578, -1.8921545617463598, 0.53148928133729068, 0.0046331164041389372, 0.013899349212416812},
      {2.3936475118069387, -1.9873904075111861, 0.5658879979027055, 0.0034818622251927556, 0.010445586675578267},
      {2.450743295230728, -2.06794904601978, 0.59574774438332101, 0.0026822508007163792, 0.0080467524021491377},
      {2.4981398605924205, -2.1368928194784025, 0.62187513816221485, 0.0021097275904709001, 0.0063291827714127002}};
  if (r < 2 || r > 12) return false;
  a[0] = tab[r][0]; a[1] = tab[r][1]; a[2] = tab[r][2]; b[0] = tab[r][3]; b[1] = tab[r][4];
  return true;
}

// ---- 1. decimation ------------------------------------------------------------------------------
// extended input of decimate(): x' = [x[0] * lag, x, x[last] * lag] (harvest.cpp:51-58), then the
// 9-sample odd reflection of matlabfunctions.cpp:189-193.  n in [0, L + 18), L = x_len + 2 lag.
__device__ __forceinline__ double harvest_xp(const double* __restrict__ x, int x_len, int lag, int j) {
  return x[min(x_len - 1, max(0, j - lag))];
}
__device__ __forceinline__ double harvest_ext(const double* __restrict__ x, int x_len, int lag, int L, int n) {
  if (n < 9) return 2.0 * harvest_xp(x, x_len, lag, 0) - harvest_xp(x, x_len, lag, 9 - n);
  if (n < 9 + L) return harvest_xp(x, x_len, lag, n - 9);
  return 2.0 * harvest_xp(x, x_len, lag, L - 1) - harvest_xp(x, x_len, lag, L - 2 - (n - (9 + L)));
}

// forward pass: B[n] for the chunk of this thread
__global__ void harvest_iir_fwd_kernel(const double* __restrict__ x_all, const long long* __restrict__ x_off,
                                       const int* __restrict__ x_len_all, const long long* __restrict__ B_off,
                                       HarvestConst c, int n_chunks_max, double* __restrict__ B) {
  const int u = blockIdx.y;
  const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
  if (chunk >= n_chunks_max) return;
  const int x_len = x_len_all[u];
  const int L = x_len + 2 * c.lag, M = L + 18;
  const int lo = chunk * kIirChunk;
  if (lo >= M) return;
  const int hi = min(M, lo + kIirChunk);
  const double* __restrict__ x = x_all + x_off[u];
  double* __restrict__ out = B + B_off[u];
  double w0 = 0.0, w1 = 0.0, w2 = 0.0;
  for (int n = max(0, lo - kIirWarm); n < hi; ++n) {
    const double wt = harvest_ext(x, x_len, c.lag, L, n) + c.a[0] * w0 + c.a[1] * w1 + c.a[2] * w2;
    if (n >= lo) out[n] = c.b[0] * wt + c.b[1] * w0 + c.b[1] * w1 + c.b[0] * w2;
    w2 = w1; w1 = w0; w0 = wt;
  }
}

// backward pass over B, keeping every r-th sample: y[i] = Z[nbeg + (lag / r + i) r + 8]
__global__ void harvest_iir_bwd_kernel(const double* __restrict__ B, const long long* __restrict__ B_off,
                                       const int* __restrict__ x_len_all, const long long* __restrict__ y_off,
                                       const int* __restrict__ y_len_all, HarvestConst c, int n_chunks_max,
                                       double* __restrict__ y_all) {
  const int u = blockIdx.y;
  const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
  if (chunk >= n_chunks_max) return;
  const int x_len = x_len_all[u];
  const int L = x_len + 2 * c.lag, M = L + 18;
  const int lo = chunk * kIirChunk;
  if (lo >= M) return;
  const int hi = min(M, lo + kIirChunk);
  const double* __restrict__ in = B + B_off[u];
  double* __restrict__ y = y_all + y_off[u];
  const int y_len = y_len_all[u];
  const int nout = (L - 1) / c.r + 1;                       // matlabfunctions.cpp:202-203
  const int nbeg = c.r - c.r * nout + L;
  const int first = nbeg + (c.lag / c.r) * c.r + 8;         // Z index of y[0]
  double w0 = 0.0, w1 = 0.0, w2 = 0.0;
  for (int n = min(M - 1, hi - 1 + kIirWarm); n >= lo; --n) {
    const double wt = in[n] + c.a[0] * w0 + c.a[1] * w1 + c.a[2] * w2;
    if (n < hi) {
      const int d = n - first;
      if (d >= 0 && d % c.r == 0 && d / c.r < y_len)
        y[d / c.r] = c.b[0] * wt + c.b[1] * w0 + c.b[1] * w1 + c.b[0] * w2;
    }
    w2 = w1; w1 = w0; w0 = wt;
  }
}

__global__ void harvest_copy_kernel(const double* __restrict__ x_all, const long long* __restrict__ x_off,
                                    const long long* __restrict__ y_off, const int* __restrict__ y_len_all,
                                    double* __restrict__ y_all) {
  const int u = blockIdx.y;
  const int n = y_len_all[u];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    y_all[y_off[u] + i] = x_all[x_off[u] + i];
}

__global__ void harvest_mean_kernel(const double* __restrict__ y_all, const long long* __restrict__ y_off,
                                    const int* __restrict__ y_len_all, double* __restrict__ mean) {
  __shared__ double red[96];
  const int u = blockIdx.x;
  const double* __restrict__ y = y_all + y_off[u];
  const int n = y_len_all[u];
  double v[1] = {0.0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) v[0] += y[i];
  block_sum<1>(v, red);
  if (threadIdx.x == 0) mean[u] = v[0] / n;
}

// ---- 2. raw candidates per (utterance, channel, 1 ms frame) (:240-293) ---------------------------
// One CTA = 128 consecutive 1 ms frames of one (utterance, channel).  Every frame interpolates the four
// interval contours at its time: a binary search over the edge list each.  The frames of a CTA span 128 ms,
// i.e. a handful to ~120 consecutive knots of each list, so the CTA first finds that range once (one thread
// per list), copies the edges of the range into shared memory, and the frames search there -- ~10 dependent
// L2 round trips per frame and list became one per CTA and list.
constexpr int kRawWin = 224;                 // edges per list kept in shared memory (a CTA needs <= ~125)
__global__ void __launch_bounds__(128)
harvest_raw_kernel(const double* __restrict__ edges, const long long* __restrict__ list_off,
                   const int* __restrict__ list_cnt, const int* __restrict__ g_off,
                   const int* __restrict__ g_len, const double* __restrict__ boundary,
                   HarvestConst c, int utt0, const long long* __restrict__ raw_off,
                   double* __restrict__ raw) {
  __shared__ double win[4][kRawWin];
  __shared__ int w_g0[4], w_lo[4], w_hi[4], w_nint[4], w_ok;
  const int u_local = blockIdx.y / c.nch, ch = blockIdx.y % c.nch;
  const int u = utt0 + u_local;
  const int n_fr = g_len[u];
  const int i0 = blockIdx.x * blockDim.x;
  if (i0 >= n_fr) return;
  const int i = i0 + threadIdx.x;
  const size_t l0 = ((size_t)u_local * c.nch + ch) * 4;
  const int i_last = min(n_fr, i0 + (int)blockDim.x) - 1;
  if (threadIdx.x == 0) w_ok = 1;
  __syncthreads();
  if (threadIdx.x < 4) {
    const int q = threadIdx.x;
    const int cnt = list_cnt[l0 + q];
    const int n_int = cnt < 2 ? 0 : cnt - 1;
    w_nint[q] = n_int;
    if (!(n_int - 2 > 0)) atomicAnd(&w_ok, 0);       // CheckEvent (:262-269)
    else {
      const double* __restrict__ e = edges + list_off[l0 + q];
      const int lo = zc_upper_bound(e, 0, n_int, c.actual_fs, div_rn((double)i0, 1000.0));
      const int hi = zc_upper_bound(e, lo, n_int, c.actual_fs, div_rn((double)i_last, 1000.0));
      w_lo[q] = lo; w_hi[q] = hi;
      w_g0[q] = max(0, min(lo, n_int - 1) - 1);      // lowest edge any frame of the CTA reads: e[k - 1], k = clamp(ub, 1, n_int - 1)
    }
  }
  __syncthreads();
  const bool ok = w_ok != 0;
  bool fits = true;
  if (ok) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int g1 = min(w_nint[q], max(1, min(w_hi[q], w_nint[q] - 1)) + 1);   // highest edge read: e[k + 1]
      const int n = g1 - w_g0[q] + 1;
      fits = fits && n <= kRawWin;
      if (n <= kRawWin) {
        const double* __restrict__ e = edges + list_off[l0 + q] + w_g0[q];
        for (int j = threadIdx.x; j < n; j += blockDim.x) win[q][j] = e[j];
      }
    }
  }
  __syncthreads();
  if (i >= n_fr) return;
  double cd = 0.0;
  if (ok) {
    const double t = div_rn((double)i, 1000.0);          // i * 1 / 1000.0 (:1176)
    double v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      v[q] = fits ? zc_interp_window(win[q], w_g0[q], w_lo[q], w_hi[q], w_nint[q], c.actual_fs, t)
                  : zc_interp(edges + list_off[l0 + q], w_nint[q], c.actual_fs, t);
    cd = div_rn(add_rn(add_rn(add_rn(v[0], v[1]), v[2]), v[3]), 4.0);
    const double bf = boundary[ch];
    if (cd > mul_rn(bf, 1.1) || cd < mul_rn(bf, 0.9) || cd > c.f0_ceil || cd < c.f0_floor) cd = 0.0;   // :243-253
  }
  raw[raw_off[u_local] + (size_t)ch * n_fr + i] = cd;
}

// ---- 3. base candidates per frame: runs of >= 10 adjacent voiced channels (:348-412) ------------
__global__ void harvest_detect_kernel(const double* __restrict__ raw, const long long* __restrict__ raw_off,
                                      const int* __restrict__ g_off, const int* __restrict__ g_len, int nch,
                                      int max_base, int utt0, double* __restrict__ base, int* __restrict__ nc_utt) {
  const int u_local = blockIdx.y, u = utt0 + u_local;
  const int n_fr = g_len[u];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_fr) return;
  const double* __restrict__ r = raw + raw_off[u_local] + i;
  double* __restrict__ out = base + ((size_t)g_off[u] + i) * max_base;
  int count = 0, st = 0, prev = 0;
  double acc = 0.0;
  for (int ch = 1; ch < nch; ++ch) {
    const double v = r[(size_t)ch * n_fr];
    const int cur = (ch < nch - 1 && v > 0.0) ? 1 : 0;           // vuv[0] = vuv[nch-1] = 0
    if (cur && !prev) { st = ch; acc = 0.0; }
    if (!cur && prev) {
      if (ch - st >= 10 && count < max_base) out[count++] = acc / (ch - st);
    }
    if (cur) acc += v;
    prev = cur;
  }
  for (int k = count; k < max_base; ++k) out[k] = 0.0;
  if (count > 0) atomicMax(&nc_utt[u], count);
}

// candidate of (frame k, slot s) after OverlapF0Candidates (:417-429)
__device__ __forceinline__ double harvest_overlapped(const double* __restrict__ base_u, int max_base, int nc,
                                                     int n_fr, int k, int s) {
  const int q = s / nc, j = s - q * nc;
  int src = k;
  if (q >= 1 && q <= 3) src = k - q;
  else if (q >= 4) src = k + (q - 3);
  if (src < 0 || src >= n_fr) return 0.0;
  return base_u[(size_t)src * max_base + j];
}

// ---- 4. refinement: one warp per 1 ms frame, looping over its non-zero candidates (:433-631) ------
// GetRefinedF0 for one candidate, evaluated by a full warp; lane 0 returns the result.
__device__ __forceinline__ void harvest_refine_one(const double* __restrict__ y, int y_len, double mean,
                                                   const HarvestConst& c, int k, double f0c, int lane,
                                                   const double2* __restrict__ tw_c_base,
                                                   double* refined_out, double* score_out) {
  const double fs = c.actual_fs;
  const double pos = div_rn((double)k, 1000.0);
  const int hwl = static_cast<int>(add_rn(div_rn(mul_rn(1.5, fs), f0c), 1.0));          // :586
  const int W = 2 * hwl + 1;
  const double wlen = div_rn(add_rn(mul_rn(2.0, (double)hwl), 1.0), fs);                // :587
  const int log2fft = 2 + (31 - __clz(W));                                               // :591-592 (W odd)
  const int nfft = 1 << log2fft;
  const int basic_index = matlab_round(add_rn(mul_rn(add_rn(pos, div_rn((double)(-hwl), fs)), fs), 0.001));   // :436-437
  const int nh = min(static_cast<int>(fs / 2.0 / f0c), 6);                               // :570-571
  int bins[6];
#pragma unroll
  for (int h = 0; h < 6; ++h)
    bins[h] = matlab_round(mul_rn(div_rn(mul_rn(f0c, (double)nfft), fs), (double)(h + 1)));   // :513
  // window phase a_n = 2 pi tmp_n / wlen, tmp_n = (basic_index + n - 1) / fs - pos (:447-452): linear
  // in n, so one sincos for n = lane and angle-addition steps of 32 samples; the neighbours needed
  // by the differentiated window are one more angle addition (+- one sample).
  const double dturn = 2.0 / (wlen * fs);                            // angle step in units of pi
  double cd, sd, c32, s32, cs, sn;
  sincospi(dturn, &sd, &cd);
  sincospi(32.0 * dturn, &s32, &c32);
  sincospi(2.0 * add_rn(div_rn(basic_index + lane - 1.0, fs), -pos) / wlen, &sn, &cs);
  double acc[6][4];
#pragma unroll
  for (int h = 0; h < 6; ++h) { acc[h][0] = acc[h][1] = acc[h][2] = acc[h][3] = 0.0; }
  // phasors e^{-2 pi i bin n / nfft} for n = lane, advanced by 32 samples per iteration
  // (start and step come from the compact twiddle table of size nfft: exp(-2 pi i m / nfft) for
  // m <= nfft/2, negated for the other half turn -- twelve table loads instead of twelve sincospi)
  double pc[6], ps[6], qc[6], qs[6];
  {
    const double2* __restrict__ tw = tw_c_base + Context::tw_c_offset(log2fft);
    const int nhalf = nfft >> 1;
#pragma unroll
    for (int h = 0; h < 6; ++h) {
      const int m0 = (int)(((long long)bins[h] * lane) & (nfft - 1));
      const int m1 = (int)(((long long)bins[h] * 32) & (nfft - 1));
      double2 a = __ldg(&tw[m0 & (nhalf - 1)]), b = __ldg(&tw[m1 & (nhalf - 1)]);
      if (m0 & nhalf) { a.x = -a.x; a.y = -a.y; }
      if (m1 & nhalf) { b.x = -b.x; b.y = -b.y; }
      pc[h] = a.x; ps[h] = a.y; qc[h] = b.x; qs[h] = b.y;
    }
  }
  auto blackman = [](double cv) { return 0.42 + 0.5 * cv + 0.08 * (2.0 * cv * cv - 1.0); };
  for (int n = lane; n < W; n += 32) {
    const double w = blackman(cs);
    const double w_next = blackman(cs * cd - sn * sd), w_prev = blackman(cs * cd + sn * sd);
    double dw;                                                       // GetDiffWindow (:459-465)
    if (n == 0) dw = -w_next / 2.0;
    else if (n == W - 1) dw = w_prev / 2.0;
    else dw = -(w_next - w_prev) / 2.0;
    {
      const double t = cs * c32 - sn * s32;
      sn = sn * c32 + cs * s32;
      cs = t;
    }
    const int idx = max(0, min(y_len - 1, basic_index + n - 1));
    const double xv = y[idx] - mean;
    const double xm = xv * w, xd = xv * dw;
#pragma unroll
    for (int h = 0; h < 6; ++h) {
      acc[h][0] += xm * pc[h]; acc[h][1] += xm * ps[h];
      acc[h][2] += xd * pc[h]; acc[h][3] += xd * ps[h];
      const double t = pc[h] * qc[h] - ps[h] * qs[h];
      ps[h] = ps[h] * qc[h] + pc[h] * qs[h];
      pc[h] = t;
    }
  }
  // 24 sums over the warp by recursive halving: at every step a lane keeps half of its values and
  // hands the other half to its partner, so 12 + 6 + 3 values cross instead of 24 per step; the
  // last three values are finished with plain butterflies.  Lane (4 g + q') ends up with ... no
  // particular owner is needed: the totals are broadcast back below.
  {
    double v[24];
#pragma unroll
    for (int h = 0; h < 6; ++h)
#pragma unroll
      for (int q = 0; q < 4; ++q) v[4 * h + q] = acc[h][q];
    // step 1 (partner lane ^ 16): keep 12
    double a12[12];
    {
      const bool up = lane & 16;
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const double send = up ? v[i] : v[12 + i];
        const double keep = up ? v[12 + i] : v[i];
        a12[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
    }
    double a6[6];
    {
      const bool up = lane & 8;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const double send = up ? a12[i] : a12[6 + i];
        const double keep = up ? a12[6 + i] : a12[i];
        a6[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
    }
    double a3[3];
    {
      const bool up = lane & 4;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double send = up ? a6[i] : a6[3 + i];
        const double keep = up ? a6[3 + i] : a6[i];
        a3[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      a3[i] += __shfl_xor_sync(0xffffffffu, a3[i], 2);
      a3[i] += __shfl_xor_sync(0xffffffffu, a3[i], 1);
    }
    // lane group g = (lane >> 2) & 7 = (bit16, bit8, bit4) holds values  12 b16 + 6 b8 + 3 b4 + {0,1,2}
#pragma unroll
    for (int idx = 0; idx < 24; ++idx) {
      const int b16 = idx / 12, r12 = idx % 12, b8 = r12 / 6, r6 = r12 % 6, b4 = r6 / 3, i = r6 % 3;
      const int src = (b16 << 4) | (b8 << 3) | (b4 << 2);
      acc[idx >> 2][idx & 3] = __shfl_sync(0xffffffffu, a3[i], src);
    }
  }
  double numerator = 0.0, denominator = 0.0, sc = 0.0;             // FixF0 (:504-536)
  for (int h = 0; h < nh; ++h) {
    const double re = acc[h][0], im = acc[h][1], dre = acc[h][2], dim = acc[h][3];
    const double power = re * re + im * im;
    const double numer = re * dim - im * dre;
    const int index = bins[h];
    const double inst = power == 0.0 ? 0.0
        : add_rn(div_rn(mul_rn((double)index, fs), (double)nfft),
                 div_rn(div_rn(mul_rn(div_rn(numer, power), fs), 2.0), kPi));
    const double amp = sqrt(power);
    numerator += amp * inst;
    denominator += amp * (h + 1.0);
    sc += fabs((inst / (h + 1.0) - f0c) / f0c);
  }
  double refined = numerator / (denominator + kMySafeGuardMinimum);
  double rscore = 1.0 / (sc / nh + kMySafeGuardMinimum);
  if (refined < c.f0_floor || refined > c.f0_ceil || rscore < 2.5) { refined = 0.0; rscore = 0.0; }   // :598-602
  *refined_out = refined;
  *score_out = rscore;
}

__global__ void __launch_bounds__(256, 2)
harvest_refine_kernel(const double* __restrict__ y_all, const long long* __restrict__ y_off,
                      const int* __restrict__ y_len_all, const double* __restrict__ mean_all,
                      const double* __restrict__ base, const int* __restrict__ g_off,
                      const int* __restrict__ g_len, const int* __restrict__ nc_utt,
                      const long long* __restrict__ cand_off, int max_base, HarvestConst c, int n_utt,
                      long long total_frames, const double2* __restrict__ tw_c_base,
                      double* __restrict__ cand, double* __restrict__ score) {
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= total_frames) return;
  int lo = 0, hi = n_utt - 1;                       // utterance of this 1 ms frame
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (g_off[mid] <= wid) lo = mid; else hi = mid - 1; }
  const int u = lo;
  const int nc = nc_utt[u], slots = nc * kOverlap, n_fr = g_len[u];
  const int k = (int)(wid - g_off[u]);
  const double* __restrict__ base_u = base + (size_t)g_off[u] * max_base;
  const double* __restrict__ y = y_all + y_off[u];
  const int y_len = y_len_all[u];
  const double mean = mean_all[u];
  const size_t o = cand_off[u] + (size_t)k * slots;
  for (int s = 0; s < slots; ++s) {
    const double f0c = harvest_overlapped(base_u, max_base, nc, n_fr, k, s);
    double refined = 0.0, rscore = 0.0;
    if (f0c > 0.0) harvest_refine_one(y, y_len, mean, c, k, f0c, lane, tw_c_base, &refined, &rscore);
    if (lane == 0) { cand[o + s] = refined; score[o + s] = rscore; }
  }
}

// ---- 4b. refinement, one THREAD per candidate (default) ------------------------------------------------
// At the ~8 kHz the reference decimates to, a candidate's window is 2 (1.5 fs / f0 + 1) + 1 = 100 - 420
// samples: a warp per candidate spends more on its set-up (three sincospi, twelve table loads) and on
// the 24-value warp reduction than on the window itself.  Here every candidate is one thread that walks
// its window sample by sample: the phasors start at 1 and advance by one table value per bin, the
// window phase advances by one sample (the differentiated window needs w[n-1], w[n], w[n+1]: a
// three-value slide, one new Blackman value per sample), nothing is exchanged between threads.  Threads
// of a warp are consecutive 1 ms frames of the same candidate slot, i.e. neighbouring points of one
// F0 track: similar window lengths (little divergence) and overlapping sample ranges (L1).
__global__ void __launch_bounds__(128, 4)      // 128 registers: four CTAs per SM (measured -11 % against 156 registers / three CTAs)
harvest_refine_thread_kernel(const double* __restrict__ y_all, const long long* __restrict__ y_off,
                             const int* __restrict__ y_len_all, const double* __restrict__ mean_all,
                             const double* __restrict__ base, const int* __restrict__ g_off,
                             const int* __restrict__ g_len, const int* __restrict__ nc_utt,
                             const long long* __restrict__ cand_off, int max_base, HarvestConst c, int n_utt,
                             long long total_work, const double2* __restrict__ tw_c_base,
                             double* __restrict__ cand, double* __restrict__ score) {
  // One thread per (1 ms frame k, base candidate j); it serves the kOverlap slots q * nc + j of the frame,
  // i.e. the candidates the same F0 track had at frames k-3 .. k+3 (OverlapF0Candidates :417-429).  Their
  // values differ by a fraction of a per cent, so most of them share the window length and all six
  // harmonic bins -- and then the two windowed spectra at those bins are the same numbers: the window walk
  // is repeated only when (hwl, bins) changes (2-3 walks instead of 7), the rest of FixF0 (:504-536), which
  // does depend on the candidate itself, runs per slot.  Bit-identical to one walk per slot.
  const long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (w >= total_work) return;
  int lo = 0, hi = n_utt - 1;                       // cand_off / kOverlap is the work prefix: g_len * nc per utterance
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (cand_off[mid] <= w * kOverlap) lo = mid; else hi = mid - 1; }
  const int u = lo;
  const int nc = nc_utt[u], slots = nc * kOverlap, n_fr = g_len[u];
  const long long local = w - cand_off[u] / kOverlap;
  const int j = (int)(local / n_fr), k = (int)(local - (long long)j * n_fr);      // frame fastest
  const double* __restrict__ base_u = base + (size_t)g_off[u] * max_base;
  const double* __restrict__ y = y_all + y_off[u];
  const int y_len = y_len_all[u];
  const double mean = mean_all[u];
  const double fs = c.actual_fs;
  const double pos = div_rn((double)k, 1000.0);
  auto blackman = [](double cv) { return 0.42 + 0.5 * cv + 0.08 * (2.0 * cv * cv - 1.0); };
  int key_hwl = -1, key_bins[6] = {0, 0, 0, 0, 0, 0};
  double acc[6][4];
#pragma unroll 1
  for (int q = 0; q < kOverlap; ++q) {
    const int s = q * nc + j;
    const size_t o = cand_off[u] + (size_t)k * slots + s;
    const double f0c = harvest_overlapped(base_u, max_base, nc, n_fr, k, s);
    if (!(f0c > 0.0)) { cand[o] = 0.0; score[o] = 0.0; continue; }
    const int hwl = static_cast<int>(add_rn(div_rn(mul_rn(1.5, fs), f0c), 1.0));          // :586
    const int W = 2 * hwl + 1;
    const int log2fft = 2 + (31 - __clz(W));                                               // :591-592 (W odd)
    const int nfft = 1 << log2fft, nhalf = nfft >> 1;
    const int nh = min(static_cast<int>(fs / 2.0 / f0c), 6);                               // :570-571
    int bins[6];
    bool same = hwl == key_hwl;
#pragma unroll
    for (int h = 0; h < 6; ++h) {
      bins[h] = matlab_round(mul_rn(div_rn(mul_rn(f0c, (double)nfft), fs), (double)(h + 1)));   // :513
      same = same && bins[h] == key_bins[h];
    }
    if (!same) {
      key_hwl = hwl;
      const double wlen = div_rn(add_rn(mul_rn(2.0, (double)hwl), 1.0), fs);              // :587
      const int basic_index = matlab_round(add_rn(mul_rn(add_rn(pos, div_rn((double)(-hwl), fs)), fs), 0.001));   // :436-437
      // The two spectra (window, differentiated window) at the <= 6 harmonic bins by GOERTZEL recurrences:
      // s[n] = x[n] + 2 cos(w) s[n-1] - s[n-2] costs one addition and one FMA per sample, bin and sequence
      // (24 FP64 operations per sample for 6 bins x 2 sequences) where a rotating phasor with its four
      // multiply-adds costs 48; at the end  sum x[n] e^{-i w n} = e^{-i w (W-1)} (s[W-1] - e^{-i w} s[W-2]),
      // the closing phasor taken from the twiddle table (index bin (W - 1) mod nfft, exact).  The error of
      // the recurrence grows like W / sin(w): <= 1e-12 here (W <= 340, w >= 0.05), F0 to ~1e-10.
      double k2[6], qc[6], qs[6], sm1[6], sm2[6], sd1[6], sd2[6];
      const double2* __restrict__ tw = tw_c_base + Context::tw_c_offset(log2fft);
      auto tw_at = [&](int m) {                                   // exp(-2 pi i m / nfft), m in [0, nfft)
        double2 b = __ldg(&tw[m & (nhalf - 1)]);
        if (m & nhalf) { b.x = -b.x; b.y = -b.y; }
        return b;
      };
#pragma unroll
      for (int h = 0; h < 6; ++h) {
        key_bins[h] = bins[h];
        const double2 b = tw_at(bins[h] & (nfft - 1));
        qc[h] = b.x; qs[h] = b.y; k2[h] = 2.0 * b.x;
        sm1[h] = sm2[h] = sd1[h] = sd2[h] = 0.0;
      }
      // window phase a_n = 2 pi ((basic_index + n - 1) / fs - pos) / wlen (:447-452), advanced one sample at a
      // time; the differentiated window (:459-465) needs w[n-1], w[n], w[n+1]: a three-value slide
      const double dturn = 2.0 / (wlen * fs);
      double cd, sd, cs, sn;
      sincospi(dturn, &sd, &cd);
      sincospi(2.0 * add_rn(div_rn(basic_index - 1.0, fs), -pos) / wlen, &sn, &cs);      // n = 0
      double w_prev = 0.0, w_cur = blackman(cs);
      auto sample = [&](int n, double* xm, double* xd) {
        {                                                             // phase of sample n + 1
          const double t = cs * cd - sn * sd;
          sn = sn * cd + cs * sd;
          cs = t;
        }
        const double w_next = blackman(cs);
        double dw;
        if (n == 0) dw = -w_next / 2.0;
        else if (n == W - 1) dw = w_prev / 2.0;
        else dw = -(w_next - w_prev) / 2.0;
        const int idx = max(0, min(y_len - 1, basic_index + n - 1));
        const double xv = y[idx] - mean;
        *xm = xv * w_cur; *xd = xv * dw;
        w_prev = w_cur;
        w_cur = w_next;
      };
      // two samples per trip: the roles of s[n-1] / s[n-2] alternate, no register moves; W is odd
      for (int n = 0; n + 1 < W; n += 2) {
        double xm0, xd0, xm1, xd1;
        sample(n, &xm0, &xd0);
        sample(n + 1, &xm1, &xd1);
#pragma unroll
        for (int h = 0; h < 6; ++h) {
          sm2[h] = fma(k2[h], sm1[h], xm0 - sm2[h]);
          sd2[h] = fma(k2[h], sd1[h], xd0 - sd2[h]);
          sm1[h] = fma(k2[h], sm2[h], xm1 - sm1[h]);
          sd1[h] = fma(k2[h], sd2[h], xd1 - sd1[h]);
        }
      }
      {
        double xm0, xd0;
        sample(W - 1, &xm0, &xd0);
#pragma unroll
        for (int h = 0; h < 6; ++h) {
          const double tm = fma(k2[h], sm1[h], xm0 - sm2[h]), td = fma(k2[h], sd1[h], xd0 - sd2[h]);
          // y = s[W-1] - e^{-i w} s[W-2];  X = e^{-i w (W-1)} y
          const double ymr = tm - qc[h] * sm1[h], ymi = -qs[h] * sm1[h];
          const double ydr = td - qc[h] * sd1[h], ydi = -qs[h] * sd1[h];
          const double2 e = tw_at((int)(((long long)bins[h] * (W - 1)) & (nfft - 1)));
          acc[h][0] = e.x * ymr - e.y * ymi; acc[h][1] = e.x * ymi + e.y * ymr;
          acc[h][2] = e.x * ydr - e.y * ydi; acc[h][3] = e.x * ydi + e.y * ydr;
        }
      }
    }
    double numerator = 0.0, denominator = 0.0, sc = 0.0;           // FixF0 (:504-536)
#pragma unroll
    for (int h = 0; h < 6; ++h) {
      if (h >= nh) break;
      const double re = acc[h][0], im = acc[h][1], dre = acc[h][2], dim = acc[h][3];
      const double power = re * re + im * im;
      const double numer = re * dim - im * dre;
      const double inst = power == 0.0 ? 0.0
          : add_rn(div_rn(mul_rn((double)bins[h], fs), (double)nfft),
                   div_rn(div_rn(mul_rn(div_rn(numer, power), fs), 2.0), kPi));
      const double amp = sqrt(power);
      numerator += amp * inst;
      denominator += amp * (h + 1.0);
      sc += fabs((inst / (h + 1.0) - f0c) / f0c);
    }
    double refined = numerator / (denominator + kMySafeGuardMinimum);
    double rscore = 1.0 / (sc / nh + kMySafeGuardMinimum);
    if (refined < c.f0_floor || refined > c.f0_ceil || rscore < 2.5) { refined = 0.0; rscore = 0.0; }   // :598-602
    cand[o] = refined;
    score[o] = rscore;
  }
}

// ---- RemoveUnreliableCandidates (:652-688) --------------------------------------------------------
__device__ __forceinline__ double harvest_min_rel_error(double ref, const double* __restrict__ row, int n) {
  double best = 1.0;                               // SelectBestF0 with allowed_range 1.0 (:636-650)
  for (int i = 0; i < n; ++i) {
    const double t = fabs(ref - row[i]) / ref;
    if (t > best) continue;
    best = t;
  }
  return best;
}

__global__ void harvest_unreliable_kernel(const double* __restrict__ cand_in, const double* __restrict__ score_in,
                                          const int* __restrict__ g_len, const int* __restrict__ nc_utt,
                                          const long long* __restrict__ cand_off, int n_utt,
                                          const long long* __restrict__ work_first, long long total_work,
                                          double* __restrict__ cand_out, double* __restrict__ score_out) {
  const long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (w >= total_work) return;
  int lo = 0, hi = n_utt - 1;
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (work_first[mid] <= w) lo = mid; else hi = mid - 1; }
  const int u = lo;
  const int slots = nc_utt[u] * kOverlap, n_fr = g_len[u];
  const long long local = w - work_first[u];
  const int k = (int)(local / slots), s = (int)(local - (long long)k * slots);
  const size_t o = cand_off[u] + (size_t)k * slots + s;
  double cv = cand_in[o], sv = score_in[o];
  if (k >= 1 && k < n_fr - 1 && cv != 0.0) {
    // rows 0 and n_fr-1 of the reference's copy are never written (uninitialised): zeros here
    const double* next = cand_in + cand_off[u] + (size_t)(k + 1) * slots;
    const double* prev = cand_in + cand_off[u] + (size_t)(k - 1) * slots;
    const double e1 = (k + 1 < n_fr - 1) ? harvest_min_rel_error(cv, next, slots) : 1.0;
    const double e2 = (k - 1 >= 1) ? harvest_min_rel_error(cv, prev, slots) : 1.0;
    if (fmin(e1, e2) > 0.05) { cv = 0.0; sv = 0.0; }
  }
  cand_out[o] = cv;
  score_out[o] = sv;
}

// ---- 5. contour logic ---------------------------------------------------------------------------------
// GetBoundaryList (:729-745): returns the number of boundaries; sections are [bl[2i], bl[2i+1]].
__device__ int harvest_boundaries(const double* __restrict__ f0, int n, int* __restrict__ bl) {
  int nb = 0, prev = 0;
  for (int i = 1; i < n; ++i) {
    const int cur = (i < n - 1 && f0[i] > 0.0) ? 1 : 0;
    if (cur != prev) { bl[nb] = i - nb % 2; ++nb; }
    prev = cur;
  }
  return nb;
}

// SelectBestF0 (:636-650)
__device__ double harvest_select_best(double ref, const double* __restrict__ row, int n, double allowed) {
  double best = 0.0, best_error = allowed;
  for (int i = 0; i < n; ++i) {
    const double t = fabs(ref - row[i]) / ref;
    if (t > best_error) continue;
    best = row[i];
    best_error = t;
  }
  return best;
}

// phase A: SearchF0Base, FixStep1 (allowed 0.008), FixStep2 (minimum 6) -> step2, section count
__global__ void harvest_fix_a_kernel(const double* __restrict__ cand, const double* __restrict__ score,
                                     const int* __restrict__ g_off, const int* __restrict__ g_len,
                                     const int* __restrict__ nc_utt, const long long* __restrict__ cand_off,
                                     double* tmp1, double* tmp2, int* bl_all,      // exchanged between threads: not __restrict__
                                     int* __restrict__ n_sections) {
  const int u = blockIdx.x;
  const int n = g_len[u], off = g_off[u], slots = nc_utt[u] * kOverlap;
  double* basef = tmp1 + off;
  double* step1 = tmp2 + off;
  const double* __restrict__ cu = cand + cand_off[u];
  const double* __restrict__ su = score + cand_off[u];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {              // SearchF0Base (:693-706)
    double bf = 0.0, bs = 0.0;
    for (int j = 0; j < slots; ++j) {
      const double sc = su[(size_t)i * slots + j];
      if (sc > bs) { bf = cu[(size_t)i * slots + j]; bs = sc; }
    }
    basef[i] = bf;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {              // FixStep1 (:711-724)
    double v = 0.0;
    if (i >= 2 && basef[i] != 0.0) {
      const double ref = basef[i - 1] * 2 - basef[i - 2];
      v = (fabs((basef[i] - ref) / ref) > 0.008 && fabs(basef[i] - basef[i - 1]) / basef[i - 1] > 0.008) ? 0.0 : basef[i];
    }
    step1[i] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {                                          // FixStep2 (:750-764)
    int* bl = bl_all + 2 * off;
    const int nb = harvest_boundaries(step1, n, bl);
    for (int i = 0; i < n; ++i) basef[i] = step1[i];               // step2 lives in tmp1
    for (int i = 0; i < nb / 2; ++i) {
      if (bl[i * 2 + 1] - bl[i * 2] >= 6) continue;
      for (int j = bl[i * 2]; j <= bl[i * 2 + 1]; ++j) basef[j] = 0.0;
    }
    n_sections[u] = harvest_boundaries(basef, n, bl) / 2;
  }
}

// SelectBestF0 (:636-650) by a warp: the sequential rule keeps the LAST candidate among those with the
// smallest relative error <= allowed; lanes scan strided subsets with the same rule and the partial
// results are joined by (smaller error, then larger index).  Every lane returns the result.
__device__ __forceinline__ double harvest_select_best_warp(double ref, const double* __restrict__ row, int n, double allowed,
                                                           int lane) {
  double best_error = allowed;
  int best_i = -1;
  for (int i = lane; i < n; i += 32) {
    const double t = fabs(ref - row[i]) / ref;
    if (t > best_error) continue;
    best_error = t;
    best_i = i;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double e = __shfl_xor_sync(0xffffffffu, best_error, o);
    const int i = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (i >= 0 && (best_i < 0 || e < best_error || (e == best_error && i > best_i))) { best_error = e; best_i = i; }
  }
  return best_i >= 0 ? row[best_i] : 0.0;
}

// phase B: FixStep3 (:976-1004) and FixStep4 (:1009-1032); one warp per utterance.  The walk along the
// contour is sequential by nature (every step depends on the F0 just chosen), but each step's search over
// the candidate slots, the score look-ups of MergeF0 and the contour copies are spread over the lanes;
// the scalar bookkeeping (means, boundaries, merge order) stays on lane 0 in the reference's order.
// mc: [sections][n] scratch of this utterance; chan: pointer permutation for Swap (:828-843).
__global__ void harvest_fix_b_kernel(const double* __restrict__ cand, const double* __restrict__ score,
                                     const int* __restrict__ g_off, const int* __restrict__ g_len,
                                     const int* __restrict__ nc_utt, const long long* __restrict__ cand_off,
                                     double* tmp1, double* tmp2, int* bl_all,
                                     const long long* __restrict__ mc_off, double* mc_all,
                                     int* chan_all, int* order_all, int dbg_mode = 3, int stop_after = 0) {
  // NOTE: the arrays the lanes exchange data through (tmp1 / tmp2 / bl / mc / chan / order) must NOT be
  // __restrict__: that qualifier promises the compiler that nobody else -- which includes the other
  // lanes -- writes the object, and it then keeps values in registers across __syncwarp() (measured: the
  // merged contour differed from the one-lane version on real speech; the state after Extend did not).
  const int u = blockIdx.x;
  const int lane = threadIdx.x;
  const int n = g_len[u], off = g_off[u], slots = nc_utt[u] * kOverlap;
  const double* step2 = tmp1 + off;
  double* step3 = tmp2 + off;
  const double* __restrict__ cu = cand + cand_off[u];
  const double* __restrict__ su = score + cand_off[u];
  int* bl = bl_all + 2 * off;
  double* mc = mc_all + mc_off[u];
  int* chan = chan_all + off;
  int* order = order_all + off;
  __shared__ int nb_s, nchn_s;
  for (int i = lane; i < n; i += 32) step3[i] = step2[i];
  if (lane == 0) nb_s = harvest_boundaries(step2, n, bl);
  __syncwarp();
  const int nsec = nb_s / 2;
  // GetMultiChannelF0 (:769-781)
  for (int s = 0; s < nsec; ++s)
    for (int j = lane; j < n; j += 32)
      mc[(size_t)s * n + j] = (j >= bl[2 * s] && j <= bl[2 * s + 1]) ? step2[j] : 0.0;
  for (int s = lane; s < nsec; s += 32) chan[s] = s;
  __syncwarp();
  // Extend (:867-883) with ExtendF0 (:794-823), in place on mc and bl
  for (int s = 0; s < nsec; ++s) {
    double* ext = mc + (size_t)s * n;
    for (int dir = 0; dir < 2; ++dir) {
      const int shift = dir == 0 ? 1 : -1;
      const int origin = dir == 0 ? bl[2 * s + 1] : bl[2 * s];
      const int last_point = dir == 0 ? min(n - 2, bl[2 * s + 1] + 100) : max(1, bl[2 * s] - 100);
      double tmp_f0 = ext[origin];
      int shifted_origin = origin, count = 0;
      const int distance = abs(last_point - origin);
      __syncwarp();                                                 // bl / ext read by every lane before lane 0 rewrites them
      for (int i = 0; i <= distance; ++i) {
        const int idx = origin + shift * i + shift;
        double v;
        if (dbg_mode & 1) v = harvest_select_best_warp(tmp_f0, cu + (size_t)idx * slots, slots, 0.18, lane);
        else {
          v = 0.0;
          if (lane == 0) v = harvest_select_best(tmp_f0, cu + (size_t)idx * slots, slots, 0.18);
          v = __shfl_sync(0xffffffffu, v, 0);
        }
        if (lane == 0) ext[idx] = v;
        if (v == 0.0) ++count;
        else { tmp_f0 = v; count = 0; shifted_origin = idx; }
        if (count == 4) break;
      }
      if (lane == 0) { if (dir == 0) bl[2 * s + 1] = shifted_origin; else bl[2 * s] = shifted_origin; }
      __syncwarp();
    }
  }
  if (stop_after == 1) return;
  if (lane == 0) {
    // ExtendSub (:845-862); mean_f0 is deliberately not reset between sections (as in the reference)
    int nchn = 0;
    double mean_f0 = 0.0;
    for (int s = 0; s < nsec; ++s) {
      const int st = bl[2 * s], ed = bl[2 * s + 1];
      const double* ext = mc + (size_t)chan[s] * n;
      for (int j = st; j < ed; ++j) mean_f0 += ext[j];
      mean_f0 /= ed - st;
      if (2200.0 / mean_f0 < ed - st) {
        const int a = nchn++, b2 = s;
        int t = chan[a]; chan[a] = chan[b2]; chan[b2] = t;
        t = bl[2 * a]; bl[2 * a] = bl[2 * b2]; bl[2 * b2] = t;
        t = bl[2 * a + 1]; bl[2 * a + 1] = bl[2 * b2 + 1]; bl[2 * b2 + 1] = t;
      }
    }
    if (nchn != 0) {
      for (int i = 0; i < nchn; ++i) order[i] = i;                  // MakeSortedOrder (:888-901)
      for (int i = 1; i < nchn; ++i)
        for (int j = i - 1; j >= 0; --j) {
          if (bl[order[j] * 2] > bl[order[i] * 2]) { const int t = order[i]; order[i] = order[j]; order[j] = t; }
          else break;
        }
    }
    nchn_s = nchn;
  }
  __syncwarp();
  if (stop_after == 2) return;
  const int nchn = nchn_s;
  if (nchn != 0) {
    // MergeF0 (:944-971)
    const double* ch0 = mc + (size_t)chan[0] * n;
    for (int i = lane; i < n; i += 32) step3[i] = ch0[i];
    __syncwarp();
    // bl[0] / bl[1] double as the merged contour's boundaries exactly as in the reference (which rewrites
    // boundary_list[0..1] in place, so a later channel whose slot is 0 sees the rewritten values)
    for (int i = 1; i < nchn; ++i) {
      const int oi = order[i];
      const double* f2 = mc + (size_t)chan[oi] * n;
      const int st2 = bl[oi * 2], ed2 = bl[oi * 2 + 1];
      const int st1 = bl[0], ed1 = bl[1];
      __syncwarp();                                                 // every lane has read bl before lane 0 rewrites it
      if (st2 - ed1 > 0) {
        for (int j = st2 + lane; j <= ed2; j += 32) step3[j] = f2[j];
        if (lane == 0) { bl[0] = st2; bl[1] = ed2; }
      } else if (!(st1 <= st2 && ed1 >= ed2)) {                     // MergeF0Sub (:917-939); otherwise bl[1] stays ed1
        double score1 = 0.0, score2 = 0.0;
        for (int j = st2; j <= ed1; ++j) {
          double s1 = 0.0, s2 = 0.0;                                // SearchScore (:906-912): a maximum, any order
          const double v1 = step3[j], v2 = f2[j];
          for (int q = (dbg_mode & 2) ? lane : 0; q < slots; q += (dbg_mode & 2) ? 32 : 1) {
            const double cq = cu[(size_t)j * slots + q], sq = su[(size_t)j * slots + q];
            if (v1 == cq && s1 < sq) s1 = sq;
            if (v2 == cq && s2 < sq) s2 = sq;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            s1 = fmax(s1, __shfl_xor_sync(0xffffffffu, s1, o));
            s2 = fmax(s2, __shfl_xor_sync(0xffffffffu, s2, o));
          }
          score1 += s1;
          score2 += s2;
        }
        __syncwarp();
        if (score1 > score2) { for (int j = ed1 + lane; j <= ed2; j += 32) step3[j] = f2[j]; }
        else { for (int j = st2 + lane; j <= ed2; j += 32) step3[j] = f2[j]; }
        if (lane == 0) bl[1] = ed2;
      }
      __syncwarp();
    }
  }
  __syncwarp();
  // FixStep4 (:1009-1032), threshold 9; result back into tmp1
  double* step4 = tmp1 + off;
  for (int i = lane; i < n; i += 32) step4[i] = step3[i];
  __syncwarp();
  if (lane == 0) {
    const int nb4 = harvest_boundaries(step3, n, bl);
    for (int i = 0; i < nb4 / 2 - 1; ++i) {
      const int distance = bl[(i + 1) * 2] - bl[i * 2 + 1] - 1;
      if (distance >= 9) continue;
      const double t0 = step3[bl[i * 2 + 1]] + 1, t1 = step3[bl[(i + 1) * 2]] - 1;
      const double coefficient = (t1 - t0) / (distance + 1.0);
      int count = 1;
      for (int j = bl[i * 2 + 1] + 1; j <= bl[(i + 1) * 2] - 1; ++j) step4[j] = t0 + coefficient * count++;
    }
  }
}

// phase B, reference implementation of the contour logic: lane 0 walks everything (WB_HARVEST_FIX_WARP=0).
// mc: [sections][n] scratch of this utterance; chan: pointer permutation for Swap (:828-843).
__global__ void harvest_fix_b_serial_kernel(const double* __restrict__ cand, const double* __restrict__ score,
                                     const int* __restrict__ g_off, const int* __restrict__ g_len,
                                     const int* __restrict__ nc_utt, const long long* __restrict__ cand_off,
                                     double* __restrict__ tmp1, double* __restrict__ tmp2, int* __restrict__ bl_all,
                                     const long long* __restrict__ mc_off, double* __restrict__ mc_all,
                                     int* __restrict__ chan_all, int* __restrict__ order_all, int stop_after = 0) {
  const int u = blockIdx.x;
  const int n = g_len[u], off = g_off[u], slots = nc_utt[u] * kOverlap;
  const double* __restrict__ step2 = tmp1 + off;
  double* __restrict__ step3 = tmp2 + off;
  const double* __restrict__ cu = cand + cand_off[u];
  const double* __restrict__ su = score + cand_off[u];
  int* bl = bl_all + 2 * off;
  double* mc = mc_all + mc_off[u];
  int* chan = chan_all + off;
  int* order = order_all + off;
  __shared__ int nb_s;
  for (int i = threadIdx.x; i < n; i += blockDim.x) step3[i] = step2[i];
  if (threadIdx.x == 0) nb_s = harvest_boundaries(step2, n, bl);
  __syncthreads();
  const int nsec = nb_s / 2;
  // GetMultiChannelF0 (:769-781)
  for (int s = 0; s < nsec; ++s)
    for (int j = threadIdx.x; j < n; j += blockDim.x)
      mc[(size_t)s * n + j] = (j >= bl[2 * s] && j <= bl[2 * s + 1]) ? step2[j] : 0.0;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int s = 0; s < nsec; ++s) chan[s] = s;
    // Extend (:867-883) with ExtendF0 (:794-823), in place on mc and bl
    for (int s = 0; s < nsec; ++s) {
      double* ext = mc + (size_t)s * n;
      for (int dir = 0; dir < 2; ++dir) {
        const int shift = dir == 0 ? 1 : -1;
        const int origin = dir == 0 ? bl[2 * s + 1] : bl[2 * s];
        const int last_point = dir == 0 ? min(n - 2, bl[2 * s + 1] + 100) : max(1, bl[2 * s] - 100);
        double tmp_f0 = ext[origin];
        int shifted_origin = origin, count = 0;
        const int distance = abs(last_point - origin);
        for (int i = 0; i <= distance; ++i) {
          const int idx = origin + shift * i + shift;
          const double v = harvest_select_best(tmp_f0, cu + (size_t)idx * slots, slots, 0.18);
          ext[idx] = v;
          if (v == 0.0) ++count;
          else { tmp_f0 = v; count = 0; shifted_origin = idx; }
          if (count == 4) break;
        }
        if (dir == 0) bl[2 * s + 1] = shifted_origin; else bl[2 * s] = shifted_origin;
      }
    }
    if (stop_after == 1) return;
    // ExtendSub (:845-862); mean_f0 is deliberately not reset between sections (as in the reference)
    int nchn = 0;
    double mean_f0 = 0.0;
    for (int s = 0; s < nsec; ++s) {
      const int st = bl[2 * s], ed = bl[2 * s + 1];
      const double* ext = mc + (size_t)chan[s] * n;
      for (int j = st; j < ed; ++j) mean_f0 += ext[j];
      mean_f0 /= ed - st;
      if (2200.0 / mean_f0 < ed - st) {
        const int a = nchn++, b2 = s;
        int t = chan[a]; chan[a] = chan[b2]; chan[b2] = t;
        t = bl[2 * a]; bl[2 * a] = bl[2 * b2]; bl[2 * b2] = t;
        t = bl[2 * a + 1]; bl[2 * a + 1] = bl[2 * b2 + 1]; bl[2 * b2 + 1] = t;
      }
    }
    if (stop_after == 2) return;
    if (nchn != 0) {
      // MergeF0 (:944-971)
      for (int i = 0; i < nchn; ++i) order[i] = i;                  // MakeSortedOrder (:888-901)
      for (int i = 1; i < nchn; ++i)
        for (int j = i - 1; j >= 0; --j) {
          if (bl[order[j] * 2] > bl[order[i] * 2]) { const int t = order[i]; order[i] = order[j]; order[j] = t; }
          else break;
        }
      const double* ch0 = mc + (size_t)chan[0] * n;
      for (int i = 0; i < n; ++i) step3[i] = ch0[i];
      for (int i = 1; i < nchn; ++i) {
        const int oi = order[i];
        const double* f2 = mc + (size_t)chan[oi] * n;
        const int st2 = bl[oi * 2], ed2 = bl[oi * 2 + 1];
        if (st2 - bl[1] > 0) {
          for (int j = st2; j <= ed2; ++j) step3[j] = f2[j];
          bl[0] = st2;
          bl[1] = ed2;
        } else {
          const int st1 = bl[0], ed1 = bl[1];                       // MergeF0Sub (:917-939)
          if (st1 <= st2 && ed1 >= ed2) { bl[1] = ed1; continue; }
          double score1 = 0.0, score2 = 0.0;
          for (int j = st2; j <= ed1; ++j) {
            double s1 = 0.0, s2 = 0.0;                              // SearchScore (:906-912)
            const double v1 = step3[j], v2 = f2[j];
            for (int q = 0; q < slots; ++q) {
              const double cq = cu[(size_t)j * slots + q], sq = su[(size_t)j * slots + q];
              if (v1 == cq && s1 < sq) s1 = sq;
              if (v2 == cq && s2 < sq) s2 = sq;
            }
            score1 += s1;
            score2 += s2;
          }
          if (score1 > score2) { for (int j = ed1; j <= ed2; ++j) step3[j] = f2[j]; }
          else { for (int j = st2; j <= ed2; ++j) step3[j] = f2[j]; }
          bl[1] = ed2;
        }
      }
    }
    // FixStep4 (:1009-1032), threshold 9; result back into tmp1
    double* step4 = tmp1 + off;
    for (int i = 0; i < n; ++i) step4[i] = step3[i];
    const int nb4 = harvest_boundaries(step3, n, bl);
    for (int i = 0; i < nb4 / 2 - 1; ++i) {
      const int distance = bl[(i + 1) * 2] - bl[i * 2 + 1] - 1;
      if (distance >= 9) continue;
      const double t0 = step3[bl[i * 2 + 1]] + 1, t1 = step3[bl[(i + 1) * 2]] - 1;
      const double coefficient = (t1 - t0) / (distance + 1.0);
      int count = 1;
      for (int j = bl[i * 2 + 1] + 1; j <= bl[(i + 1) * 2] - 1; ++j) step4[j] = t0 + coefficient * count++;
    }
  }
}

// ---- SmoothF0Contour (:1078-1113) ------------------------------------------------------------------
// one CTA per utterance finds the voiced sections of the zero-padded contour; one thread per
// section then runs the zero-lag Butterworth filter of FilteringF0 (:1049-1073) over the whole
// padded length exactly as the reference does.
constexpr int kSmoothLag = 300;
__global__ void harvest_sections_kernel(const double* __restrict__ tmp1, const int* __restrict__ g_off,
                                        const int* __restrict__ g_len, int* __restrict__ bl_all,
                                        int* __restrict__ n_sections) {
  const int u = blockIdx.x;
  if (threadIdx.x != 0) return;
  const int n = g_len[u], off = g_off[u];
  const double* f0 = tmp1 + off;
  // boundaries of the padded contour [0 * lag, f0, 0 * lag]: same as on f0 itself, shifted by lag,
  // except that f0[0] and f0[n-1] are no longer forced to be unvoiced
  int* bl = bl_all + 2 * off;
  int nb = 0, prev = 0;
  for (int i = 1; i < n + 2 * kSmoothLag; ++i) {
    const int j = i - kSmoothLag;
    const int cur = (j >= 0 && j < n && f0[j] > 0.0) ? 1 : 0;
    if (cur != prev) { bl[nb] = i - nb % 2; ++nb; }
    prev = cur;
  }
  n_sections[u] = nb / 2;
}

__global__ void harvest_smooth_kernel(const double* __restrict__ tmp1, const int* __restrict__ g_off,
                                      const int* __restrict__ g_len, const int* __restrict__ bl_all,
                                      const int* __restrict__ sec_first, int n_utt, int total_sections,
                                      const long long* __restrict__ scratch_off, double* __restrict__ scratch,
                                      double* __restrict__ f0_out) {
  const int sidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (sidx >= total_sections) return;
  int lo = 0, hi = n_utt - 1;
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (sec_first[mid] <= sidx) lo = mid; else hi = mid - 1; }
  const int u = lo, s = sidx - sec_first[u];
  const int n = g_len[u], off = g_off[u], len = n + 2 * kSmoothLag;
  const double* f0 = tmp1 + off;
  const int st = bl_all[2 * off + 2 * s], ed = bl_all[2 * off + 2 * s + 1];
  double* tmp_x = scratch + scratch_off[u] + (size_t)s * len;
  const double b0 = 0.0078202080334971724, b1 = 0.015640416066994345;
  const double a0 = 1.7347257688092754, a1 = -0.76600660094326412;
  auto xin = [&](int i) {                            // channel s with the edge extension of :1054-1055
    const int j = min(ed, max(st, i)) - kSmoothLag;
    return f0[j];
  };
  double w0 = 0.0, w1 = 0.0;
  for (int i = 0; i < len; ++i) {
    const double wt = xin(i) + a0 * w0 + a1 * w1;
    tmp_x[len - i - 1] = b0 * wt + b1 * w0 + b0 * w1;
    w1 = w0; w0 = wt;
  }
  w0 = w1 = 0.0;
  for (int i = 0; i < len; ++i) {
    const double wt = tmp_x[i] + a0 * w0 + a1 * w1;
    const double yv = b0 * wt + b1 * w0 + b0 * w1;
    const int pos = len - i - 1;
    if (pos >= st && pos <= ed) f0_out[off + pos - kSmoothLag] = yv;
    w1 = w0; w0 = wt;
  }
}

// final pick onto the caller's frame grid (:1246-1251)
__global__ void harvest_pick_kernel(const double* __restrict__ basic_f0, const int* __restrict__ g_off,
                                    const int* __restrict__ g_len, const int* __restrict__ f_off,
                                    const int* __restrict__ f_len, double frame_period, double* __restrict__ f0) {
  const int u = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= f_len[u]) return;
  const double t = div_rn(mul_rn((double)i, frame_period), 1000.0);
  const int k = min(g_len[u] - 1, matlab_round(mul_rn(t, 1000.0)));
  f0[f_off[u] + i] = basic_f0[g_off[u] + k];
}

std::map<std::vector<double>, HarvestBank*> g_hbanks;

bool build_harvest_bank(double actual_fs, double f0_floor, double f0_ceil, HarvestBank* hb) {
  const double channels_in_octave = 40.0;
  const double afloor = f0_floor * 0.9, aceil = f0_ceil * 1.1;                                  // :1149-1150
  hb->nch = 1 + static_cast<int>(log(aceil / afloor) / kLog2 * channels_in_octave);              // :1151-1153
  if (hb->nch < 12) { set_error("Harvest: only %d channels (f0 range too narrow)", hb->nch); return false; }
  hb->boundary.resize(hb->nch);
  int hmax = 0;
  std::vector<int> h(hb->nch);
  for (int i = 0; i < hb->nch; ++i) {
    hb->boundary[i] = afloor * pow(2.0, (i + 1) / channels_in_octave);                           // :1155-1157
    h[i] = matlab_round(actual_fs / hb->boundary[i] * 2.0);                                       // :101
    hmax = std::max(hmax, h[i]);
  }
  int bn = 1024;
  while (bn < 6 * hmax && bn < 8192) bn <<= 1;
  if (bn < 2 * hmax + 64) { set_error("Harvest: band-pass filters of %d taps do not fit the block FFT", 2 * hmax + 1); return false; }
  hb->bn = bn;
  hb->log2bn = 0; while ((1 << hb->log2bn) < bn) ++hb->log2bn;
  hb->D = hmax - 1;
  hb->V = bn - 2 * hmax;
  std::vector<double2> G((size_t)hb->nch * (bn / 2 + 1));
  std::vector<int> shift(hb->nch);
  for (int c = 0; c < hb->nch; ++c) {
    const int ln = 2 * h[c] + 1;
    std::vector<double> re(bn, 0.0), im(bn, 0.0);
    for (int i = 0; i < ln; ++i) {                       // NuttallWindow x cosine (:102-106)
      const double tmp = i / (ln - 1.0);
      const double nut = 0.355768 - 0.487396 * cos(2.0 * kPi * tmp) + 0.144232 * cos(4.0 * kPi * tmp) -
                         0.012604 * cos(6.0 * kPi * tmp);
      re[i] = nut * cos(2 * kPi * hb->boundary[c] * (i - h[c]) / actual_fs);
    }
    host_fft(re, im);
    for (int k = 0; k <= bn / 2; ++k) G[(size_t)c * (bn / 2 + 1) + k] = make_double2(re[k], im[k]);
    shift[c] = hb->D + h[c] + 1;                         // index_bias = filter_length_half + 1 (:140)
  }
  if (!hb->G.alloc(G.size()) || !hb->shift.alloc(hb->nch) || !hb->d_boundary.alloc(hb->nch)) return false;
  return WB_CUDA(cudaMemcpy(hb->G.p, G.data(), G.size() * sizeof(double2), cudaMemcpyHostToDevice)) &&
         WB_CUDA(cudaMemcpy(hb->shift.p, shift.data(), hb->nch * sizeof(int), cudaMemcpyHostToDevice)) &&
         WB_CUDA(cudaMemcpy(hb->d_boundary.p, hb->boundary.data(), hb->nch * sizeof(double), cudaMemcpyHostToDevice));
}

}  // namespace

// the filter table above for the host-side decimate() of wb_compat.cu
bool decimate_filter_coefficients(int r, double* a, double* b) { return decimate_coefficients(r, a, b); }

#ifndef WB_HOST_EMU      // the launchers; tests/emu has its own
// decimate() of W/src/matlabfunctions.cpp:184-210 for every utterance of the batch (no edge
// extension: lag = 0), as Dio uses it when option.speed > 1 (W/src/dio.cpp:69-71).  out_len[u] =
// min(want_len[u], number of values the reference's loop writes); the caller treats the rest as 0.
bool decimate_run(const Batch* b, int r, const std::vector<int>& want_len, DevBuf<double>* y,
                  DevBuf<long long>* y_off, DevBuf<int>* y_len, std::vector<int>* out_len) {
  Context* ctxp = ctx();
  if (!ctxp) return false;
  cudaStream_t st = ctxp->stream;
  const int n_utt = b->n_utt;
  HarvestConst c;
  c.fs = b->fs; c.r = r; c.nch = 0; c.lag = 0; c.actual_fs = (double)b->fs / r; c.f0_floor = c.f0_ceil = 0.0;
  if (!decimate_coefficients(r, c.a, c.b)) { set_error("decimate: unsupported ratio %d", r); return false; }
  std::vector<long long> h_yoff(n_utt), h_boff(n_utt);
  out_len->resize(n_utt);
  long long ytot = 0, btot = 0;
  int max_M = 0;
  for (int u = 0; u < n_utt; ++u) {
    const int L = b->h_x_len[u];
    if (L < 20) { set_error("decimate: utterance %d is too short (%d samples)", u, L); return false; }
    const int nout = (L - 1) / r + 1, nbeg = r - r * nout + L;
    const int written = (L + 9 - nbeg + r - 1) / r;                  // trips of the loop at :205-206
    (*out_len)[u] = std::min(want_len[u], written);
    h_yoff[u] = ytot; ytot += ((*out_len)[u] + 1) & ~1;
    h_boff[u] = btot; btot += (L + 18 + 1) & ~1;
    max_M = std::max(max_M, L + 18);
  }
  DevBuf<double> d_B;
  DevBuf<long long> d_boff;
  if (!y->alloc(ytot + 2) || !y_off->alloc(n_utt) || !y_len->alloc(n_utt) || !d_B.alloc(btot) || !d_boff.alloc(n_utt)) return false;
  if (!write_dev(y_off->p, h_yoff.data(), n_utt * sizeof(long long))) return false;
  if (!write_dev(y_len->p, out_len->data(), n_utt * sizeof(int))) return false;
  if (!write_dev(d_boff.p, h_boff.data(), n_utt * sizeof(long long))) return false;
  WB_CUDA_OR_RETURN(cudaStreamSynchronize(st), false);
  const int n_chunks = (max_M + kIirChunk - 1) / kIirChunk;
  harvest_iir_fwd_kernel<<<dim3((n_chunks + 63) / 64, n_utt), 64, 0, st>>>(b->x.p, b->x_off.p, b->x_len.p, d_boff.p, c, n_chunks, d_B.p);
  WB_LAUNCH_CHECK();
  harvest_iir_bwd_kernel<<<dim3((n_chunks + 63) / 64, n_utt), 64, 0, st>>>(d_B.p, d_boff.p, b->x_len.p, y_off->p, y_len->p, c, n_chunks, y->p);
  WB_LAUNCH_CHECK();
  WB_CUDA_OR_RETURN(cudaStreamSynchronize(st), false);                // d_B dies with this scope
  return true;
}

bool harvest_run(Batch* b, const HarvestParams& p, double* d_f0_out) {
  Context* ctxp = ctx();
  if (!ctxp) return false;
  cudaStream_t st = ctxp->stream;
  const int n_utt = b->n_utt;
  if (n_utt == 0) return true;
  HarvestConst c;
  c.fs = b->fs;
  c.r = std::max(std::min(matlab_round(b->fs / 8000.0), 12), 1);                                  // :1227, :1160
  c.actual_fs = (double)b->fs / c.r;
  c.f0_floor = p.f0_floor; c.f0_ceil = p.f0_ceil;
  c.lag = c.r == 1 ? 0 : static_cast<int>(ceil(140.0 / c.r) * c.r);                              // :49-50
  c.a[0] = c.a[1] = c.a[2] = c.b[0] = c.b[1] = 0.0;
  if (c.r > 1 && !decimate_coefficients(c.r, c.a, c.b)) { set_error("Harvest: unsupported decimation ratio %d", c.r); return false; }
  const std::vector<double> key = {c.actual_fs, p.f0_floor, p.f0_ceil};
  HarvestBank* hb = nullptr;
  auto it = g_hbanks.find(key);
  if (it == g_hbanks.end()) {
    hb = new HarvestBank();
    if (!build_harvest_bank(c.actual_fs, p.f0_floor, p.f0_ceil, hb)) { delete hb; return false; }
    g_hbanks[key] = hb;
  } else {
    hb = it->second;
  }
  c.nch = hb->nch;
  const int max_base = matlab_round(c.nch / 10.0);                                                // :1186-1187
  // WB_HARVEST_TRACE=1: wall time of every phase of this call (stream drained at each mark) on stderr
  static const bool trace = getenv("WB_HARVEST_TRACE") != nullptr;
  auto t_mark = std::chrono::steady_clock::now();
  auto phase = [&](const char* what) {
    if (!trace) return;
    cudaStreamSynchronize(st);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[harvest] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_mark).count());
    t_mark = now;
  };

  // ---- per-utterance sizes: decimated length, 1 ms frame grid -------------------------------------
  std::vector<int> h_ylen(n_utt), h_glen(n_utt), h_goff(n_utt);
  std::vector<long long> h_yoff(n_utt), h_boff(n_utt);
  long long ytot = 0, btot = 0, gtot = 0;
  int max_M = 0, max_y = 0, max_g = 0;
  for (int u = 0; u < n_utt; ++u) {
    const int xl = b->h_x_len[u];
    if (xl < 2 * 9 + 2) { set_error("Harvest: utterance %d is too short (%d samples)", u, xl); return false; }
    h_ylen[u] = static_cast<int>(ceil(static_cast<double>(xl) / c.r));                            // :1161-1162
    h_glen[u] = static_cast<int>(1000.0 * xl / b->fs / 1.0) + 1;                                  // GetSamplesForHarvest, 1 ms
    h_yoff[u] = ytot; ytot += (h_ylen[u] + 1) & ~1;
    const int M = xl + 2 * c.lag + 18;
    h_boff[u] = btot; btot += (M + 1) & ~1;
    h_goff[u] = (int)gtot; gtot += h_glen[u];
    max_M = std::max(max_M, M); max_y = std::max(max_y, h_ylen[u]); max_g = std::max(max_g, h_glen[u]);
  }
  if (gtot > 0x3fffffffLL) { set_error("Harvest: batch has too many 1 ms frames"); return false; }
  DevBuf<int> d_ylen, d_glen, d_goff, d_mask, d_nc;
  DevBuf<long long> d_yoff, d_boff;
  DevBuf<double> d_y, d_B, d_mean, d_base;
  if (!d_ylen.alloc(n_utt) || !d_glen.alloc(n_utt) || !d_goff.alloc(n_utt) || !d_mask.alloc(n_utt) || !d_nc.alloc(n_utt) ||
      !d_yoff.alloc(n_utt) || !d_boff.alloc(n_utt) || !d_y.alloc(ytot) || !d_mean.alloc(n_utt) ||
      !d_base.alloc((size_t)gtot * max_base))
    return false;
  std::vector<int> h_mask(n_utt, 0x3fffffff);
  auto up = [&](void* dst, const void* src, size_t bytes) { return write_dev(dst, src, bytes); };
  if (!up(d_ylen.p, h_ylen.data(), n_utt * sizeof(int)) || !up(d_glen.p, h_glen.data(), n_utt * sizeof(int)) ||
      !up(d_goff.p, h_goff.data(), n_utt * sizeof(int)) || !up(d_mask.p, h_mask.data(), n_utt * sizeof(int)) ||
      !up(d_yoff.p, h_yoff.data(), n_utt * sizeof(long long)) || !up(d_boff.p, h_boff.data(), n_utt * sizeof(long long)))
    return false;
  if (!dev_fill(d_nc.p, 0, n_utt * sizeof(int))) return false;
  WB_CUDA_OR_RETURN(cudaStreamSynchronize(st), false);

  phase("tables");
  // ---- 1. decimation -----------------------------------------------------------------------------------
  if (c.r > 1) {
    if (!d_B.alloc(btot)) return false;
    const int n_chunks = (max_M + kIirChunk - 1) / kIirChunk;
    KernelTimer kt("harvest_iir_kernel");
    harvest_iir_fwd_kernel<<<dim3((n_chunks + 63) / 64, n_utt), 64, 0, st>>>(b->x.p, b->x_off.p, b->x_len.p, d_boff.p, c, n_chunks, d_B.p);
    WB_LAUNCH_CHECK();
    harvest_iir_bwd_kernel<<<dim3((n_chunks + 63) / 64, n_utt), 64, 0, st>>>(d_B.p, d_boff.p, b->x_len.p, d_yoff.p, d_ylen.p, c, n_chunks, d_y.p);
    WB_LAUNCH_CHECK(); kt.stop();
  } else {
    harvest_copy_kernel<<<dim3(64, n_utt), 256, 0, st>>>(b->x.p, b->x_off.p, d_yoff.p, d_ylen.p, d_y.p);
    WB_LAUNCH_CHECK();
  }
  harvest_mean_kernel<<<n_utt, 256, 0, st>>>(d_y.p, d_yoff.p, d_ylen.p, d_mean.p);
  WB_LAUNCH_CHECK();

  phase("decimation + mean");
  // ---- 2./3. band filtering, zero crossings, raw and base candidates, in sub-batches -------------------
  OlsConst oc = {hb->nch, hb->bn, hb->log2bn, hb->D, hb->V};
  const size_t smem = 2 * cpad_size(hb->bn / 2) * sizeof(double2);
  WB_CUDA_OR_RETURN(cudaFuncSetAttribute(ols_filter_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), false);
  WB_CUDA_OR_RETURN(cudaFuncSetAttribute(ols_filter_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), false);
  WB_CUDA_OR_RETURN(cudaFuncSetAttribute(ols_filter_zc_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), false);
  WB_CUDA_OR_RETURN(cudaFuncSetAttribute(ols_filter_zc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), false);
  const size_t kMaxScratchDoubles = (size_t)2 << 30;           // 16 GiB of filtered signals at a time
  const size_t kMaxSegDoubles = (size_t)1 << 30;               // 8 GiB of zero-crossing segments at a time
  constexpr int kSegCap = 256;                                 // events per (list, block): a block is ~1 550 samples at 8 kHz (0.19 s), the
                                                               // highest channel (880 Hz) yields ~170 events of each type in it
  const bool fused_wanted = option("harvest_fused") && hb->V > 64;
  // scratch of the sub-batches: declared once, grow-only (DevBuf::alloc keeps a buffer that is large enough), so
  // the sub-batches of a call -- and, through the pool, of the next call -- reuse the same blocks
  DevBuf<double> d_F, d_edges, d_raw, d_seg;
  DevBuf<long long> d_foff, d_loff, d_roff;
  DevBuf<int> d_counts, d_ltot, d_segcnt, d_segoff;
  int u0 = 0;
  while (u0 < n_utt) {
    int u1 = u0;
    size_t tot = 0, rtot = 0;
    int sub_max_y = 0, sub_max_g = 0;
    std::vector<long long> h_foff, h_roff;
    while (u1 < n_utt) {
      const size_t need = (size_t)c.nch * h_ylen[u1] + 8;
      if (u1 > u0 && (tot + need > kMaxScratchDoubles || (long long)(u1 - u0 + 1) * c.nch > 65535)) break;   // grid.y of the per-(utterance, channel) kernels
      if (u1 > u0 && fused_wanted) {
        // the fused path's event segments: (lists) x (blocks of the longest utterance) x kSegCap doubles; bounded so
        // that the scratch pool settles on a few GiB whatever the batch (a 1 132-utterance batch asked for 22 GB
        // per sub-batch and the pool kept growing for several calls)
        const int my = std::max(sub_max_y, h_ylen[u1]);
        const size_t nb = (size_t)(my - 1 + (hb->V - 2) - 1) / (hb->V - 2);
        if ((size_t)(u1 - u0 + 1) * c.nch * 4 * nb * kSegCap > kMaxSegDoubles) break;
      }
      h_foff.push_back((long long)tot);
      h_roff.push_back((long long)rtot);
      tot += need;
      rtot += (size_t)c.nch * h_glen[u1];
      sub_max_y = std::max(sub_max_y, h_ylen[u1]);
      sub_max_g = std::max(sub_max_g, h_glen[u1]);
      ++u1;
    }
    const int nu = u1 - u0;
    const int n_lists = nu * c.nch * 4;
    if (!d_roff.alloc(nu) || !d_ltot.alloc(n_lists + 1) || !d_loff.alloc(n_lists) || !d_raw.alloc(rtot)) return false;
    if (!up(d_roff.p, h_roff.data(), nu * sizeof(long long))) return false;
    phase("sub-batch: allocations");
    std::vector<int> h_ltot(n_lists + 1);
    std::vector<long long> h_loff(n_lists);
    bool fused_done = false;
    if (fused_wanted) {
      // band-pass filters + zero crossings in one kernel: the 152 channel signals never leave shared memory
      constexpr int kCap = kSegCap;
      OlsConst ocz = oc;
      ocz.V = hb->V - 2;
      const int n_blocks = (sub_max_y - 1 + ocz.V - 1) / ocz.V;
      if (!d_segcnt.alloc((size_t)n_lists * n_blocks) || !d_segoff.alloc((size_t)n_lists * n_blocks) ||
          !d_seg.alloc((size_t)n_lists * n_blocks * kCap))
        return false;
      if (!dev_fill(d_segcnt.p, 0, (size_t)n_lists * n_blocks * sizeof(int))) return false;
      if (!dev_fill(d_ltot.p + n_lists, 0, sizeof(int))) return false;
      phase("sub-batch: segment buffers");
      {
        KernelTimer kt("harvest_filter_kernel");
        if (hb->log2bn == 11)
          ols_filter_zc_kernel<11><<<dim3(n_blocks, nu), 256, smem, st>>>(d_y.p, d_yoff.p, d_ylen.p, d_ylen.p, d_mask.p, d_mean.p, hb->G.p,
                                                                         ctxp->tw_c(11), ocz, hb->shift.p, u0, n_blocks, kCap, d_segcnt.p, d_seg.p);
        else
          ols_filter_zc_kernel<0><<<dim3(n_blocks, nu), 256, smem, st>>>(d_y.p, d_yoff.p, d_ylen.p, d_ylen.p, d_mask.p, d_mean.p, hb->G.p,
                                                                        ctxp->d_twiddle, ocz, hb->shift.p, u0, n_blocks, kCap, d_segcnt.p, d_seg.p);
        WB_LAUNCH_CHECK(); kt.stop();
      }
      phase("sub-batch: filter + zc");
      zc_seg_scan_kernel<<<(n_lists + 127) / 128, 128, 0, st>>>(d_segcnt.p, n_lists, n_blocks, kCap, d_segoff.p, d_ltot.p, d_ltot.p + n_lists);
      WB_LAUNCH_CHECK();
      if (!read_back(h_ltot.data(), d_ltot.p, (n_lists + 1) * sizeof(int))) return false;
      phase("sub-batch: seg scan + read");
      if (h_ltot[n_lists] == 0) {
        long long etot = 0;
        for (int l = 0; l < n_lists; ++l) { h_loff[l] = etot; etot += h_ltot[l]; }
        if (!d_edges.alloc((size_t)etot + 2)) return false;
        if (!up(d_loff.p, h_loff.data(), n_lists * sizeof(long long))) return false;
        KernelTimer kt("harvest_zc_kernel");
        zc_seg_gather_kernel<<<n_lists, 128, 0, st>>>(d_segcnt.p, d_segoff.p, d_seg.p, n_blocks, kCap, d_loff.p, d_edges.p);
        WB_LAUNCH_CHECK(); kt.stop();
        WB_CUDA_OR_RETURN(cudaStreamSynchronize(st), false);  // the segment buffers die with this scope
        phase("sub-batch: gather");
        fused_done = true;
      }
    }
    if (!fused_done) {
      const int n_blocks = (sub_max_y + hb->V - 1) / hb->V;
      const int n_chunks = (sub_max_y + kZcChunk - 1) / kZcChunk;
      if (!d_F.alloc(tot) || !d_foff.alloc(nu) || !d_counts.alloc((size_t)n_lists * n_chunks)) return false;
      if (!up(d_foff.p, h_foff.data(), nu * sizeof(long long))) return false;
      {
        KernelTimer kt("harvest_filter_kernel");
        // the decimated signals play the role of Dio's x: offsets d_yoff, length = y_len for both "x" and "y"
        if (hb->log2bn == 11)
          ols_filter_kernel<11><<<dim3(n_blocks, nu), 256, smem, st>>>(d_y.p, d_yoff.p, d_ylen.p, d_ylen.p, d_mask.p, d_mean.p, d_foff.p,
                                                                      hb->G.p, ctxp->tw_c(11), oc, hb->shift.p, u0, d_F.p);
        else
          ols_filter_kernel<0><<<dim3(n_blocks, nu), 256, smem, st>>>(d_y.p, d_yoff.p, d_ylen.p, d_ylen.p, d_mask.p, d_mean.p, d_foff.p,
                                                                     hb->G.p, ctxp->d_twiddle, oc, hb->shift.p, u0, d_F.p);
        WB_LAUNCH_CHECK(); kt.stop();
      }
      {
        KernelTimer kt("harvest_zc_kernel");
        zc_kernel<false><<<dim3(n_chunks, nu * c.nch), 256, 0, st>>>(d_F.p, d_foff.p, d_ylen.p, c.nch, u0, n_chunks, d_counts.p, nullptr, nullptr);
        WB_LAUNCH_CHECK();
        zc_scan_kernel<<<(n_lists + 127) / 128, 128, 0, st>>>(d_counts.p, n_lists, n_chunks, d_ltot.p);
        WB_LAUNCH_CHECK(); kt.stop();
      }
      if (!read_back(h_ltot.data(), d_ltot.p, n_lists * sizeof(int))) return false;
      long long etot = 0;
      for (int l = 0; l < n_lists; ++l) { h_loff[l] = etot; etot += h_ltot[l]; }
      if (!d_edges.alloc((size_t)etot + 2)) return false;
      if (!up(d_loff.p, h_loff.data(), n_lists * sizeof(long long))) return false;
      {
        KernelTimer kt("harvest_zc_kernel");
        zc_kernel<true><<<dim3(n_chunks, nu * c.nch), 256, 0, st>>>(d_F.p, d_foff.p, d_ylen.p, c.nch, u0, n_chunks, d_counts.p, d_loff.p, d_edges.p);
        WB_LAUNCH_CHECK(); kt.stop();
      }
    }
    KernelTimer ktr("harvest_raw_kernel");
    harvest_raw_kernel<<<dim3((sub_max_g + 127) / 128, nu * c.nch), 128, 0, st>>>(d_edges.p, d_loff.p, d_ltot.p, d_goff.p, d_glen.p,
                                                                                hb->d_boundary.p, c, u0, d_roff.p, d_raw.p);
    WB_LAUNCH_CHECK();
    harvest_detect_kernel<<<dim3((sub_max_g + 127) / 128, nu), 128, 0, st>>>(d_raw.p, d_roff.p, d_goff.p, d_glen.p, c.nch, max_base, u0,
                                                                           d_base.p, d_nc.p);
    WB_LAUNCH_CHECK(); ktr.stop();
    WB_CUDA_OR_RETURN(cudaStreamSynchronize(st), false);      // scratch buffers die with this scope
    phase("sub-batch: raw + detect");
    u0 = u1;
  }

  // ---- 4. refinement of the overlapped candidates --------------------------------------------------------
  std::vector<int> h_nc(n_utt);
  if (!read_back(h_nc.data(), d_nc.p, n_utt * sizeof(int))) return false;
  std::vector<long long> h_coff(n_utt), h_wfirst(n_utt);
  long long ctot = 0;
  for (int u = 0; u < n_utt; ++u) {
    h_coff[u] = ctot; h_wfirst[u] = ctot;
    ctot += (long long)h_glen[u] * h_nc[u] * kOverlap;
  }
  DevBuf<double> d_tmp1, d_tmp2;
  DevBuf<int> d_bl, d_nsec;
  if (!d_tmp1.alloc(gtot) || !d_tmp2.alloc(gtot) || !d_bl.alloc(2 * (gtot + 2LL * kSmoothLag * n_utt) + 8) || !d_nsec.alloc(n_utt)) return false;
  DevBuf<long long> d_coff, d_wfirst;
  DevBuf<double> d_cand, d_score, d_cand2, d_score2;
  if (!d_coff.alloc(n_utt) || !d_wfirst.alloc(n_utt) || !d_cand.alloc(ctot + 1) || !d_score.alloc(ctot + 1) ||
      !d_cand2.alloc(ctot + 1) || !d_score2.alloc(ctot + 1))
    return false;
  if (!up(d_coff.p, h_coff.data(), n_utt * sizeof(long long)) || !up(d_wfirst.p, h_wfirst.data(), n_utt * sizeof(long long))) return false;
  phase("candidate buffers");
  if (ctot > 0) {
    {
      KernelTimer kt("harvest_refine_kernel");
      if (option("harvest_refine_thread")) {
        harvest_refine_thread_kernel<<<(unsigned)((ctot / kOverlap + 127) / 128), 128, 0, st>>>(d_y.p, d_yoff.p, d_ylen.p, d_mean.p, d_base.p, d_goff.p, d_glen.p,
                                                                                    d_nc.p, d_coff.p, max_base, c, n_utt, ctot / kOverlap, ctxp->d_twiddle_c, d_cand.p, d_score.p);
      } else {
        const long long threads = gtot * 32;
        harvest_refine_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_y.p, d_yoff.p, d_ylen.p, d_mean.p, d_base.p, d_goff.p, d_glen.p,
                                                                                d_nc.p, d_coff.p, max_base, c, n_utt, gtot, ctxp->d_twiddle_c, d_cand.p, d_score.p);
      }
      WB_LAUNCH_CHECK(); kt.stop();
    }
    KernelTimer ktu("harvest_unreliable_kernel");
    harvest_unreliable_kernel<<<(unsigned)((ctot + 255) / 256), 256, 0, st>>>(d_cand.p, d_score.p, d_glen.p, d_nc.p, d_coff.p, n_utt, d_wfirst.p,
                                                                             ctot, d_cand2.p, d_score2.p);
    WB_LAUNCH_CHECK(); ktu.stop();
  }
  // ---- 5. contour logic -------------------------------------------------------------------------------------
  // bl_all is indexed by 2 * g_off[u]; the smoothing stage needs 2 * (g_len + 600) entries at most, and
  // sections are at least 1 frame apart, so 2 * g_off[u] + ... stays inside because g_len >= 602 is not
  // guaranteed: use a separate, padded offset table for the smoothing boundaries below.
  phase("refine + unreliable");
  KernelTimer kta("harvest_fix_a_kernel");
  harvest_fix_a_kernel<<<n_utt, 256, 0, st>>>(d_cand2.p, d_score2.p, d_goff.p, d_glen.p, d_nc.p, d_coff.p, d_tmp1.p, d_tmp2.p, d_bl.p, d_nsec.p);
  WB_LAUNCH_CHECK(); kta.stop();
  std::vector<int> h_nsec(n_utt);
  if (!read_back(h_nsec.data(), d_nsec.p, n_utt * sizeof(int))) return false;
  {
    std::vector<long long> h_mcoff(n_utt);
    long long mtot = 0;
    for (int u = 0; u < n_utt; ++u) { h_mcoff[u] = mtot; mtot += (long long)h_nsec[u] * h_glen[u]; }
    DevBuf<long long> d_mcoff;
    DevBuf<double> d_mc;
    DevBuf<int> d_chan, d_order;
    if (!d_mcoff.alloc(n_utt) || !d_mc.alloc(mtot + 1) || !d_chan.alloc(gtot) || !d_order.alloc(gtot)) return false;
    if (!up(d_mcoff.p, h_mcoff.data(), n_utt * sizeof(long long))) return false;
    KernelTimer kt("harvest_fix_kernel");
    if (getenv("WB_HARVEST_FIX_CHECK")) {
      // debugging aid: both versions of the contour logic on copies of the same inputs, differences reported
      DevBuf<double> t1b, t2b, mcb;
      DevBuf<int> blb, chb, orb;
      if (!t1b.alloc(gtot) || !t2b.alloc(gtot) || !mcb.alloc(mtot + 1) || !blb.alloc(d_bl.n) || !chb.alloc(gtot) || !orb.alloc(gtot)) return false;
      cudaMemcpyAsync(t1b.p, d_tmp1.p, gtot * sizeof(double), cudaMemcpyDeviceToDevice, st);
      cudaMemcpyAsync(t2b.p, d_tmp2.p, gtot * sizeof(double), cudaMemcpyDeviceToDevice, st);
      cudaMemcpyAsync(blb.p, d_bl.p, d_bl.n * sizeof(int), cudaMemcpyDeviceToDevice, st);
      const int stop = getenv("WB_HARVEST_FIX_STOP") ? atoi(getenv("WB_HARVEST_FIX_STOP")) : 0;
      harvest_fix_b_serial_kernel<<<n_utt, 32, 0, st>>>(d_cand2.p, d_score2.p, d_goff.p, d_glen.p, d_nc.p, d_coff.p, t1b.p, t2b.p, blb.p,
                                                       d_mcoff.p, mcb.p, chb.p, orb.p, stop);
      harvest_fix_b_kernel<<<n_utt, 32, 0, st>>>(d_cand2.p, d_score2.p, d_goff.p, d_glen.p, d_nc.p, d_coff.p, d_tmp1.p, d_tmp2.p, d_bl.p,
                                                d_mcoff.p, d_mc.p, d_chan.p, d_order.p, atoi(getenv("WB_HARVEST_FIX_CHECK")), stop);
      if (stop) {
        std::vector<double> ma(mtot + 1), mb(mtot + 1);
        std::vector<int> ba(d_bl.n), bb(d_bl.n), ca(gtot), cb(gtot);
        cudaMemcpyAsync(ma.data(), mcb.p, mtot * sizeof(double), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(mb.data(), d_mc.p, mtot * sizeof(double), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(ba.data(), blb.p, d_bl.n * sizeof(int), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(bb.data(), d_bl.p, d_bl.n * sizeof(int), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(ca.data(), chb.p, gtot * sizeof(int), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(cb.data(), d_chan.p, gtot * sizeof(int), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        long long dm = 0, db = 0, dc = 0;
        const int nsec0 = h_nsec[0], n0 = h_glen[0];
        for (long long i = 0; i < (long long)nsec0 * n0; ++i)
          if (memcmp(&ma[i], &mb[i], 8) != 0 && dm++ < 6) fprintf(stderr, "[fix check] mc section %lld frame %lld: serial %.9g warp %.9g\n", i / n0, i % n0, ma[i], mb[i]);
        for (int i = 0; i < 2 * nsec0; ++i) if (ba[i] != bb[i] && db++ < 6) fprintf(stderr, "[fix check] bl[%d]: serial %d warp %d\n", i, ba[i], bb[i]);
        for (int i = 0; i < nsec0; ++i) if (ca[i] != cb[i] && dc++ < 6) fprintf(stderr, "[fix check] chan[%d]: serial %d warp %d\n", i, ca[i], cb[i]);
        fprintf(stderr, "[fix check] stop %d, utterance 0: %d sections, %d frames: mc differs at %lld, bl at %lld, chan at %lld\n", stop, nsec0, n0, dm, db, dc);
      }
      std::vector<double> ha(gtot), hb2(gtot), h3a(gtot), h3b(gtot);
      cudaMemcpyAsync(ha.data(), t1b.p, gtot * sizeof(double), cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(hb2.data(), d_tmp1.p, gtot * sizeof(double), cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(h3a.data(), t2b.p, gtot * sizeof(double), cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(h3b.data(), d_tmp2.p, gtot * sizeof(double), cudaMemcpyDeviceToHost, st);
      cudaStreamSynchronize(st);
      long long nd4 = 0, nd3 = 0;
      for (long long i = 0; i < gtot; ++i) {
        if (memcmp(&h3a[i], &h3b[i], 8) != 0 && nd3++ < 5) fprintf(stderr, "[fix check] step3 frame %lld: serial %.9g warp %.9g\n", i, h3a[i], h3b[i]);
        if (memcmp(&ha[i], &hb2[i], 8) != 0 && nd4++ < 5) fprintf(stderr, "[fix check] step4 frame %lld: serial %.9g warp %.9g\n", i, ha[i], hb2[i]);
      }
      fprintf(stderr, "[fix check] %lld frames: step3 differs at %lld, step4 at %lld\n", gtot, nd3, nd4);
    } else if (option("harvest_fix_warp")) {
      harvest_fix_b_kernel<<<n_utt, 32, 0, st>>>(d_cand2.p, d_score2.p, d_goff.p, d_glen.p, d_nc.p, d_coff.p, d_tmp1.p, d_tmp2.p, d_bl.p,
                                                d_mcoff.p, d_mc.p, d_chan.p, d_order.p);
    } else {
      harvest_fix_b_serial_kernel<<<n_utt, 32, 0, st>>>(d_cand2.p, d_score2.p, d_goff.p, d_glen.p, d_nc.p, d_coff.p, d_tmp1.p, d_tmp2.p, d_bl.p,
                                                       d_mcoff.p, d_mc.p, d_chan.p, d_order.p);
    }
    WB_LAUNCH_CHECK(); kt.stop();
    WB_CUDA_OR_RETURN(cudaStreamSynchronize(st), false);
  }
  phase("contour logic");
  // smoothing: step4 is in tmp1; the smoothed basic contour goes to tmp2 (zero where unvoiced)
  if (!dev_fill(d_tmp2.p, 0, (size_t)gtot * sizeof(double))) return false;
  harvest_sections_kernel<<<n_utt, 32, 0, st>>>(d_tmp1.p, d_goff.p, d_glen.p, d_bl.p, d_nsec.p);
  WB_LAUNCH_CHECK();
  if (!read_back(h_nsec.data(), d_nsec.p, n_utt * sizeof(int))) return false;
  {
    std::vector<int> h_sfirst(n_utt);
    std::vector<long long> h_soff(n_utt);
    int stot = 0;
    long long scr = 0;
    for (int u = 0; u < n_utt; ++u) {
      h_sfirst[u] = stot; h_soff[u] = scr;
      stot += h_nsec[u];
      scr += (long long)h_nsec[u] * (h_glen[u] + 2 * kSmoothLag);
    }
    if (stot > 0) {
      DevBuf<int> d_sfirst;
      DevBuf<long long> d_soff;
      DevBuf<double> d_scr;
      if (!d_sfirst.alloc(n_utt) || !d_soff.alloc(n_utt) || !d_scr.alloc(scr + 1)) return false;
      if (!up(d_sfirst.p, h_sfirst.data(), n_utt * sizeof(int)) || !up(d_soff.p, h_soff.data(), n_utt * sizeof(long long))) return false;
      KernelTimer kt("harvest_smooth_kernel");
      harvest_smooth_kernel<<<(stot + 63) / 64, 64, 0, st>>>(d_tmp1.p, d_goff.p, d_glen.p, d_bl.p, d_sfirst.p, n_utt, stot, d_soff.p, d_scr.p, d_tmp2.p);
      WB_LAUNCH_CHECK(); kt.stop();
      WB_CUDA_OR_RETURN(cudaStreamSynchronize(st), false);
    }
  }
  int max_f = 0;
  for (int u = 0; u < n_utt; ++u) max_f = std::max(max_f, b->h_f_len[u]);
  if (max_f > 0) {
    harvest_pick_kernel<<<dim3((max_f + 127) / 128, n_utt), 128, 0, st>>>(d_tmp2.p, d_goff.p, d_glen.p, b->f_off.p, b->f_len.p, b->frame_period, d_f0_out);
    WB_LAUNCH_CHECK();
  }
  WB_CUDA_OR_RETURN(cudaStreamSynchronize(st), false);
  phase("smoothing + pick");
  return true;
}

#endif  // WB_HOST_EMU

}  // namespace wb
