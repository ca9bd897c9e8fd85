// world-b200: zero-crossing machinery shared by Dio and Harvest (both estimate F0 from the
// intervals between four kinds of zero-crossing events of band-limited copies of the signal).
// Reference: W/src/dio.cpp ZeroCrossingEngine :357-393, GetFourZeroCrossingIntervals :402-435;
// W/src/harvest.cpp :162-238 (identical logic); W/src/matlabfunctions.cpp interp1 :157-182.
#pragma once
#include <math.h>
#include <algorithm>
#include <vector>
#include "wb_common.cuh"
#include "wb_fft.cuh"

namespace wb {

constexpr int kZcChunk = 2048;       // sample pairs per CTA in the zero-crossing kernels

// ---- host: build the combined filters and their spectra ------------------------------------------
inline void host_fft(std::vector<double>& re, std::vector<double>& im) {   // in-place radix-2, forward
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

// ---- zero crossings ------------------------------------------------------------------------------
// Event types on the filtered signal s (GetFourZeroCrossingIntervals :402-435):
//   0 negative-going  s[i] > 0 >= s[i+1]            1 positive-going  -s
//   2 peaks           d[i] = s[i+1] - s[i], d[i] > 0 >= d[i+1]      3 dips  -d
__device__ __forceinline__ bool zc_event(double a, double b) { return 0.0 < a && b <= 0.0; }
__device__ __forceinline__ double zc_fine(int e, double a, double b) {   // :376-379
  return add_rn((double)e, -div_rn(a, add_rn(b, -a)));
}

// One CTA scans kZcChunk sample pairs of one (utterance, band).  Events keep the order of the
// sample index: thread t looks at samples i0 + k * 256 + t (k = 0..7, coalesced), all eight
// rounds are evaluated first, the per-(round, type, warp) event counts go to shared memory and
// ONE barrier later every event knows its slot (prefix over rounds and warps + ballot prefix
// inside the warp).  WRITE = false only counts.
template <bool WRITE>
static __global__ void __launch_bounds__(256)
zc_kernel(const double* __restrict__ F, const long long* __restrict__ F_off,
              const int* __restrict__ y_len_all, int nb, int utt0, int n_chunks_max,
              int* __restrict__ counts,              // [lists][n_chunks_max] (count pass: out; write: exclusive offsets)
              const long long* __restrict__ list_off, double* __restrict__ edges) {
  constexpr int kRounds = kZcChunk / 256;
  __shared__ int wcnt[kRounds][4][8];
  __shared__ unsigned lpre_s[WRITE ? kRounds : 1][256];
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
  unsigned evbits = 0;                               // bit (4 k + t): event of type t in round k
#pragma unroll
  for (int k = 0; k < kRounds; ++k) {
    const int i = i0 + k * 256 + tid;
    bool ev[4] = {false, false, false, false};
    if (i < y_len - 1) {
      const double s0 = s[i], s1 = s[i + 1];
      ev[0] = zc_event(s0, s1);
      ev[1] = zc_event(-s0, -s1);
      if (i < y_len - 2) {
        const double s2 = s[i + 2];
        const double d0 = add_rn(s1, -s0), d1 = add_rn(s2, -s1);
        ev[2] = zc_event(d0, d1);
        ev[3] = zc_event(-d0, -d1);
      }
    }
    unsigned pre = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const unsigned bal = __ballot_sync(0xffffffffu, ev[t]);
      if (lane == 0) wcnt[k][t][wid] = __popc(bal);
      if (ev[t]) evbits |= 1u << (4 * k + t);
      pre |= (unsigned)__popc(bal & ((1u << lane) - 1u)) << (8 * t);   // events of lower lanes (<= 31)
    }
    if (WRITE) lpre_s[k][tid] = pre;
  }
  __syncthreads();
  {
    // exclusive prefix over (round, warp) of every type, in place: the 4 x 64 counters are one per
    // thread (type = tid / 64, entry = round * 8 + warp), scanned with shuffles inside each warp and
    // joined across the two warps of a type -- four threads walking 64 entries each kept the other
    // 252 waiting for ~2000 cycles per CTA
    static_assert(kRounds * 8 == 64, "one counter per thread of a 256-thread CTA");
    __shared__ int half_tot[4];
    const int ty = tid >> 6, idx = tid & 63;
    const int v = wcnt[idx >> 3][ty][idx & 7];
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (idx == 31) half_tot[ty] = inc;
    __syncthreads();
    const int first_half = half_tot[ty];
    const int base = WRITE ? counts[((size_t)ub * 4 + ty) * n_chunks_max + chunk] : 0;
    wcnt[idx >> 3][ty][idx & 7] = base + (inc - v) + (idx >= 32 ? first_half : 0);
    if (!WRITE && idx == 63) counts[((size_t)ub * 4 + ty) * n_chunks_max + chunk] = inc + first_half;
  }
  if (!WRITE) return;
  __syncthreads();
  if (evbits == 0) return;
#pragma unroll 1
  for (int k = 0; k < kRounds; ++k) {
    const unsigned e4 = (evbits >> (4 * k)) & 0xfu;
    if (e4 == 0) continue;
    const int i = i0 + k * 256 + tid;
    const double s0 = s[i], s1 = s[i + 1];
    const double s2 = i < y_len - 2 ? s[i + 2] : 0.0;
    const double d0 = add_rn(s1, -s0), d1 = add_rn(s2, -s1);
    const unsigned pre = lpre_s[k][tid];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (!((e4 >> t) & 1u)) continue;
      const int pos = wcnt[k][t][wid] + (int)((pre >> (8 * t)) & 0xffu);
      const double fine = t == 0 ? zc_fine(i + 1, s0, s1) : t == 1 ? zc_fine(i + 1, -s0, -s1)
                        : t == 2 ? zc_fine(i + 1, d0, d1) : zc_fine(i + 1, -d0, -d1);
      edges[list_off[(size_t)ub * 4 + t] + pos] = fine;
    }
  }
}

// one thread per list: exclusive scan over chunks (in place), list totals out
static __global__ void zc_scan_kernel(int* __restrict__ counts, int n_lists, int n_chunks_max,
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
  return div_rn(mul_rn(add_rn(e[i], e[i + 1]), 0.5), fs);       // halving is exact
}
__device__ __forceinline__ double zc_val(const double* __restrict__ e, int i, double fs) {
  return div_rn(fs, add_rn(e[i + 1], -e[i]));
}
// zc_loc(e, i, fs) <= t without the two divisions in all but borderline cases: with s = (e_i + e_{i+1}) / 2
// (exact halving) the knot is rn(s / fs), and rounding is monotonic: s / fs <= t implies rn(s / fs) <= t;
// s / fs > t allows rn(s / fs) == t only within half an ulp of t.  The sign of fma(t, fs, -s) is the exact
// sign of t fs - s, so only |t fs - s| <= fs ulp(t) needs the reference's own arithmetic.
__device__ __forceinline__ bool zc_loc_le(const double* __restrict__ e, int i, double fs, double t) {
  const double s = mul_rn(add_rn(e[i], e[i + 1]), 0.5);
  const double r = fma(t, fs, -s);
  if (r >= 0.0) return true;
  if (-r > fs * t * 4.5e-16) return false;
  return div_rn(s, fs) <= t;
}
// first knot of e[0 .. n_int] whose location exceeds t (upper_bound), searched in [lo, hi)
__device__ __forceinline__ int zc_upper_bound(const double* __restrict__ e, int lo, int hi, double fs, double t) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (zc_loc_le(e, mid, fs, t)) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// interp1 at t when the upper bound is known to lie in [lo, hi] and the edges e[g0 ..] sit in `win`
// (shared memory): the same arithmetic as zc_interp on the same values.
__device__ __forceinline__ double zc_interp_window(const double* win, int g0, int lo, int hi, int n_int, double fs, double t) {
  const double* e = win - g0;                 // e[i] for i in the window
  const int ub = zc_upper_bound(e, lo, hi, fs, t);
  const int k = max(1, min(n_int - 1, ub));
  const double x0 = zc_loc(e, k - 1, fs), x1 = zc_loc(e, k, fs);
  const double y0 = zc_val(e, k - 1, fs), y1 = zc_val(e, k, fs);
  const double sfrac = div_rn(add_rn(t, -x0), add_rn(x1, -x0));
  return add_rn(y0, mul_rn(sfrac, add_rn(y1, -y0)));
}
__device__ inline double zc_interp(const double* __restrict__ e, int n_int, double fs, double t) {
  int lo = 0, hi = n_int;                 // upper_bound: first knot with loc > t
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (zc_loc_le(e, mid, fs, t)) lo = mid + 1; else hi = mid;
  }
  const int k = max(1, min(n_int - 1, lo));
  const double x0 = zc_loc(e, k - 1, fs), x1 = zc_loc(e, k, fs);
  const double y0 = zc_val(e, k - 1, fs), y1 = zc_val(e, k, fs);
  const double sfrac = div_rn(add_rn(t, -x0), add_rn(x1, -x0));
  return add_rn(y0, mul_rn(sfrac, add_rn(y1, -y0)));
}


// ---- band filtering by overlap-save -------------------------------------------------------------
// Both F0 estimators filter the whole utterance with a bank of FIR filters (Dio: low-cut x
// Nuttall low-pass, 7 bands, W/src/dio.cpp:296-343; Harvest: Nuttall x cosine band-pass, ~150
// channels, W/src/harvest.cpp:99-148) through FFTs of 2^15 - 2^19 points.  Here the same
// convolution is evaluated block by block with transforms that live in shared memory: one
// forward transform of a bn-sample block, then per band a multiply by the precomputed filter
// spectrum G[b] and one inverse transform.  Input index m = (n0 - D + i) & mask reproduces the
// reference's circular wrap-around when its FFT is short (Dio); a mask of
// 2^30 - 1 turns the wrap off (negative indices become huge and read as zero padding).
struct OlsConst { int nb, bn, log2bn, D, V; };

// dynamic shared memory: [ xs: cpad_size(bn/2) double2 | ws: cpad_size(bn/2) double2 ]
template <int LOG2BN, int THREADS = 256, int MAXK = 4>     // LOG2BN 0: block size given at run time (c.log2bn)
static __global__ void __launch_bounds__(THREADS)
ols_filter_kernel(const double* __restrict__ x_all, const long long* __restrict__ x_off,
                  const int* __restrict__ x_len, const int* __restrict__ y_len_all,
                  const int* __restrict__ fft_mask_all, const double* __restrict__ mean_all,
                  const long long* __restrict__ F_off, const double2* __restrict__ G,
                  const double2* __restrict__ tw, OlsConst c, const int* __restrict__ shift_all, int utt0,
                  double* __restrict__ F) {
  WB_DYN_SMEM(double2, smem2);
  const int u = utt0 + blockIdx.y;
  const int y_len = y_len_all[u];
  const int n0 = blockIdx.x * c.V;
  if (n0 >= y_len) return;
  constexpr int LM = LOG2BN > 0 ? LOG2BN - 1 : 0;
  constexpr int TWL = LOG2BN > 0 ? LOG2BN : kTwLog2;     // compact twiddle table of this size, or the master table
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
  fft_dit<LM, false, THREADS, MAXK, TWL>(xs, log2m, tw);
  // half spectrum in place: slot k = X[k] (k < M), slot 0 = (X[0], X[M])
  for (int k = tid; k <= M / 2; k += T) {
    if (k == 0) {
      const double2 z0 = xs[0];
      xs[0] = make_double2(z0.x + z0.y, z0.x - z0.y);
    } else {
      const double2 a = rfft_bin<TWL>(xs, log2m, k, tw);
      const double2 b = rfft_bin<TWL>(xs, log2m, M - k, tw);
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
        ws[cpad(brev(0, log2m))] = c2r_pack<TWL>(y0, yM, 0, log2m, tw);
      } else {
        const double2 yk = cmul(xs[cpad(k)], Gb[k]);
        const double2 ym = cmul(xs[cpad(M - k)], Gb[M - k]);
        ws[cpad(brev(k, log2m))] = c2r_pack<TWL>(yk, ym, k, log2m, tw);
        if (k != M - k) ws[cpad(brev(M - k, log2m))] = c2r_pack<TWL>(ym, yk, M - k, log2m, tw);
      }
    }
    fft_dit<LM, true, THREADS, MAXK, TWL>(ws, log2m, tw);
    const int shift = shift_all[b];                    // filtered_b[n] = conv[n - n0 + shift]
    double* __restrict__ dst = Fu + (size_t)b * y_len + n0;
    for (int i = tid; i < n_out; i += T) dst[i] = wsd[rfft_out_slot(i + shift)];
    __syncthreads();
  }
}


// ---- band filtering with the zero crossings taken from the block while it is still in shared memory --
// ols_filter_kernel writes every band signal to HBM (7 x the input for Dio, 152 x for Harvest) and the two
// zc_kernel passes read all of it back twice.  Here the events of a block are detected right behind the
// inverse transform of each band: thread t owns `per` consecutive output samples (plus a two-sample
// halo, so a block owns V - 2 outputs instead of V), marks the four event types in bit masks, one block
// scan of the packed counts orders them, and the sub-sample positions go to this block's segment of a
// staging buffer -- seg_cap events per (list, block); a block that finds more only counts them and the
// caller falls back to the two-pass path for that sub-batch.  A scan over the blocks of each list and a
// gather then give the same contiguous, ordered edge lists as zc_kernel<true>.
constexpr int kZcSegCap = 256;      // Dio's capacity (8 192-sample blocks at the full rate); Harvest passes 128

template <int LOG2BN, int THREADS = 256, int MAXK = 4>     // LOG2BN 0: block size given at run time (c.log2bn)
static __global__ void __launch_bounds__(THREADS)
ols_filter_zc_kernel(const double* __restrict__ x_all, const long long* __restrict__ x_off,
                     const int* __restrict__ x_len, const int* __restrict__ y_len_all,
                     const int* __restrict__ fft_mask_all, const double* __restrict__ mean_all,
                     const double2* __restrict__ G, const double2* __restrict__ tw, OlsConst c,   // c.V: outputs OWNED per block
                     const int* __restrict__ shift_all, int utt0, int n_blocks_max, int seg_cap,
                     int* __restrict__ seg_count,          // [lists][n_blocks_max], zeroed by the caller
                     double* __restrict__ seg_edges) {     // [lists][n_blocks_max][seg_cap]
  WB_DYN_SMEM(double2, smem2);
  __shared__ unsigned long long wtot[THREADS / 32];
  const int u = utt0 + blockIdx.y;
  const int y_len = y_len_all[u];
  const int n0 = blockIdx.x * c.V;
  if (n0 >= y_len - 1) return;                      // no event can start at the last sample
  constexpr int LM = LOG2BN > 0 ? LOG2BN - 1 : 0;
  constexpr int TWL = LOG2BN > 0 ? LOG2BN : kTwLog2;
  const int log2m = LOG2BN > 0 ? LOG2BN - 1 : c.log2bn - 1, M = 1 << log2m;
  double2* xs = smem2;
  double2* ws = smem2 + cpad_size(M);
  double* xsd = reinterpret_cast<double*>(xs);
  double* wsd = reinterpret_cast<double*>(ws);
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, wid = tid >> 5;
  const double* __restrict__ x = x_all + x_off[u];
  const int xl = x_len[u];
  const int mask = fft_mask_all[u];
  const double mean = mean_all[u];
  for (int i = tid; i < c.bn; i += T) {             // input block: ypad[(n0 - D + i) mod FS]
    const int m = (n0 - c.D + i) & mask;
    double v = 0.0;
    if (m < y_len) v = (m < xl ? x[m] : 0.0) - mean;
    xsd[rfft_in_slot(i, log2m)] = v;
  }
  fft_dit<LM, false, THREADS, MAXK, TWL>(xs, log2m, tw);
  for (int k = tid; k <= M / 2; k += T) {           // half spectrum in place: slot k = X[k] (k < M), slot 0 = (X[0], X[M])
    if (k == 0) {
      const double2 z0 = xs[0];
      xs[0] = make_double2(z0.x + z0.y, z0.x - z0.y);
    } else {
      const double2 a = rfft_bin<TWL>(xs, log2m, k, tw);
      const double2 b = rfft_bin<TWL>(xs, log2m, M - k, tw);
      xs[cpad(k)] = a;
      xs[cpad(M - k)] = b;
    }
  }
  __syncthreads();
  const int n_own = min(c.V, y_len - 1 - n0);       // samples of this block at which an event may start
  const int per = (c.V + T - 1) / T;
  const int lo = min(n_own, tid * per), hi = min(n_own, lo + per);
  for (int b = 0; b < c.nb; ++b) {
    const double2* __restrict__ Gb = G + (size_t)b * (M + 1);
    for (int k = tid; k <= M / 2; k += T) {
      if (k == 0) {
        const double2 x0 = xs[0];
        const double2 y0 = make_double2(x0.x * Gb[0].x, 0.0), yM = make_double2(x0.y * Gb[M].x, 0.0);
        ws[cpad(brev(0, log2m))] = c2r_pack<TWL>(y0, yM, 0, log2m, tw);
      } else {
        const double2 yk = cmul(xs[cpad(k)], Gb[k]);
        const double2 ym = cmul(xs[cpad(M - k)], Gb[M - k]);
        ws[cpad(brev(k, log2m))] = c2r_pack<TWL>(yk, ym, k, log2m, tw);
        if (k != M - k) ws[cpad(brev(M - k, log2m))] = c2r_pack<TWL>(ym, yk, M - k, log2m, tw);
      }
    }
    fft_dit<LM, true, THREADS, MAXK, TWL>(ws, log2m, tw);
    const int shift = shift_all[b];                 // filtered_b[n0 + i] = conv[i + shift]
    auto S = [&](int i) { return wsd[rfft_out_slot(i + shift)]; };
    // events that start at my samples (GetFourZeroCrossingIntervals :402-435): bit (i - lo) of m[type]
    unsigned m[4] = {0u, 0u, 0u, 0u};
    if (lo < hi) {
      double s0 = S(lo), s1 = S(lo + 1);
      for (int i = lo; i < hi; ++i) {
        const double s2 = S(i + 2);
        const unsigned bit = 1u << (i - lo);
        if (zc_event(s0, s1)) m[0] |= bit;
        if (zc_event(-s0, -s1)) m[1] |= bit;
        if (n0 + i < y_len - 2) {
          const double d0 = add_rn(s1, -s0), d1 = add_rn(s2, -s1);
          if (zc_event(d0, d1)) m[2] |= bit;
          if (zc_event(-d0, -d1)) m[3] |= bit;
        }
        s0 = s1; s1 = s2;
      }
    }
    // ordered slots: block-wide exclusive scan of the four counts, 16 bits each (a block holds < 2^16 samples)
    const unsigned long long mine = (unsigned long long)__popc(m[0]) | ((unsigned long long)__popc(m[1]) << 16) |
                                    ((unsigned long long)__popc(m[2]) << 32) | ((unsigned long long)__popc(m[3]) << 48);
    unsigned long long inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wtot[wid] = inc;
    __syncthreads();
    unsigned long long before = inc - mine;
    for (int w = 0; w < wid; ++w) before += wtot[w];
    const size_t list0 = ((size_t)blockIdx.y * c.nb + b) * 4;
    if (tid == T - 1) {
      const unsigned long long tot = before + mine;
#pragma unroll
      for (int t = 0; t < 4; ++t)
        seg_count[(list0 + t) * n_blocks_max + blockIdx.x] = (int)((tot >> (16 * t)) & 0xffffull);
    }
#pragma unroll 1
    for (int t = 0; t < 4; ++t) {
      unsigned mm = m[t];
      if (mm == 0u) continue;
      int pos = (int)((before >> (16 * t)) & 0xffffull);
      double* __restrict__ seg = seg_edges + ((list0 + t) * n_blocks_max + blockIdx.x) * seg_cap;
      while (mm) {
        const int j = __ffs(mm) - 1;
        mm &= mm - 1u;
        const int i = lo + j;
        const double s0 = S(i), s1 = S(i + 1);
        double a = s0, bb = s1;
        if (t >= 2) { const double s2 = S(i + 2); a = add_rn(s1, -s0); bb = add_rn(s2, -s1); }
        if (t & 1) { a = -a; bb = -bb; }
        if (pos < seg_cap) seg[pos] = zc_fine(n0 + i + 1, a, bb);
        ++pos;
      }
    }
    __syncthreads();                                // ws and wtot are rewritten by the next band
  }
}

// one thread per list: exclusive offsets of its blocks' segments, list total, overflow flag
static __global__ void zc_seg_scan_kernel(const int* __restrict__ seg_count, int n_lists, int n_blocks_max, int seg_cap,
                                          int* __restrict__ seg_off, int* __restrict__ totals, int* __restrict__ overflow) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_lists) return;
  const int* c = seg_count + (size_t)l * n_blocks_max;
  int* o = seg_off + (size_t)l * n_blocks_max;
  int acc = 0, over = 0;
  for (int i = 0; i < n_blocks_max; ++i) { const int v = c[i]; o[i] = acc; acc += v; over |= v > seg_cap; }
  totals[l] = acc;
  if (over) atomicOr(overflow, 1);
}

// one CTA per list: segments -> the contiguous edge list
static __global__ void __launch_bounds__(128)
zc_seg_gather_kernel(const int* __restrict__ seg_count, const int* __restrict__ seg_off, const double* __restrict__ seg_edges,
                     int n_blocks_max, int seg_cap, const long long* __restrict__ list_off, double* __restrict__ edges) {
  const size_t l = blockIdx.x;
  double* __restrict__ dst = edges + list_off[l];
  for (int k = 0; k < n_blocks_max; ++k) {
    const int n = min(seg_cap, seg_count[l * n_blocks_max + k]);
    const double* __restrict__ src = seg_edges + (l * n_blocks_max + k) * seg_cap;
    const int o = seg_off[l * n_blocks_max + k];
    for (int j = threadIdx.x; j < n; j += blockDim.x) dst[o + j] = src[j];
  }
}


}  // namespace wb
