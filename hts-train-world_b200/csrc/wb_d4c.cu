// world-b200: D4C band aperiodicity, one CTA per frame.
//
// Reference: W/src/d4c.cpp — D4C :337-397, D4CLoveTrain(+Sub) :225-282, D4CGeneralBody :290-316,
// GetCentroid :90-119, GetStaticCentroid :125-142, GetSmoothedPowerSpectrum :148-164,
// GetStaticGroupDelay :170-186, GetCoarseAperiodicity :192-223, GetAperiodicity :325-333,
// GetWindowedWaveform :52-84.
//
// What changed relative to the reference's one-core loop:
//   * the two real FFTs of GetCentroid (x.w and (n+1).x.w) are one complex FFT; the centroid
//     is then (a d + b c)/2 with Z[k] = a+ib, Z[N-k] = c+id — no spectrum is ever unpacked;
//   * band spectra are transformed two at a time the same way;
//   * std::sort + cumulative sum (25 % of the reference's CPU time) is replaced by an exact
//     MSB-first bitwise selection of the (boundary+1)-th largest power held in registers:
//     sorted_cumsum[N/2-boundary-1] == sum of everything below that element;
//   * the randn dither is read from the precomputed table at the offset the sequential
//     reference would have reached (LoveTrain draws of all voiced frames first, then
//     3 windows per processed frame).
#include <stdlib.h>
#include "wb_batch.h"
#include "wb_fft.cuh"
#include "wb_spectral.cuh"

namespace wb {

namespace {

constexpr int kHanning = 1, kBlackman = 2;
constexpr int kLtStage = 2048 + 2;          // LoveTrain: doubles of shared memory for the staged sample window
constexpr int kMaxBands = 8;

__device__ __forceinline__ int d4c_hwl(double ratio, int fs, double f0) {
  return matlab_round(div_rn(div_rn(mul_rn(ratio, (double)fs), f0), 2.0));   // d4c.cpp:55-56
}

struct D4CConst {
  int fs, log2nd, log2lt, nbands, window_length, sel_boundary;
  int lt_b0, lt_b1, lt_b2;
  int out_half;              // fft_size/2 of the output axis
  int band_top;              // highest bin of the group delay any band slice reads
  double threshold;
  int centers[kMaxBands];
};

// windowed waveform with dither and weighted-mean removal (d4c.cpp:52-84).  Sample i is
// written to base[wslot(i)].  STORE_W: its window value goes to base[vslot(i)] (scratch) for the
// mean-removal sweep; otherwise that sweep re-runs the same recurrence (bit-identical values) and
// no scratch is needed.  Returns W.  Contains block syncs; on return every thread has finished
// its own slots only.
template <bool STORE_W, typename S, typename WS, typename VS>
__device__ __forceinline__ int windowed_waveform(const double* __restrict__ x, int x_len, int fs,
                                                 double f0, double position, int window_type,
                                                 double ratio, const uint32_t* __restrict__ rn,
                                                 S* base, WS wslot, VS vslot, double* red,
                                                 const double* staged = nullptr, uint64_t* staged_bar = nullptr,
                                                 const double* phasors = nullptr) {   // {cs0, sn0, cs_step, sn_step} when the caller has them
  const int T = blockDim.x, tid = threadIdx.x;
  const int hwl = d4c_hwl(ratio, fs, f0);
  const int W = 2 * hwl + 1;
  const int origin = matlab_round(add_rn(mul_rn(position, (double)fs), 0.001));
  const double turn_step = 2.0 * f0 / (ratio * fs);          // angle step in units of pi (sincospi: no range reduction)
  double s[2] = {0.0, 0.0};
  // cos(a_i), a_i = (i - hwl) * ang_step, i = tid + j T: one sincos per thread for j = 0, then
  // the angle-addition recurrence with the block-uniform step T * ang_step (<= 16 steps, so
  // the accumulated rounding stays below 1e-15); cos(2a) = 2 cos^2(a) - 1.
  double cs0, sn0, cs_step, sn_step;
  if (phasors) { cs0 = phasors[0]; sn0 = phasors[1]; cs_step = phasors[2]; sn_step = phasors[3]; }
  else {
    sincospi((double)(tid - hwl) * turn_step, &sn0, &cs0);
    sincospi((double)T * turn_step, &sn_step, &cs_step);
  }
  auto window_at = [&](double cs) {
    return window_type == kHanning ? 0.5 * cs + 0.5 : 0.42 + 0.5 * cs + 0.08 * (2.0 * cs * cs - 1.0);
  };
  double cs = cs0, sn = sn0;
  if (staged_bar) mbar_wait(staged_bar, 0);          // the bulk copy of the samples has landed (phase 0)
  for (int i = tid; i < W; i += T) {
    const double w = window_at(cs);
    {
      const double c2 = cs * cs_step - sn * sn_step;
      sn = sn * cs_step + cs * sn_step;
      cs = c2;
    }
    const double xv = staged ? staged[i] : x[min(x_len - 1, max(0, origin + i - hwl))];
    const double wave = xv * w + randn_from_u32(rn[i]) * kMySafeGuardMinimum;
    base[wslot(i)] = static_cast<S>(wave);
    if (STORE_W) base[vslot(i)] = static_cast<S>(w);
    s[0] += wave;
    s[1] += w;
  }
  block_sum<2>(s, red);
  const double coef = s[0] / s[1];
  if (STORE_W) {
    for (int i = tid; i < W; i += T) base[wslot(i)] = static_cast<S>((double)base[wslot(i)] - (double)base[vslot(i)] * coef);
  } else {
    cs = cs0; sn = sn0;
    for (int i = tid; i < W; i += T) {
      base[wslot(i)] = static_cast<S>((double)base[wslot(i)] - window_at(cs) * coef);
      const double c2 = cs * cs_step - sn * sn_step;
      sn = sn * cs_step + cs * sn_step;
      cs = c2;
    }
  }
  return W;
}

// ---- LoveTrain (d4c.cpp:225-282): ap0 for every voiced frame -----------------------------------
__global__ void d4c_lt_count_kernel(const double* __restrict__ f0, int n, int fs,
                                    long long* __restrict__ counts) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  const double v = f0[f];
  counts[f] = v == 0.0 ? 0 : 2LL * matlab_round(div_rn(mul_rn(1.5, (double)fs), fmax(v, 40.0))) + 1;
}

template <int LOG2LT>     // 0: size given at run time (c.log2lt)
__global__ void __launch_bounds__(256)
d4c_lovetrain_kernel(UttView u, const int* __restrict__ frame_utt, const double* __restrict__ frame_t,
                     const double* __restrict__ f0_in, const long long* __restrict__ rng_off,
                     const uint32_t* __restrict__ randn_tab, const double2* __restrict__ tw,
                     D4CConst c, double* __restrict__ ap0_out) {
  WB_DYN_SMEM(double2, smem2);
  const int f = blockIdx.x;
  const double f0 = f0_in[f];
  if (f0 == 0.0) { if (threadIdx.x == 0) ap0_out[f] = 0.0; return; }
  constexpr int LM = LOG2LT > 0 ? LOG2LT - 1 : 0;
  constexpr int TWL = LOG2LT > 0 ? LOG2LT : kTwLog2;     // compact twiddle table of this size, or the master table
  const int log2m = LOG2LT > 0 ? LOG2LT - 1 : c.log2lt - 1;
  const int M = 1 << log2m, N = M << 1;
  double2* buf = smem2;
  double* bufd = reinterpret_cast<double*>(buf);
  // layout: [ buf: 2*cpad_size(M) doubles | red: 96 ]
  const int vbase = 2 * cpad_size(M);
  double* red = bufd + vbase;
  const int tid = threadIdx.x, T = blockDim.x;
  const int utt = frame_utt[f];
  const double* __restrict__ x = u.x + u.x_off[utt];
  const double cur_f0 = fmax(f0, 40.0);
  auto wslot = [log2m](int i) { return rfft_in_slot(i, log2m); };
  auto vslot = [vbase](int i) { return vbase + i; };
  const int W = windowed_waveform<false>(x, u.x_len[utt], c.fs, cur_f0, frame_t[f], kBlackman, 3.0,
                                         randn_tab + rng_off[f], bufd, wslot, vslot, red);
  for (int i = W + tid; i < N; i += T) bufd[rfft_in_slot(i, log2m)] = 0.0;
  fft_dit<LM, false, 256, 3, TWL>(buf, log2m, tw);
  double s[2] = {0.0, 0.0};
  for (int k = c.lt_b0 + 1 + tid; k <= c.lt_b2; k += T) {
    const double2 X = rfft_bin<TWL>(buf, log2m, k, tw);
    const double p = X.x * X.x + X.y * X.y;
    if (k <= c.lt_b1) s[0] += p;
    s[1] += p;
  }
  block_sum<2>(s, red);
  if (tid == 0) ap0_out[f] = s[0] / s[1];
}

// LoveTrain with its transform in FP32 (tests/test_precision_budget.py: ap0 moves by < 1e-7 on
// every input, the only use of ap0 is the comparison with the threshold).  Window, mean removal
// and the band sums stay in FP64.  dynamic shared memory: [ buf: cpadf(M) + 4 float2 | red: 96 doubles | stage: kLtStage doubles ]
template <int LOG2LT>
__global__ void __launch_bounds__(256)
d4c_lovetrain32_kernel(UttView u, const int* __restrict__ frame_utt, const double* __restrict__ frame_t,
                       const double* __restrict__ f0_in, const long long* __restrict__ rng_off,
                       const uint32_t* __restrict__ randn_tab, const float2* __restrict__ twf,
                       D4CConst c, double* __restrict__ ap0_out) {
  WB_DYN_SMEM(double2, smem2);
  const int f = blockIdx.x;
  const double f0 = f0_in[f];
  if (f0 == 0.0) { if (threadIdx.x == 0) ap0_out[f] = 0.0; return; }
  constexpr int LM = LOG2LT - 1, M = 1 << LM, N = M << 1;
  float2* buf = reinterpret_cast<float2*>(smem2);
  float* buff = reinterpret_cast<float*>(buf);
  double* red = reinterpret_cast<double*>(buf + ((cpadf(M) + 4 + 1) & ~1));
  double* stage = red + 96;                      // kLtStage doubles: the frame's sample window, by TMA
  const int tid = threadIdx.x, T = blockDim.x;
  const int utt = frame_utt[f];
  const double* __restrict__ x = u.x + u.x_off[utt];
  const double cur_f0 = fmax(f0, 40.0);
  // the sample window is a contiguous range of the utterance: one bulk copy (cp.async.bulk + mbarrier) into
  // shared memory, issued before the window coefficients are set up; windows that cross an utterance edge
  // (the reference clamps the index) or exceed the staging area (f0 < ~70 Hz) take the clamped gather
  __shared__ uint64_t mbar;
  const int hwl = d4c_hwl(3.0, c.fs, cur_f0);
  const int g0 = matlab_round(add_rn(mul_rn(frame_t[f], (double)c.fs), 0.001)) - hwl;
  int st_a0 = 0, st_n = 0;
  const bool st_ok = bulk_window_range(g0, 2 * hwl + 1, u.x_len[utt], kLtStage, &st_a0, &st_n);
  if (st_ok) {
    if (tid == 0) mbar_init(&mbar, 1);
    __syncthreads();
    if (tid == 0) bulk_load_issue(stage, x + st_a0, (unsigned)st_n * 8u, &mbar);
  }
  auto wslot = [](int i) { return rfft_in_slot_f(i, LM); };
  auto vslot = [](int i) { return i; };
  const int W = windowed_waveform<false>(x, u.x_len[utt], c.fs, cur_f0, frame_t[f], kBlackman, 3.0,
                                         randn_tab + rng_off[f], buff, wslot, vslot, red,
                                         st_ok ? stage + (g0 - st_a0) : nullptr, st_ok ? &mbar : nullptr);
  for (int i = W + tid; i < N; i += T) buff[rfft_in_slot_f(i, LM)] = 0.f;
  fft_dit<LM, false, 256, 3, LOG2LT>(buf, LM, twf);
  double s[2] = {0.0, 0.0};
  for (int k = c.lt_b0 + 1 + tid; k <= c.lt_b2; k += T) {
    const float2 X = rfft_bin<LOG2LT>(buf, LM, k, twf);
    const double p = (double)X.x * X.x + (double)X.y * X.y;
    if (k <= c.lt_b1) s[0] += p;
    s[1] += p;
  }
  block_sum<2>(s, red);
  if (tid == 0) ap0_out[f] = s[0] / s[1];
}

__global__ void d4c_main_count_kernel(const double* __restrict__ f0, const double* __restrict__ ap0,
                                      int n, int fs, double threshold, long long* __restrict__ counts) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  const double v = f0[f];
  if (v == 0.0 || ap0[f] <= threshold) { counts[f] = 0; return; }   // d4c.cpp:380
  counts[f] = 3LL * (2LL * d4c_hwl(4.0, fs, fmax(kFloorF0D4C, v)) + 1);
}

// ---- coarse aperiodicity working set ---------------------------------------------------------
// The band spectra (GetCoarseAperiodicity :192-223) are computed in FP32: each is the FFT of
// a Nuttall-windowed slice of the static group delay, and only the ratio
// (sum of all but the boundary+1 largest powers) / (sum of all powers) is used, so an FFT
// error of 2^-24 relative to the largest bin moves the aperiodicity by ~1e-7 (tolerance 1e-4).
// Layout inside the [cbuf | pw] region (bytes): [0, fb) FP32 FFT buffer, reused as the
// selection histograms [0, nbands * 2048 * 4); then nbands power arrays of Hd + 4 floats.
constexpr int kSelBins = 2048;          // 11-bit digits of the 31-bit float keys
constexpr int kMaxSets = 6;
__host__ __device__ constexpr int d4c_band_p_off(int nd, int nbands) {
  return ((cpad_size(nd) * 8 > nbands * kSelBins * 4 ? cpad_size(nd) * 8 : nbands * kSelBins * 4) + 15) & ~15;
}
__host__ __device__ constexpr int d4c_cbuf_slots(int nd, int nbands) {   // double2 slots of cbuf
  const int need = d4c_band_p_off(nd, nbands) + nbands * (nd / 2 + 4) * 4 - (nd / 2 + 8) * 8;   // pw follows cbuf
  const int need_slots = (need + 15) / 16;
  return cpad_size(nd) > need_slots ? cpad_size(nd) : need_slots;
}

struct SelectScratch {
  unsigned long long wsum[2][32];
  int digit[kMaxSets], above[kMaxSets], cand[kMaxSets];
};

// Exact sum of everything below the K-th largest of each of `nsets` arrays of n non-negative
// floats (P + s * pstride).  MSB-first radix selection on the IEEE bit patterns (order-
// isomorphic to the values for non-negative floats) with 11-bit digits: every pass histograms
// the digit of the keys that still match the decided prefix, finds the bin holding the K-th
// largest with one block-wide suffix scan (all sets packed into two 64-bit words, 21 bits
// each) and narrows the prefix.  A set is finished as soon as all keys matching its prefix
// belong to the top set (always true once one candidate is left): 2 passes typically, 3 at
// most -- instead of the reference's std::sort of every band (25 % of its CPU time).
// Result: low[s] = sum of the (n - K) smallest values, tot[s] = sum of all (both in FP64).
template <int T>
__device__ __forceinline__ void select_low_sums(const float* __restrict__ P, int pstride, int nsets, int n,
                                                int K, int* hist, SelectScratch* sc, double* red,
                                                double* low, double* tot) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int nw = T >> 5, BPT = kSelBins / T;
  // bin b = t * BPT + q lives at [q * T + t]: a thread's own bins are conflict-free
  auto hslot = [](unsigned bin) { return (int)((bin & (BPT - 1)) * T + (bin / BPT)); };
  unsigned pre[kMaxSets], mask_hi = 0u;
  int k_rem[kMaxSets];
  bool done[kMaxSets];
#pragma unroll
  for (int s = 0; s < kMaxSets; ++s) { pre[s] = 0u; k_rem[s] = K; done[s] = s >= nsets; }
  for (int shift = 31 - 11; ; shift -= 11) {
    const int sh = shift < 0 ? 0 : shift;
    const unsigned dmask = (1u << (shift < 0 ? 11 + shift : 11)) - 1u;
    for (int i = tid; i < nsets * kSelBins / 4; i += T) reinterpret_cast<int4*>(hist)[i] = make_int4(0, 0, 0, 0);
    __syncthreads();
#pragma unroll
    for (int s = 0; s < kMaxSets; ++s) {
      if (done[s]) continue;
      const float* __restrict__ ps = P + s * pstride;
      for (int k = tid; k < n; k += T) {
        const unsigned key = __float_as_uint(ps[k]);
        if ((key & mask_hi) == pre[s]) atomicAdd(&hist[s * kSelBins + hslot((key >> sh) & dmask)], 1);
      }
    }
    __syncthreads();
    // per-thread bin totals, suffix-summed over threads; 3 sets x 21 bits per 64-bit word
    unsigned mine[kMaxSets];
    unsigned long long v[2] = {0ull, 0ull};
#pragma unroll
    for (int s = 0; s < kMaxSets; ++s) {
      mine[s] = 0u;
      if (done[s]) continue;
#pragma unroll
      for (int q = 0; q < BPT; ++q) mine[s] += hist[s * kSelBins + q * T + tid];
      v[s / 3] |= (unsigned long long)mine[s] << (21 * (s % 3));
    }
    unsigned long long inc[2] = {v[0], v[1]};
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t0 = __shfl_down_sync(0xffffffffu, inc[0], o);
      const unsigned long long t1 = __shfl_down_sync(0xffffffffu, inc[1], o);
      if (lane + o < 32) { inc[0] += t0; inc[1] += t1; }
    }
    if (lane == 0) { sc->wsum[0][wid] = inc[0]; sc->wsum[1][wid] = inc[1]; }
    __syncthreads();
    unsigned long long above[2] = {inc[0] - v[0], inc[1] - v[1]};
    for (int w = wid + 1; w < nw; ++w) { above[0] += sc->wsum[0][w]; above[1] += sc->wsum[1][w]; }
#pragma unroll
    for (int s = 0; s < kMaxSets; ++s) {
      if (done[s]) continue;
      const int ab = (int)((above[s / 3] >> (21 * (s % 3))) & 0x1fffffull);
      if (ab < k_rem[s] && k_rem[s] <= ab + (int)mine[s]) {      // the K-th largest is in my bins
        int acc = ab;
        for (int q = BPT - 1; q >= 0; --q) {
          const int h = hist[s * kSelBins + q * T + tid];
          if (acc < k_rem[s] && k_rem[s] <= acc + h) { sc->digit[s] = tid * BPT + q; sc->above[s] = acc; sc->cand[s] = h; break; }
          acc += h;
        }
      }
    }
    __syncthreads();
    mask_hi |= dmask << sh;
    bool all_done = true;
#pragma unroll
    for (int s = 0; s < kMaxSets; ++s) {
      if (done[s]) continue;
      pre[s] |= (unsigned)sc->digit[s] << sh;
      k_rem[s] -= sc->above[s];
      if (sc->cand[s] == k_rem[s]) done[s] = true;   // every key with this prefix is in the top set
      all_done = all_done && done[s];
    }
    if (all_done || sh == 0) break;
  }
  // top set = keys whose decided bits are >= pre; when the digits ran out with ties left
  // (!done), only k_rem copies of the value `pre` belong to it.
  for (int s = 0; s < nsets; ++s) {
    const float* __restrict__ ps = P + s * pstride;
    double acc[3] = {0.0, 0.0, 0.0};
    for (int k = tid; k < n; k += T) {
      const float pv = ps[k];
      const unsigned key = __float_as_uint(pv) & mask_hi;
      acc[0] += pv;
      if (key < pre[s]) acc[1] += pv;
      if (!done[s] && key == pre[s]) acc[2] += 1.0;
    }
    block_sum<3>(acc, red);
    tot[s] = acc[0];
    low[s] = acc[1];
    if (!done[s]) low[s] += (acc[2] - (double)k_rem[s]) * (double)__uint_as_float(pre[s]);
  }
}

// The same quantity for ONE array handled by ONE warp (compile-time sizes): the K-th largest key is
// found by an MSB-first binary search on the bit pattern, each step one register sweep (key >= candidate)
// plus a warp reduction -- no shared memory traffic, no atomics and no block barriers, so the bands of a
// frame are selected concurrently by different warps.  The search stops as soon as exactly K keys lie at
// or above the candidate (then they are the top set); otherwise it ends with T = the K-th largest key and
// the low set receives its surplus copies of T.
//
// Two phases.  (1) The upper 16 bits of a non-negative float are its bfloat16 truncation, and bfloat16
// values order like their bit patterns, so the first 16 bits are decided on PACKED keys: two upper halves
// per register, one HSET2.BF16 (two comparisons, a 16-bit mask per half) per pair of keys and one three-input
// integer addition per two pairs (bf16x2_ge_mask) -- a fifth of the issue cycles of the compare /
// predicated-add pair per key of the 32-bit sweep, which ran on the half-rate integer pipe.  (2) Only when the K-th and the
// (K+1)-th largest key share their upper half (a bucket 0.8 % wide) do the full keys come back from shared
// memory and the search continues on bits 15 .. 0 with 32-bit sweeps; inside one bucket the lower bits are
// as good as random, so this takes a few steps.
// (keys >= cand2) per half as a mask, 0xffff or 0.  Subtracting the masks of several registers from an integer
// accumulator counts both halves at once: -0xffff = 1 - 0x10000, so the low half of the accumulator ends up
// with the number of low hits and the high half with (high hits - low hits) mod 2^16 -- one three-input integer
// addition takes two masks, i.e. 1.5 instructions per PAIR of keys and step.
__device__ __forceinline__ unsigned bf16x2_ge_mask(unsigned keys, unsigned cand2) {
#ifdef WB_HOST_EMU
  return ((keys & 0xffffu) >= (cand2 & 0xffffu) ? 0xffffu : 0u) | ((keys >> 16) >= (cand2 >> 16) ? 0xffff0000u : 0u);
#else
  unsigned m;
  asm("set.ge.u32.bf16x2 %0, %1, %2;" : "=r"(m) : "r"(keys), "r"(cand2));
  return m;
#endif
}
__device__ __forceinline__ int packed_counts_total(unsigned acc) {
  const unsigned lo = acc & 0xffffu;
  return (int)(lo + (((acc >> 16) + lo) & 0xffffu));
}

template <int NPL>     // keys per lane: ceil(n / 32)
__device__ __forceinline__ void warp_select_low_sum(const float* __restrict__ P, int n, int K,
                                                    double* low, double* tot) {
  const int lane = threadIdx.x & 31;
  unsigned T = 0u;
  int at_or_above = 32 * NPL;                 // keys >= T
  bool exact = false;
#ifndef WB_SELECT_32BIT
  {
    constexpr int NPK = (NPL + 1) / 2;
    unsigned h[NPK];
    unsigned mx = 0u;
#pragma unroll
    for (int j = 0; j < NPK; ++j) {
      const int k0 = lane + 64 * j, k1 = k0 + 32;
      const unsigned a = k0 < n ? __float_as_uint(P[k0]) : 0u;
      const unsigned b = (2 * j + 1 < NPL && k1 < n) ? __float_as_uint(P[k1]) : 0u;
      mx = max(mx, max(a, b));
#ifdef WB_HOST_EMU
      h[j] = (a >> 16) | (b & 0xffff0000u);
#else
      h[j] = __byte_perm(a, b, 0x7632);        // upper half of a | upper half of b
#endif
    }
    mx = __reduce_max_sync(0xffffffffu, mx) >> 16;
    unsigned T16 = 0u;
#pragma unroll 1
    for (int bit = 31 - __clz(mx | 1u); bit >= 0; --bit) {      // higher bits: no key reaches the candidate
      const unsigned cand = T16 | (1u << bit);
      const unsigned cand2 = cand * 0x10001u;
      unsigned c[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int j = 0; j + 1 < NPK; j += 2)        // the pad keys are 0 < cand
        c[(j >> 1) & 3] = c[(j >> 1) & 3] - bf16x2_ge_mask(h[j], cand2) - bf16x2_ge_mask(h[j + 1], cand2);
      if (NPK & 1) c[0] -= bf16x2_ge_mask(h[NPK - 1], cand2);
      const int cnt = __reduce_add_sync(0xffffffffu, packed_counts_total((c[0] + c[1]) + (c[2] + c[3])));
      if (cnt >= K) {
        T16 = cand;
        at_or_above = cnt;
        if (cnt == K) { exact = true; break; }
      }
    }
    T = T16 << 16;
  }
#endif
  unsigned key[NPL];
#pragma unroll
  for (int j = 0; j < NPL; ++j) {
    const int k = lane + 32 * j;
    key[j] = k < n ? __float_as_uint(P[k]) : 0u;
  }
#ifdef WB_SELECT_32BIT
  unsigned mx = 0u;
#pragma unroll
  for (int j = 0; j < NPL; ++j) mx = max(mx, key[j]);
  mx = __reduce_max_sync(0xffffffffu, mx);
  const int first_bit = 31 - __clz(mx | 1u);
#else
  const int first_bit = 15;
#endif
  if (!exact) {
#pragma unroll 1
    for (int bit = first_bit; bit >= 0; --bit) {
      const unsigned cand = T | (1u << bit);
      int c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int j = 0; j < NPL; ++j)              // one compare and one predicated increment per key
#ifdef WB_HOST_EMU
        c[j & 7] += key[j] >= cand;
#else
        asm("{.reg .pred p; setp.ge.u32 p, %1, %2; @p add.s32 %0, %0, 1;}" : "+r"(c[j & 7]) : "r"(key[j]), "r"(cand));
#endif
      const int cnt = __reduce_add_sync(0xffffffffu, ((c[0] + c[1]) + (c[2] + c[3])) + ((c[4] + c[5]) + (c[6] + c[7])));
      if (cnt >= K) {
        T = cand;
        at_or_above = cnt;
        if (cnt == K) break;
      }
    }
  }
  // sums: groups of 8 keys are added in single precision (relative error <= 4 ulp of a float, the size of the
  // rounding of the FP32 transform that produced the powers), the groups in double -- one conversion and two
  // FP64 additions per 8 keys instead of per key
  double a_tot = 0.0, a_low = 0.0;
#pragma unroll
  for (int j0 = 0; j0 < NPL; j0 += 8) {
    float g_tot = 0.f, g_low = 0.f;
#pragma unroll
    for (int j = j0; j < j0 + 8 && j < NPL; ++j) {
      const float v = __uint_as_float(key[j]);
      g_tot += v;
      g_low += key[j] < T ? v : 0.f;
    }
    a_tot += (double)g_tot;
    a_low += (double)g_low;
  }
  a_tot = warp_sum(a_tot);
  a_low = warp_sum(a_low);
  *tot = a_tot;
  *low = a_low + (double)(at_or_above - K) * (double)__uint_as_float(T);
}

// First radix-16 pass of the band transform fused with its input: the Nuttall-windowed slice has
// window_length <= 3 * Nd/16 non-zero samples, so of the 16 inputs x[il + r Nd/16] of a
// first-pass group only r = 0, 1 (and r = 2 for the first few il) are non-zero and the 16-point
// DFT collapses to  out[q] = x0 + x1 w^q + x2 w^2q,  w = exp(-2 pi i / 16).  The zero fill, the
// bit-reversed scatter and the first pass's loads disappear; the later passes run unchanged.
// Element .x of every point carries band a, .y band b (two real sequences per transform).
template <int LOG2N, int THREADS>
__device__ __forceinline__ void band_fft_pruned(float2* __restrict__ fb, const double* __restrict__ cen, int ca,
                                                int cb, bool two, const double* __restrict__ nuttall, int wl,
                                                const float2* __restrict__ twf) {
  static_assert(fft_plan<LOG2N, 4>::k_of(0) == 4, "first pass must be radix-16");
  constexpr int G = 1 << (LOG2N - 4);
  // lane <-> il: the reads of cen are linear and the stores land on slots brev(il) + q, the same
  // conflict-free bit-reversed scatter as a plain input permutation
  for (int il = threadIdx.x; il < G; il += THREADS) {
    const int g = brev(il, LOG2N - 4);
    float2 x[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int i = il + r * G;
      x[r] = make_float2(0.f, 0.f);
      if (i < wl) {
        const double w = nuttall[i];
        x[r].x = static_cast<float>(cen[ca + i] * w);
        if (two) x[r].y = static_cast<float>(cen[cb + i] * w);
      }
    }
    float2* __restrict__ sb = fb + cpadf(g << 4);
    const bool has2 = il + 2 * G < wl;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float2 t = rot16<false>(x[1], q);
      float2 lo = cadd(x[0], t), hi = csub(x[0], t);
      if (has2) {                                  // w^(2q) = w^(2(q+8)); for 2q >= 8 it is -w^(2q-8)
        float2 u = rot16<false>(x[2], (2 * q) & 7);
        if (2 * q >= 8) u = make_float2(-u.x, -u.y);
        lo = cadd(lo, u);
        hi = cadd(hi, u);
      }
      sb[cpadf(q)] = lo;
      sb[cpadf(q + 8)] = hi;
    }
  }
  fft_finish_after_first_pass<LOG2N, 4, false, THREADS, LOG2N, false>(fb, twf);
}

// dynamic shared memory: [ cen: Hd+8 | cbuf: d4c_cbuf_slots double2 | pw: Hd+8 | red: 96 |
//                          SelectScratch | coarse: kMaxBands+2 ]
template <int LOG2ND, int THREADS, int MAXK = 3>    // LOG2ND 0: size given at run time (c.log2nd)
__global__ void __launch_bounds__(THREADS, 2)
d4c_main_kernel(UttView u, const int* __restrict__ frame_utt, const double* __restrict__ frame_t,
                const double* __restrict__ f0_in, const double* __restrict__ ap0,
                const long long* __restrict__ rng_off, const long long* __restrict__ lt_totals,
                const uint32_t* __restrict__ randn_tab, const double2* __restrict__ tw,
                const float2* __restrict__ twf, const double* __restrict__ nuttall, D4CConst c,
                double* __restrict__ ap_out) {
  WB_DYN_SMEM(double2, smem2);
  const int log2nd = LOG2ND > 0 ? LOG2ND : c.log2nd;
  constexpr int LMD = LOG2ND > 0 ? LOG2ND - 1 : 0;
  constexpr int TWL = LOG2ND > 0 ? LOG2ND : kTwLog2;     // compact twiddle tables of this size, or the master tables
  const int Nd = 1 << log2nd, Hd = Nd >> 1;
  double* cen = reinterpret_cast<double*>(smem2);
  double2* cbuf = reinterpret_cast<double2*>(cen + Hd + 8);
  double* cbufd = reinterpret_cast<double*>(cbuf);
  double* pw = reinterpret_cast<double*>(cbuf + d4c_cbuf_slots(Nd, c.nbands));
  double* red = pw + Hd + 8;
  SelectScratch* sc = reinterpret_cast<SelectScratch*>(red + 160);
  double* coarse = reinterpret_cast<double*>(sc + 1);
  const int tid = threadIdx.x;
  constexpr int T = THREADS;
  const int f = blockIdx.x;
  double* __restrict__ out = ap_out + (size_t)f * (c.out_half + 1);
  const double f0 = f0_in[f];
  if (f0 == 0.0 || ap0[f] <= c.threshold) {           // d4c.cpp:380 + InitializeAperiodicity
    for (int k = tid; k <= c.out_half; k += T) out[k] = 1.0 - kMySafeGuardMinimum;
    return;
  }
  // ---- GetAperiodicity (:325-333): interp1 over {0, 3k, ..., fs/2} then 10^(dB/20) ---------------
  auto write_row = [&]() {
    if (tid == 0) { coarse[0] = -60.0; coarse[c.nbands + 1] = -kMySafeGuardMinimum; }
    __syncthreads();
    // The row leaves as exp() of a single-precision argument (6e-8 relative, tolerance 1e-4 absolute), so the
    // piecewise-linear interpolation over the knots {0, 3k, ..., 3k nbands, fs/2} runs in single precision
    // as well: bin frequencies k fs / (2 out_half) and knots are exact there, the segment is found by one
    // multiplication; when its rounding puts a bin that sits ON a knot into the segment before it, s is 1
    // there and the interpolant is continuous.
    const int N_out = 2 * c.out_half;
    const float bin_hz = (float)c.fs / (float)N_out, top_hz = 0.5f * (float)c.fs;
    const float last_x0 = (float)(c.nbands * kFrequencyInterval);
    const float inv_last = 1.0f / (top_hz - last_x0);
    for (int k = tid; k <= c.out_half; k += T) {
      const float xi = (float)k * bin_hz;
      const int seg = min(c.nbands + 1, 1 + (int)(xi * (float)(1.0 / kFrequencyInterval)));   // clamp(upper_bound(axis, xi), 1, nk-1)
      const float x0 = (float)(seg - 1) * (float)kFrequencyInterval;
      const float s = (xi - x0) * (seg <= c.nbands ? (float)(1.0 / kFrequencyInterval) : inv_last);
      const float y0 = (float)coarse[seg - 1], y1 = (float)coarse[seg];
      const float v = fmaf(s, y1 - y0, y0);
      out[k] = static_cast<double>(expf(v * 0.11512925464970228420f));      // 10^(v/20)
    }
  };
  // fs < 12 kHz: no 3 kHz band fits below fs/2 - 3 kHz (d4c.cpp:351-353), the reference still runs
  // D4CGeneralBody but nothing of it reaches the row: knots {0: -60 dB, fs/2: -1e-12 dB} only
  if (c.nbands == 0) { write_row(); return; }
  const int utt = frame_utt[f];
  const double* __restrict__ x = u.x + u.x_off[utt];
  const int x_len = u.x_len[utt];
  const double t_pos = frame_t[f];
  const double cur_f0 = fmax(kFloorF0D4C, f0);
  const uint32_t* __restrict__ rn = randn_tab + lt_totals[utt] + rng_off[f];
  const int W4 = 2 * d4c_hwl(4.0, c.fs, cur_f0) + 1;
#if !defined(WB_HOST_EMU) && !defined(WB_NO_RN_PREFETCH)
  // the frame's 3 W4 dither words are read once, by this CTA only, straight from HBM: ask L2 for all of them now
  // (one 128-byte line per thread and trip), the second and third window then find them there
  for (int i = tid * 32; i < 3 * W4; i += T * 32) asm volatile("prefetch.global.L2 [%0];" :: "l"(rn + i));
#endif
  if (W4 > Nd || Hd + 2 * smoothing_boundary(cur_f0, c.fs, Nd) + 1 > 2 * cpad_size(Nd)) {
    for (int k = tid; k <= c.out_half; k += T) out[k] = __longlong_as_double(0x7ff8000000000000LL);
    return;
  }
  auto cslot = [log2nd](int i) { return cpad(brev(i, log2nd)); };
  // Nothing above the highest band slice is ever read (at 48 kHz the five bands end at 18 of 24 kHz),
  // so every per-bin sweep stops at the bin its consumer needs: the group delay up to band_top, the
  // input of each smoothing its boundary further (smoothing_input_need), all clamped to Nd / 2.
  const int b_f0 = smoothing_boundary(cur_f0, c.fs, Nd), b_half = smoothing_boundary(cur_f0 / 2.0, c.fs, Nd);
  const int k_gd = min(Hd, c.band_top);                                       // static group delay
  const int k_s2 = min(Hd, smoothing_input_need(k_gd, b_f0));                 // first smoothing of the ratio (input of the second)
  const int k_ratio = min(Hd, smoothing_input_need(k_s2, b_half));            // centroid / power ratio, smoothed power spectrum, centroid
  const int k_pw = min(Hd, smoothing_input_need(k_ratio, b_f0));              // raw power spectrum

  // The three sample windows of the frame (two centroids, power spectrum) are staged into the
  // idle `pw` array by TMA bulk copies, each issued one phase ahead so that it lands while the
  // previous window is being transformed; windows that cross an utterance edge (the reference
  // clamps the index there) or exceed the staging area fall back to a clamped gather.
  __shared__ uint64_t mbar;
  const int hwl_w = d4c_hwl(4.0, c.fs, cur_f0);
  auto window_origin = [&](int which) {                 // 0 / 1: centroid sides, 2: power spectrum
    const double pos = which == 2 ? t_pos : add_rn(t_pos, which == 0 ? -0.25 / cur_f0 : 0.25 / cur_f0);
    return matlab_round(add_rn(mul_rn(pos, (double)c.fs), 0.001)) - hwl_w;
  };
  int st_a0 = 0, st_n = 0;
  bool st_ok = bulk_window_range(window_origin(0), W4, x_len, Hd + 8, &st_a0, &st_n);
  unsigned st_parity = 0;
  if (tid == 0) mbar_init(&mbar, 1);
  __syncthreads();
  if (st_ok && tid == 0) bulk_load_issue(pw, x + st_a0, (unsigned)st_n * 8u, &mbar);

  // The three windows of the frame (two Blackman, one Hanning) have the same length and the same phase
  // cos(pi (i - hwl) 2 f0 / (4 fs)) at sample i: the thread's start phasor and the step of T samples are
  // computed once per frame (two double-precision sincospi instead of six).
  double ph[4];
  {
    const double turn_step = 2.0 * cur_f0 / (4.0 * c.fs);   // angle step in units of pi
    sincospi((double)(tid - hwl_w) * turn_step, &ph[1], &ph[0]);
    sincospi((double)T * turn_step, &ph[3], &ph[2]);
  }
  // ---- GetStaticCentroid (:125-142): two centroids, each one packed complex FFT ----------------
  for (int side = 0; side < 2; ++side) {
    const double pos = add_rn(t_pos, side == 0 ? -0.25 / cur_f0 : 0.25 / cur_f0);
    __syncthreads();                                    // previous readers of cbuf are done
    // windowed waveform (d4c.cpp:52-84), mean removal and unit-energy normalisation (:95-100) with a
    // single block reduction: with S1 = sum(wave), Sw = sum(w), S2 = sum(wave^2), Sxw = sum(wave w),
    // Sww = sum(w^2):  coef = S1 / Sw,  energy of (wave - w coef) = S2 - 2 coef Sxw + coef^2 Sww.
    const int hwl4 = d4c_hwl(4.0, c.fs, cur_f0);
    const int W = 2 * hwl4 + 1;
    {
      const int origin = matlab_round(add_rn(mul_rn(pos, (double)c.fs), 0.001));
      const uint32_t* __restrict__ rns = rn + (size_t)side * W4;
      double cs = ph[0], sn = ph[1];
      const double cs_step = ph[2], sn_step = ph[3];
      double sums[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
      const bool staged = st_ok;
      const double* xs = pw + (window_origin(side) - st_a0);
      auto window_sample = [&](int i, uint32_t rnd) {
        const double w = 0.42 + 0.5 * cs + 0.08 * (2.0 * cs * cs - 1.0);
        {
          const double c2 = cs * cs_step - sn * sn_step;
          sn = sn * cs_step + cs * sn_step;
          cs = c2;
        }
        const double xv = staged ? xs[i] : x[min(x_len - 1, max(0, origin + i - hwl4))];
        const double wave = xv * w + randn_from_u32(rnd) * kMySafeGuardMinimum;
        cbuf[cslot(i)] = make_double2(wave, w);
        sums[0] += wave; sums[1] += w; sums[2] += wave * wave; sums[3] += wave * w; sums[4] += w * w;
      };
      if (staged) { mbar_wait(&mbar, st_parity); st_parity ^= 1u; }
      for (int i = tid; i < W; i += T) window_sample(i, rns[i]);
      block_sum<5>(sums, red);                          // (its barriers also mean: every thread is done with pw)
      st_ok = bulk_window_range(window_origin(side + 1), W4, x_len, Hd + 8, &st_a0, &st_n);
      if (st_ok && tid == 0) bulk_load_issue(pw, x + st_a0, (unsigned)st_n * 8u, &mbar);   // next window
      const double coef = sums[0] / sums[1];
      double energy = sums[2] - 2.0 * coef * sums[3] + coef * coef * sums[4];
      if (!(energy > 1e-8 * sums[2])) {                 // DC-dominated frame: the expansion cancels, sum directly
        double e[1] = {0.0};
        for (int i = tid; i < W; i += T) { const double2 z = cbuf[cslot(i)]; const double v = z.x - z.y * coef; e[0] += v * v; }
        block_sum<1>(e, red);
        energy = e[0];
      }
      const double inv_sq = 1.0 / sqrt(energy);
      if constexpr (LOG2ND > 0) {
        // mean removal, normalisation, the (n + 1) factor of the second sequence and the zero padding
        // happen while the first pass loads its own slots (the block reductions above are the barrier
        // behind the window loop); slots beyond the window are not even read
        using plan = fft_plan<LOG2ND, MAXK>;
        constexpr int K0 = plan::k_of(0);
        auto load = [&](int base, int m) {
          const int i = brev(base, LOG2ND) | (brev(m, K0) << (LOG2ND - K0));     // element index of slot base + m
          double2 z = make_double2(0.0, 0.0);
          if (i < W) { const double2 q = cbuf[cpad(base) + cpad(m)]; const double v = (q.x - q.y * coef) * inv_sq; z = make_double2(v, v * (i + 1.0)); }
          return z;
        };
        fft_first_pass_from<K0, false, LOG2ND, THREADS>(cbuf, load);
        fft_finish_after_first_pass<LOG2ND, MAXK, false, THREADS, TWL>(cbuf, tw);
      } else {
        for (int i = tid; i < Nd; i += T) {
          double2 z = make_double2(0.0, 0.0);
          if (i < W) { const double2 q = cbuf[cslot(i)]; const double v = (q.x - q.y * coef) * inv_sq; z = make_double2(v, v * (i + 1.0)); }
          cbuf[cslot(i)] = z;
        }
        fft_dit<LOG2ND, false, THREADS, MAXK, TWL>(cbuf, log2nd, tw);
      }
    }
    for (int k = tid; k <= k_ratio; k += T) {
      const double2 A = cbuf[cpad(k)];
      const double2 B = cbuf[cpad((Nd - k) & (Nd - 1))];
      const double cval = 0.5 * (A.x * B.y + A.y * B.x);
      cen[k] = side == 0 ? cval : cen[k] + cval;
    }
  }
  __syncthreads();
  dc_correction<false>(cen, cbufd, cur_f0, c.fs, Nd);  // one warp; the centroid is next read after the power spectrum's barriers (cbuf is idle here as scratch for the block version; pw holds the staged power-spectrum window)

  // ---- GetSmoothedPowerSpectrum (:148-164) ----------------------------------------------------
  {
    const int log2m = log2nd - 1;
    // real FFT input in the first 2*cpad_size(Hd) doubles of cbuf, window scratch after it
    // (2*cpad_size(Nd) - 2*cpad_size(Hd) = 1.125 Nd doubles >= W)
    const int vbase = 2 * cpad_size(Hd);
    auto pwslot = [log2m](int i) { return rfft_in_slot(i, log2m); };
    auto pvslot = [vbase](int i) { return vbase + i; };
    __syncthreads();                                    // centroid readers of cbuf are done
    if (st_ok) mbar_wait(&mbar, st_parity);
    const int W = windowed_waveform<true>(x, x_len, c.fs, cur_f0, t_pos, kHanning, 4.0,
                                    rn + 2 * (size_t)W4, cbufd, pwslot, pvslot, red,
                                    st_ok ? pw + (window_origin(2) - st_a0) : nullptr, nullptr, ph);
    for (int i = W + tid; i < Nd; i += T) cbufd[rfft_in_slot(i, log2m)] = 0.0;
    fft_dit<LMD, false, THREADS, MAXK, TWL>(cbuf, log2m, tw);
    for (int k = tid; k <= k_pw; k += T) {
      const double2 X = rfft_bin<TWL>(cbuf, log2m, k, tw);
      pw[k] = X.x * X.x + X.y * X.y;
    }
    __syncthreads();
    dc_correction<true>(pw, cbufd, cur_f0, c.fs, Nd);
    linear_smoothing(pw, pw, cbufd, red, cur_f0, c.fs, Nd, k_ratio);
  }
  // ---- GetStaticGroupDelay (:170-186) ------------------------------------------------------------
  for (int k = tid; k <= k_ratio; k += T) cen[k] = cen[k] / pw[k];
  __syncthreads();
  linear_smoothing(cen, cen, cbufd, red, cur_f0 / 2.0, c.fs, Nd, k_s2);
  linear_smoothing(cen, pw, cbufd, red, cur_f0, c.fs, Nd, k_gd);
  for (int k = tid; k <= k_gd; k += T) cen[k] -= pw[k];
  __syncthreads();

  // ---- GetCoarseAperiodicity (:192-223): two bands per complex FP32 FFT, one selection for all --
  {
    const int hw = c.window_length / 2;
    float2* fb = reinterpret_cast<float2*>(cbuf);
    float* P = reinterpret_cast<float*>(reinterpret_cast<char*>(cbuf) + d4c_band_p_off(Nd, c.nbands));
    const int pstride = Hd + 4;
    for (int b0 = 0; b0 < c.nbands; b0 += 2) {
      const bool two = b0 + 1 < c.nbands;
      const int ca = c.centers[b0] - hw, cb = two ? c.centers[b0 + 1] - hw : 0;
      bool pruned = false;
      if constexpr (LOG2ND >= 8) {
        if (c.window_length <= 3 * (Nd >> 4)) {      // block-uniform
          band_fft_pruned<LOG2ND, THREADS>(fb, cen, ca, cb, two, nuttall, c.window_length, twf);
          pruned = true;
        }
      }
      if (!pruned) {
        for (int i = tid; i < Nd; i += T) {
          float2 z = make_float2(0.f, 0.f);
          if (i < c.window_length) {
            const double w = nuttall[i];
            z.x = static_cast<float>(cen[ca + i] * w);
            if (two) z.y = static_cast<float>(cen[cb + i] * w);
          }
          fb[cpadf(brev(i, log2nd))] = z;
        }
        fft_dit<LOG2ND, false, THREADS, 4, TWL>(fb, log2nd, twf);
      }
      for (int k = tid; k <= Hd; k += T) {
        const float2 A = fb[cpadf(k)];
        const float2 B = fb[cpadf((Nd - k) & (Nd - 1))];
        const float xr = 0.5f * (A.x + B.x), xi = 0.5f * (A.y - B.y);
        const float yr = 0.5f * (A.y + B.y), yi = 0.5f * (B.x - A.x);
        P[b0 * pstride + k] = xr * xr + xi * xi;
        if (two) P[(b0 + 1) * pstride + k] = yr * yr + yi * yi;
      }
      __syncthreads();
    }
    if constexpr (LOG2ND > 0) {                    // one warp per band, keys in registers
      constexpr int NPL = ((1 << (LOG2ND - 1)) + 1 + 31) / 32;
      for (int b = tid >> 5; b < c.nbands; b += THREADS / 32) {
        double low, tot;
        warp_select_low_sum<NPL>(P + b * pstride, Hd + 1, c.sel_boundary + 1, &low, &tot);
        if ((tid & 31) == 0) coarse[1 + b] = fmin(0.0, 10.0 * log10(low / tot) + (cur_f0 - 100.0) / 50.0);
      }
    } else {
      double low[kMaxSets], tot[kMaxSets];
      select_low_sums<THREADS>(P, pstride, c.nbands, Hd + 1, c.sel_boundary + 1, reinterpret_cast<int*>(cbuf), sc, red, low, tot);
      if (tid < c.nbands)
        coarse[1 + tid] = fmin(0.0, 10.0 * log10(low[tid] / tot[tid]) + (cur_f0 - 100.0) / 50.0);
    }
  }
  write_row();
}


}  // namespace

#ifndef WB_HOST_EMU      // the launcher; tests/emu has its own
bool d4c_run(const UttView& u, int fs, int total_frames, const int* frame_utt,
             const double* frame_t, const double* f0, int fft_size, double threshold,
             double* ap) {
  Context* ctxp = ctx();
  if (!ctxp) return false;
  if (total_frames <= 0) return true;
  cudaStream_t st = ctxp->stream;
  D4CConst c;
  c.fs = fs;
  c.threshold = threshold;
  c.out_half = fft_size / 2;
  const int nd = static_cast<int>(pow(2.0, 1.0 + static_cast<int>(log(4.0 * fs / kFloorF0D4C + 1) / kLog2)));  // d4c.cpp:344-346
  const int nlt = static_cast<int>(pow(2.0, 1.0 + static_cast<int>(log(3.0 * fs / 40.0 + 1) / kLog2)));        // :261-262
  c.log2nd = 0; while ((1 << c.log2nd) < nd) ++c.log2nd;
  c.log2lt = 0; while ((1 << c.log2lt) < nlt) ++c.log2lt;
  if (c.log2nd > 13 || c.log2nd < 6) { set_error("D4C: unsupported sampling rate %d", fs); return false; }
  c.nbands = static_cast<int>(fmin(kUpperLimit, fs / 2.0 - kFrequencyInterval) / kFrequencyInterval);  // :351-353
  if (c.nbands < 0 || c.nbands > kMaxBands || c.nbands > kMaxSets) { set_error("D4C: unsupported number of bands %d (fs %d)", c.nbands, fs); return false; }
  c.window_length = static_cast<int>(kFrequencyInterval * nd / fs) * 2 + 1;                         // :356-357
  c.sel_boundary = matlab_round(nd * 8.0 / c.window_length);                                        // :196-197
  c.band_top = 0;
  for (int i = 0; i < c.nbands; ++i) {
    c.centers[i] = static_cast<int>(kFrequencyInterval * (i + 1) * nd / fs);                        // :204-205
    c.band_top = std::max(c.band_top, c.centers[i] - c.window_length / 2 + c.window_length - 1);
  }
  c.lt_b0 = static_cast<int>(ceil(100.0 * nlt / fs));                                               // :267-269
  c.lt_b1 = static_cast<int>(ceil(4000.0 * nlt / fs));
  c.lt_b2 = static_cast<int>(ceil(7900.0 * nlt / fs));
  if (c.lt_b2 > nlt / 2) c.lt_b2 = nlt / 2;      // the reference reads uninitialised memory here (fs < 15.8 kHz)
  if (c.lt_b1 > c.lt_b2) c.lt_b1 = c.lt_b2;

  std::vector<double> h_win(c.window_length);
  for (int i = 0; i < c.window_length; ++i) {                                                       // common.cpp:113-121
    const double tmp = i / (c.window_length - 1.0);
    h_win[i] = 0.355768 - 0.487396 * cos(2.0 * kPi * tmp) + 0.144232 * cos(4.0 * kPi * tmp) -
               0.012604 * cos(6.0 * kPi * tmp);
  }
  DevBuf<double> d_win, d_ap0;
  DevBuf<long long> counts, offs_lt, offs_main, tot_lt, tot_main;
  if (!d_win.alloc(c.window_length) || !d_ap0.alloc(total_frames) || !counts.alloc(total_frames) ||
      !offs_lt.alloc(total_frames) || !offs_main.alloc(total_frames) || !tot_lt.alloc(u.n_utt) ||
      !tot_main.alloc(u.n_utt))
    return false;
  if (!write_dev(d_win.p, h_win.data(), c.window_length * sizeof(double))) return false;

  const int nblk = (total_frames + 255) / 256;
  std::vector<long long> h_lt(u.n_utt), h_main(u.n_utt);
  // The randn table must reach the largest per-utterance total (LoveTrain draws + three windows per processed
  // frame).  With the frame counts known on the host an upper bound does -- every voiced frame at the lowest F0
  // each part clamps to -- and the two read-backs of the totals (with their host round trips) are not needed.
  const bool bounded = u.max_f_len > 0;
  if (bounded) {
    const long long per_frame = 2LL * (static_cast<long long>(1.5 * fs / 40.0 + 0.5) + 1) + 1 +
                                3LL * (2LL * (static_cast<long long>(2.0 * fs / kFloorF0D4C + 0.5) + 1) + 1);
    if (!ensure_randn((size_t)((long long)u.max_f_len * per_frame))) return false;
  }
  auto need_randn = [&]() {
    if (bounded) return true;
    long long mx = 0;
    for (int i = 0; i < u.n_utt; ++i) mx = std::max(mx, h_lt[i] + h_main[i]);
    return ensure_randn((size_t)mx);
  };
  // LoveTrain
  d4c_lt_count_kernel<<<nblk, 256, 0, st>>>(f0, total_frames, fs, counts.p);
  WB_LAUNCH_CHECK();
  if (!segmented_exclusive_scan(counts.p, u.f_off, u.f_len, u.n_utt, offs_lt.p, tot_lt.p)) return false;
  if (!bounded && !read_back(h_lt.data(), tot_lt.p, u.n_utt * sizeof(long long))) return false;
  std::fill(h_main.begin(), h_main.end(), 0);
  if (!need_randn()) return false;
  {
    const int nl = 1 << c.log2lt;
    const size_t smem = (size_t)(2 * cpad_size(nl / 2) + 96) * sizeof(double);
    KernelTimer kt1("d4c_lovetrain_kernel");
#define WB_LT_LAUNCH(L)                                                                                             \
  do {                                                                                                              \
    WB_CUDA_OR_RETURN(cudaFuncSetAttribute(d4c_lovetrain_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), false); \
    d4c_lovetrain_kernel<L><<<total_frames, 256, smem, st>>>(u, frame_utt, frame_t, f0, offs_lt.p, ctxp->d_randn, L > 0 ? ctxp->tw_c(L > 0 ? L : 4) : ctxp->d_twiddle, c, d_ap0.p); \
  } while (0)
    const bool lt32 = option("lovetrain_fp32") != 0;
    if (lt32 && c.log2lt == 12) {
      const size_t smem32 = (size_t)((cpadf(nl / 2) + 4 + 1) & ~1) * sizeof(float2) + (96 + kLtStage) * sizeof(double);
      WB_CUDA_OR_RETURN(cudaFuncSetAttribute(d4c_lovetrain32_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32), false);
      d4c_lovetrain32_kernel<12><<<total_frames, 256, smem32, st>>>(u, frame_utt, frame_t, f0, offs_lt.p, ctxp->d_randn, ctxp->tw_cf(12), c, d_ap0.p);
    } else
    switch (c.log2lt) {
      case 11: WB_LT_LAUNCH(11); break;
      case 12: WB_LT_LAUNCH(12); break;
      default: WB_LT_LAUNCH(0); break;
    }
#undef WB_LT_LAUNCH
    WB_LAUNCH_CHECK(); kt1.stop();
  }
  // main
  d4c_main_count_kernel<<<nblk, 256, 0, st>>>(f0, d_ap0.p, total_frames, fs, threshold, counts.p);
  WB_LAUNCH_CHECK();
  if (!segmented_exclusive_scan(counts.p, u.f_off, u.f_len, u.n_utt, offs_main.p, tot_main.p)) return false;
  if (!bounded && !read_back(h_main.data(), tot_main.p, u.n_utt * sizeof(long long))) return false;
  if (!need_randn()) return false;
  {
    const int hd = nd / 2;
    const size_t smem = d4c_cbuf_slots(nd, c.nbands) * sizeof(double2) + (size_t)(2 * (hd + 8) + 160) * sizeof(double) +
                        sizeof(SelectScratch) + (kMaxBands + 2) * sizeof(double);
    const int threads = nd > 4096 ? 512 : 256;
    flush_deferred_copies();      // deferred uploads / downloads of a pipelined caller start under this kernel
    KernelTimer kt2("d4c_main_kernel");
#define WB_D4C_LAUNCH(L, TH, ...)                                                                                   \
  do {                                                                                                              \
    WB_CUDA_OR_RETURN(cudaFuncSetAttribute(d4c_main_kernel<L, TH, ##__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), false); \
    d4c_main_kernel<L, TH, ##__VA_ARGS__><<<total_frames, TH, smem, st>>>(u, frame_utt, frame_t, f0, d_ap0.p, offs_main.p, tot_lt.p, ctxp->d_randn, \
                                                           L > 0 ? ctxp->tw_c(L > 0 ? L : 4) : ctxp->d_twiddle, L > 0 ? ctxp->tw_cf(L > 0 ? L : 4) : ctxp->d_twiddle_f, d_win.p, c, ap);                         \
  } while (0)
    if (threads == 512) WB_D4C_LAUNCH(0, 512);
    else if (c.log2nd == 12) WB_D4C_LAUNCH(12, 256, 4);      // radix-16 passes: 3 instead of 4 round trips
    else if (c.log2nd == 11) WB_D4C_LAUNCH(11, 256, 4);
    else WB_D4C_LAUNCH(0, 256);
#undef WB_D4C_LAUNCH
    WB_LAUNCH_CHECK(); kt2.stop();
  }
  WB_CUDA_OR_RETURN(cudaStreamSynchronize(st), false);
  return true;
}
#endif  // WB_HOST_EMU

}  // namespace wb
