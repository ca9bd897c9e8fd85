// world-b200: Synthesis — pulse time base (one CTA per utterance) and per-pulse minimum-phase
// responses with overlap-add (one CTA per pulse).
//
// Reference: W/src/synthesis.cpp — Synthesis :338-397, GetTimeBase :287-320,
// GetTemporalParametersForTimeBase :223-240, GetPulseLocationsForTimeBase :242-285,
// GetOneFrameSegment :183-221, GetSpectralEnvelope :140-157, GetAperiodicRatio :159-178,
// GetPeriodicResponse :105-138, GetAperiodicResponse :38-68, GetNoiseSpectrum :19-33,
// GetSpectrumWithFractionalTimeShift :88-100, RemoveDCComponent :73-82, GetDCRemover :322-334;
// W/src/common.cpp GetMinimumPhaseSpectrum :182-220.
//
// FFT budget per pulse (reference: 7 real/complex transforms of N on one core):
//   C: noise r2c (half-size complex FFT)
//   A: the two log spectra (periodic, aperiodic) are real and even, so one complex FFT of
//      z = L1 + i L2 returns both cepstra as Re Z and Im Z;
//   B: the two folded cepstra again share one complex FFT (split by Hermitian symmetry);
//   D: the two responses come out of one inverse complex FFT of P + i A.
// The noise of pulse i is table[pulse_index[i] - pulse_index[0] + n], so
// the output is independent of pulse scheduling.  Overlap-add uses FP64 atomics (RED.ADD.F64).
#include <stdlib.h>
#include <algorithm>
#include "wb_batch.h"
#include "wb_fft.cuh"

namespace wb {
namespace {

constexpr double kTwoPi = 2.0 * kPi;

struct SynthConst {
  int fs, log2n, f0_max_len;
  double frame_period_s;   // seconds
  double lowest_f0;
};

// interp1 (matlabfunctions.cpp:157-182) on the uniform knot axis x[j] = j * fp, j in [0, n)
__device__ __forceinline__ int uniform_segment(double t, double fp, int n) {
  int g = static_cast<int>(t / fp);
  g = max(0, min(n - 1, g));
  while (g + 1 <= n - 1 && mul_rn((double)(g + 1), fp) <= t) ++g;
  while (g > 0 && mul_rn((double)g, fp) > t) --g;
  if (mul_rn((double)g, fp) > t) return 1;          // t below the first knot
  return max(1, min(n - 1, g + 1));                 // k = clamp(upper_bound, 1, n-1)
}

__device__ __forceinline__ double coarse_f0_at(const double* __restrict__ f0, int n_frames, int j,
                                               double lowest_f0) {
  if (j < n_frames) { const double v = f0[j]; return v < lowest_f0 ? 0.0 : v; }
  const double a = f0[n_frames - 1], b = f0[n_frames - 2];
  const double ca = a < lowest_f0 ? 0.0 : a, cb = b < lowest_f0 ? 0.0 : b;
  return add_rn(mul_rn(ca, 2.0), -cb);                // :236-237
}
__device__ __forceinline__ double coarse_vuv_at(const double* __restrict__ f0, int n_frames, int j,
                                                double lowest_f0) {
  if (j < n_frames) return f0[j] < lowest_f0 || f0[j] == 0.0 ? 0.0 : 1.0;
  const double fa = f0[n_frames - 1], fb = f0[n_frames - 2];
  const double a = fa < lowest_f0 || fa == 0.0 ? 0.0 : 1.0;
  const double b = fb < lowest_f0 || fb == 0.0 ? 0.0 : 1.0;
  return a * 2 - b;                                   // :238-239
}

// Upper bound on the number of pulses of an utterance: a pulse needs the phase to advance by
// 2 pi, and inside a frame interval the interpolated f0 never exceeds the larger knot (or the
// 500 Hz default of unvoiced samples), so pulses <= sum_k max(f_k, f_k+1, 500) * frame_period + 2.
__global__ void synth_pulse_bound_kernel(const double* __restrict__ f0_all, const int* __restrict__ f_off,
                                         const int* __restrict__ f_len, SynthConst c, int* __restrict__ bound) {
  __shared__ double red[96];
  const int u = blockIdx.x;
  const double* __restrict__ f0 = f0_all + f_off[u];
  const int n = f_len[u];
  double v[1] = {0.0};
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const double a = coarse_f0_at(f0, n, k, c.lowest_f0), b = coarse_f0_at(f0, n, k + 1, c.lowest_f0);
    v[0] += fmax(fmax(fabs(a), fabs(b)), kDefaultF0);
  }
  block_sum<1>(v, red);
  if (threadIdx.x == 0) bound[u] = static_cast<int>(v[0] * c.frame_period_s * 1.001) + 4;
}

// ---- time base (GetTimeBase :287-320, GetPulseLocationsForTimeBase :242-285) -------------------------
// Three steps.  Only the middle one is sequential, and it is reduced to the bare running sum:
//   (1) synth_inc_kernel     per sample, fully parallel: interpolated f0 / vuv -> phase increment
//   (2) synth_phase_kernel   one CTA per utterance: total_phase[i] = total_phase[i-1] + inc[i]
//   (3) synth_pulses_kernel  per sample pair, fully parallel: wrap, pulse detection; ordered
//                            compaction by count / scan / write over 1024-sample chunks
// so the latency of an utterance is ~280 light chunk iterations instead of ~1100 heavy ones.
constexpr int kTbChunk = 1024;

__global__ void __launch_bounds__(256)
synth_inc_kernel(const double* __restrict__ f0_all, const int* __restrict__ f_off,
                 const int* __restrict__ f_len, const int* __restrict__ y_len_all,
                 const long long* __restrict__ y_off, SynthConst c, double* __restrict__ inc_all,
                 unsigned char* __restrict__ vuv_all) {
  const int u = blockIdx.y;
  const int y_len = y_len_all[u];
  if (blockIdx.x * kTbChunk >= y_len) return;
  const double* __restrict__ f0 = f0_all + f_off[u];
  const int n_frames = f_len[u];
  const int n_knots = n_frames + 1;
  const double fp = c.frame_period_s;
  const size_t off = (size_t)y_off[u];
  // four samples per thread: four independent chains of exact divisions in flight (a CTA per 256
  // samples was bound by the latency of one chain and by the launch rate of a million tiny CTAs)
#pragma unroll
  for (int q = 0; q < kTbChunk / 256; ++q) {
    const int i = blockIdx.x * kTbChunk + q * 256 + threadIdx.x;
    if (i >= y_len) continue;
    const double t = (double)i / (double)c.fs;                       // :227-228
    const int k = uniform_segment(t, fp, n_knots);
    const double x0 = mul_rn((double)(k - 1), fp), x1 = mul_rn((double)k, fp);
    const double s = div_rn(add_rn(t, -x0), add_rn(x1, -x0));
    const double fa = coarse_f0_at(f0, n_frames, k - 1, c.lowest_f0), fb = coarse_f0_at(f0, n_frames, k, c.lowest_f0);
    const double va = coarse_vuv_at(f0, n_frames, k - 1, c.lowest_f0), vb = coarse_vuv_at(f0, n_frames, k, c.lowest_f0);
    double fi = add_rn(fa, mul_rn(s, add_rn(fb, -fa)));
    const double vi = add_rn(va, mul_rn(s, add_rn(vb, -va)));
    const unsigned char vuv = vi > 0.5 ? 1 : 0;                       // :305-309
    if (!vuv) fi = kDefaultF0;
    inc_all[off + i] = div_rn(mul_rn(kTwoPi, fi), (double)c.fs);      // :250,253
    vuv_all[off + i] = vuv;
  }
}

// total_phase[i] = total_phase[i-1] + inc[i] (:252-253) must be accumulated in exactly the
// reference's order: with fs / 500 Hz an integer (48 kHz, 16 kHz) every unvoiced pulse sits on a
// wrap-around that is decided by the rounding of this running sum, so a tree-shaped floating-point
// scan moves pulses by one sample.
// Exact parallel form: while the running sum s stays inside one binade [2^e, 2^(e+1)], every
// IEEE addition is  s <- s + rn_u(a)  with u = ulp = 2^(e-52) and rn_u = round-to-nearest
// multiple of u (s is a multiple of u, so the rounding does not depend on s unless a / u ends
// in exactly .5).  In units of u the chunk is then an INTEGER prefix sum -- associative, so a
// warp-shuffle scan reproduces the sequential result bit for bit.  Chunks that see a tie, leave
// the binade (the sum doubles ~20 times per utterance) or start from 0 take a sequential walk
// by one thread (loads in batches of eight, only the dependent DADDs are serial).
__global__ void __launch_bounds__(kTbChunk)
synth_phase_kernel(const double* __restrict__ inc_all, const long long* __restrict__ y_off,
                   const int* __restrict__ y_len_all, double* __restrict__ tot_all) {
  __shared__ double inc_s[kTbChunk];
  __shared__ double tot_s[kTbChunk];
  // double-buffered by chunk parity: a chunk's values are read after its second barrier while the
  // next chunk already writes the other set before its first one -- two barriers per chunk suffice
  __shared__ long long wsum_s[2][32];     // inclusive warp totals, then (warp 0) exclusive warp offsets
  __shared__ int wslow_s[2][32];
  __shared__ long long chunk_total_s[2];
  __shared__ int chunk_slow_s[2];
  const int u = blockIdx.x;
  const int y_len = y_len_all[u];
  const size_t off = (size_t)y_off[u];
  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = T >> 5;
  double carry = 0.0;                      // running sum before this chunk, the same in every thread
  double inc_next = tid < y_len ? inc_all[off + tid] : 0.0;
  int par = 0;
  for (int base = 0; base < y_len; base += T, par ^= 1) {
    const int i = base + tid;
    const double inc = inc_next;
    if (i + T < y_len) inc_next = inc_all[off + i + T];          // the next chunk's load is in flight during this one
    else inc_next = 0.0;
    const double s0 = carry;
    bool slow = !(s0 > 0.0);
    long long r = 0, s0int = 0;
    int e = 0;
    if (!slow) {
      e = ilogb(s0);
      const double m = scalbn(inc, 52 - e);
      const double rr = rint(m);
      // a negative increment (the extrapolated last F0 knot can be negative while the interpolated V/UV
      // flag is still voiced) breaks the monotonicity the end-of-chunk binade test below relies on
      slow = fabs(m - rr) == 0.5 || !(m < 4503599627370496.0) || inc < 0.0;
      r = static_cast<long long>(rr);
      s0int = static_cast<long long>(scalbn(s0, 52 - e));
    }
    long long pre = r;                                   // inclusive scan inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += t;
    }
    const bool warp_slow = __any_sync(0xffffffffu, slow);
    if (lane == 31) { wsum_s[par][wid] = pre; wslow_s[par][wid] = warp_slow; }
    __syncthreads();
    if (wid == 0) {                                      // scan of the warp totals, chunk total, slow flag
      const long long wv = lane < nw ? wsum_s[par][lane] : 0;
      long long winc = wv;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      const bool any_slow = __any_sync(0xffffffffu, lane < nw && wslow_s[par][lane] != 0);
      wsum_s[par][lane] = winc - wv;                     // exclusive offset of warp `lane`
      if (lane == 31) {
        chunk_total_s[par] = winc;
        // the sums are non-decreasing (inc >= 0): the chunk stays inside the binade iff its end does
        chunk_slow_s[par] = any_slow || (s0int + winc > 9007199254740992LL) || (s0int + winc < 4503599627370496LL);
      }
    }
    __syncthreads();
    double total;
    if (chunk_slow_s[par]) {                             // block-uniform
      inc_s[tid] = inc;
      __syncthreads();
      if (tid == 0) {
        double run = carry;
        for (int q = 0; q < T; q += 8) {
          double v[8];
#pragma unroll
          for (int r8 = 0; r8 < 8; ++r8) v[r8] = inc_s[q + r8];
#pragma unroll
          for (int r8 = 0; r8 < 8; ++r8) { run = add_rn(run, v[r8]); v[r8] = run; }
#pragma unroll
          for (int r8 = 0; r8 < 8; ++r8) tot_s[q + r8] = v[r8];
        }
      }
      __syncthreads();
      total = tot_s[tid];
      carry = tot_s[T - 1];
      __syncthreads();                                   // tot_s / inc_s are rewritten by the next slow chunk
    } else {
      total = scalbn(static_cast<double>(s0int + pre + wsum_s[par][wid]), e - 52);
      carry = scalbn(static_cast<double>(s0int + chunk_total_s[par]), e - 52);
    }
    if (i < y_len) tot_all[off + i] = total;
  }
}

// The same running sum with FOUR consecutive samples per thread (256 threads per chunk of 1 024): the
// thread sums its own four increments first, so the two scans handle a quarter of the values, a CTA is 8
// warps instead of 32 (cheaper barriers; every utterance of a 1 132-utterance batch is resident at once
// instead of in four waves) and the powers of two are applied by exact multiplications instead of scalbn().
// Bit-identical to synth_phase_kernel by construction: the same integer prefix sum in units of ulp(carry).
constexpr int kTbPer = 4;
__global__ void __launch_bounds__(kTbChunk / kTbPer)
synth_phase4_kernel(const double* __restrict__ inc_all, const long long* __restrict__ y_off,
                    const int* __restrict__ y_len_all, double* __restrict__ tot_all) {
  constexpr int T = kTbChunk / kTbPer, NW = T / 32;
  __shared__ double inc_s[kTbChunk];
  __shared__ double tot_s[kTbChunk];
  __shared__ long long wsum_s[2][NW];      // double-buffered by chunk parity (two barriers per chunk suffice)
  __shared__ int wslow_s[2][NW];
  const int u = blockIdx.x;
  const int y_len = y_len_all[u];
  const size_t off = (size_t)y_off[u];      // even: 16-byte aligned rows
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  double carry = 0.0;                       // running sum before this chunk, the same in every thread
  auto load4 = [&](int i0, double* v) {
    if (i0 + kTbPer <= y_len) {
      const double2 a = *reinterpret_cast<const double2*>(inc_all + off + i0);
      const double2 b = *reinterpret_cast<const double2*>(inc_all + off + i0 + 2);
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
#pragma unroll
      for (int j = 0; j < kTbPer; ++j) v[j] = i0 + j < y_len ? inc_all[off + i0 + j] : 0.0;
    }
  };
  double nxt[kTbPer];
  load4(kTbPer * tid, nxt);
  int par = 0;
  for (int base = 0; base < y_len; base += kTbChunk, par ^= 1) {
    const int i0 = base + kTbPer * tid;
    double inc[kTbPer];
#pragma unroll
    for (int j = 0; j < kTbPer; ++j) inc[j] = nxt[j];
    if (base + kTbChunk < y_len) load4(i0 + kTbChunk, nxt);          // the next chunk's load is in flight during this one
    const double s0 = carry;
    bool slow = !(s0 > 0.0);
    long long p[kTbPer] = {0, 0, 0, 0}, s0int = 0;
    double down = 0.0;                                               // 2^(e-52)
    if (!slow) {
      const int e = ((__double2hiint(s0) >> 20) & 0x7ff) - 1023;     // ilogb of a positive normal number
      const double up = __hiloint2double((1023 + 52 - e) << 20, 0);  // 2^(52-e), exact
      down = __hiloint2double((1023 + e - 52) << 20, 0);
      long long run = 0;
#pragma unroll
      for (int j = 0; j < kTbPer; ++j) {
        const double m = inc[j] * up;                                // exact (a power of two, no underflow here)
        const double rr = rint(m);
        slow = slow || fabs(m - rr) == 0.5 || !(m < 4503599627370496.0) || inc[j] < 0.0;
        run += static_cast<long long>(rr);
        p[j] = run;
      }
      s0int = static_cast<long long>(s0 * up);
    }
    long long pre = p[kTbPer - 1];                                   // inclusive scan of the thread totals inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += t;
    }
    const bool warp_slow = __any_sync(0xffffffffu, slow);
    if (lane == 31) { wsum_s[par][wid] = pre; wslow_s[par][wid] = warp_slow; }
    __syncthreads();
    long long woff = 0, chunk_total = 0;
    bool any_slow = false;
#pragma unroll
    for (int w = 0; w < NW; ++w) {                                   // 8 warp totals: every thread adds them itself
      const long long v = wsum_s[par][w];
      if (w < wid) woff += v;
      chunk_total += v;
      any_slow = any_slow || wslow_s[par][w] != 0;
    }
    // the sums are non-decreasing (inc >= 0): the chunk stays inside the binade iff its end does
    const bool chunk_slow = any_slow || (s0int + chunk_total > 9007199254740992LL) || (s0int + chunk_total < 4503599627370496LL);
    double total[kTbPer];
    if (chunk_slow) {                                                // block-uniform
#pragma unroll
      for (int j = 0; j < kTbPer; ++j) inc_s[kTbPer * tid + j] = inc[j];
      __syncthreads();
      if (tid == 0) {
        double run = carry;
        for (int q = 0; q < kTbChunk; q += 8) {
          double v[8];
#pragma unroll
          for (int r8 = 0; r8 < 8; ++r8) v[r8] = inc_s[q + r8];
#pragma unroll
          for (int r8 = 0; r8 < 8; ++r8) { run = add_rn(run, v[r8]); v[r8] = run; }
#pragma unroll
          for (int r8 = 0; r8 < 8; ++r8) tot_s[q + r8] = v[r8];
        }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < kTbPer; ++j) total[j] = tot_s[kTbPer * tid + j];
      carry = tot_s[kTbChunk - 1];
      __syncthreads();                                               // tot_s / inc_s are rewritten by the next slow chunk
    } else {
      const long long before = s0int + woff + (pre - p[kTbPer - 1]); // everything in front of this thread's samples
#pragma unroll
      for (int j = 0; j < kTbPer; ++j) total[j] = static_cast<double>(before + p[j]) * down;
      carry = static_cast<double>(s0int + chunk_total) * down;
    }
    if (i0 + kTbPer <= y_len) {
      *reinterpret_cast<double2*>(tot_all + off + i0) = make_double2(total[0], total[1]);
      *reinterpret_cast<double2*>(tot_all + off + i0 + 2) = make_double2(total[2], total[3]);
    } else {
#pragma unroll
      for (int j = 0; j < kTbPer; ++j) if (i0 + j < y_len) tot_all[off + i0 + j] = total[j];
    }
  }
}

// fmod(x, 2 pi) for 0 <= x < 2^20, exact like the C library's: with n = floor(x / 2 pi) (possibly off
// by one) the residual x - n * y is a multiple of ulp(y) = 2^-50 and smaller than 8 in magnitude,
// so it is representable and the FMA delivers it without rounding; one exact +- y repairs n.
__device__ __forceinline__ double fmod_two_pi(double x) {
  if (!(x >= 0.0 && x < 1048576.0)) return fmod(x, kTwoPi);
  const double n = floor(x * (1.0 / kTwoPi));
  double r = fma(-n, kTwoPi, x);
  if (r < 0.0) r += kTwoPi;
  else if (r >= kTwoPi) r -= kTwoPi;
  return r;
}

// Pulses: sample pair (j, j+1) holds one when the wrapped phase jumps by more than pi (:255-259).
// One CTA per 1024-pair chunk of one utterance.  WRITE = false stores the chunk's pulse count;
// after synth_pulse_scan_kernel turned the counts into exclusive offsets, WRITE = true stores the
// pulses in sample order.
template <bool WRITE>
__global__ void __launch_bounds__(256)
synth_pulses_kernel(const double* __restrict__ tot_all, const unsigned char* __restrict__ vuv_all,
                    const long long* __restrict__ y_off, const int* __restrict__ y_len_all, SynthConst c,
                    int n_chunks_max, int* __restrict__ counts, const int* __restrict__ pulse_off,
                    const int* __restrict__ pulse_cap, int* __restrict__ p_index,
                    double* __restrict__ p_shift, unsigned char* __restrict__ p_vuv, int* __restrict__ p_utt) {
  __shared__ double first_s[8 + 1];         // wrapped phase of the first sample of every warp (+ of the next chunk)
  __shared__ int wcnt[8];
  const int u = blockIdx.y, chunk = blockIdx.x;
  const int y_len = y_len_all[u];
  const int j0 = chunk * kTbChunk;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (j0 + 1 >= y_len) {                                  // no pair starts in this chunk
    if (!WRITE && tid == 0) counts[(size_t)u * n_chunks_max + chunk] = 0;
    return;
  }
  const size_t off = (size_t)y_off[u];
  const int jb = j0 + 4 * tid;                            // this thread: samples jb .. jb + 3 (and the pair into jb + 4)
  double w[5];
  if (jb + 3 < y_len) {                                   // 32-byte aligned: utterance offsets are even, jb is a multiple of 4
    const double2 a = *reinterpret_cast<const double2*>(tot_all + off + jb);
    const double2 b2 = *reinterpret_cast<const double2*>(tot_all + off + jb + 2);
    w[0] = fmod_two_pi(a.x); w[1] = fmod_two_pi(a.y); w[2] = fmod_two_pi(b2.x); w[3] = fmod_two_pi(b2.y);   // :251,254
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) w[q] = jb + q < y_len ? fmod_two_pi(tot_all[off + jb + q]) : 0.0;
  }
  if (lane == 0) first_s[wid] = w[0];
  if (tid == 255) first_s[8] = j0 + kTbChunk < y_len ? fmod_two_pi(tot_all[off + j0 + kTbChunk]) : 0.0;
  w[4] = __shfl_down_sync(0xffffffffu, w[0], 1);
  __syncthreads();
  if (lane == 31) w[4] = first_s[wid + 1];
  unsigned mask = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (jb + q + 1 < y_len && fabs(w[q + 1] - w[q]) > kPi) mask |= 1u << q;                                  // :255,259
  const int cnt = __popc(mask);
  int inc_scan = cnt;                                     // inclusive scan of the per-thread counts inside the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc_scan, o);
    if (lane >= o) inc_scan += t;
  }
  if (lane == 31) wcnt[wid] = inc_scan;
  __syncthreads();
  if (!WRITE) {
    if (tid == 0) {
      int tot = 0;
#pragma unroll
      for (int wq = 0; wq < 8; ++wq) tot += wcnt[wq];
      counts[(size_t)u * n_chunks_max + chunk] = tot;
    }
    return;
  }
  if (mask == 0) return;
  int pos = counts[(size_t)u * n_chunks_max + chunk] + (inc_scan - cnt);
  for (int wq = 0; wq < wid; ++wq) pos += wcnt[wq];
  const int cap = pulse_cap[u], base_out = pulse_off[u];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (!((mask >> q) & 1u)) continue;
    if (pos < cap) {
      const int o = base_out + pos, j = jb + q;
      p_index[o] = j;
      const double yy1 = w[q] - kTwoPi;                                   // :271-274
      const double xx = -yy1 / (w[q + 1] - yy1);
      p_shift[o] = xx / c.fs;
      p_vuv[o] = vuv_all[off + j];
      p_utt[o] = u;
    }
    ++pos;
  }
}

// one thread per utterance: exclusive scan of the chunk counts in place, total -> pulse_count
__global__ void synth_pulse_scan_kernel(int* __restrict__ counts, int n_utt, int n_chunks_max,
                                        int* __restrict__ pulse_count) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_utt) return;
  int* cu = counts + (size_t)u * n_chunks_max;
  int acc = 0;
  for (int k = 0; k < n_chunks_max; ++k) { const int v = cu[k]; cu[k] = acc; acc += v; }
  pulse_count[u] = acc;
}

// ---- pulse classification -------------------------------------------------------------------
// A pulse has a periodic response when it is voiced and its aperiodicity at DC is <= 0.999
// (GetPeriodicResponse :110).  Pulses are split into two work lists: "periodic" pulses need
// two minimum-phase spectra (periodic + aperiodic part), all others need one.  The pulse
// kernel transforms two real sequences per complex FFT, so a periodic pulse is one work item
// and two non-periodic pulses share one.
struct PulseFrames { int fr_floor, fr_ceil; double interp; };
__device__ __forceinline__ PulseFrames pulse_frames(int index, int fs, double frame_period_s, int n_frames) {
  const double current_time = (double)index / (double)fs;
  const double pos_f = current_time / frame_period_s;                 // :145-146, :164-165
  PulseFrames r;
  r.fr_floor = min(n_frames - 1, static_cast<int>(floor(pos_f)));
  r.fr_ceil = min(n_frames - 1, static_cast<int>(ceil(pos_f)));
  r.interp = pos_f - r.fr_floor;
  return r;
}
__device__ __forceinline__ double safe_ap(double x) { return fmax(0.001, fmin(0.999999999999, x)); }   // common.h:111-113

// Work lists in PULSE ORDER (count pass, scan, write pass): the two non-periodic pulses that share one
// complex transform are always the same two, so the waveform does not depend on which thread reached an
// atomic counter first (the partner's rounding noise leaks into a channel at the 1e-7 level).
// blk_cnt: [2][n_blocks] -- count pass: events of each kind per CTA; write pass: exclusive offsets.
template <bool WRITE>
__global__ void __launch_bounds__(256)
synth_classify_kernel(const double* __restrict__ ap_all, const int* __restrict__ f_off,
                      const int* __restrict__ f_len, const int* __restrict__ p_index,
                      const unsigned char* __restrict__ p_vuv, const int* __restrict__ p_utt,
                      int total_p, SynthConst c, int* __restrict__ blk_cnt,
                      int* __restrict__ list_per, int* __restrict__ list_aper) {
  __shared__ int wc[2][8];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int kind = 0;                                     // 0: unused slot, 1: periodic, 2: non-periodic
  if (p < total_p && p_utt[p] >= 0) {
    bool periodic = false;
    if (p_vuv[p]) {
      const int u = p_utt[p];
      const int half = (1 << c.log2n) >> 1;
      const PulseFrames fr = pulse_frames(p_index[p], c.fs, c.frame_period_s, f_len[u]);
      const double a0 = safe_ap(ap_all[((size_t)f_off[u] + fr.fr_floor) * (half + 1)]);
      double ar;
      if (fr.fr_floor == fr.fr_ceil) ar = a0 * a0;
      else {
        const double a1 = safe_ap(ap_all[((size_t)f_off[u] + fr.fr_ceil) * (half + 1)]);
        const double m = add_rn(mul_rn(1.0 - fr.interp, a0), mul_rn(fr.interp, a1));
        ar = m * m;
      }
      periodic = !(ar > 0.999);
    }
    kind = periodic ? 1 : 2;
  }
  const unsigned m1 = __ballot_sync(0xffffffffu, kind == 1), m2 = __ballot_sync(0xffffffffu, kind == 2);
  if (lane == 0) { wc[0][wid] = __popc(m1); wc[1][wid] = __popc(m2); }
  __syncthreads();
  if (!WRITE) {
    if (threadIdx.x < 2) {
      int t = 0;
      for (int w = 0; w < 8; ++w) t += wc[threadIdx.x][w];
      blk_cnt[threadIdx.x * gridDim.x + blockIdx.x] = t;
    }
    return;
  }
  if (kind == 0) return;
  const int k = kind - 1;
  int pos = blk_cnt[k * gridDim.x + blockIdx.x] + __popc((k == 0 ? m1 : m2) & ((1u << lane) - 1u));
  for (int w = 0; w < wid; ++w) pos += wc[k][w];
  (k == 0 ? list_per : list_aper)[pos] = p;
}

// exclusive scan of the two rows of blk_cnt (one warp per row), totals -> cnt2[0..1]
__global__ void synth_classify_scan_kernel(int* __restrict__ blk_cnt, int n_blocks, int* __restrict__ cnt2) {
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* row = blk_cnt + (size_t)k * n_blocks;
  int carry = 0;
  for (int i0 = 0; i0 < n_blocks; i0 += 32) {
    const int i = i0 + lane;
    const int v = i < n_blocks ? row[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (i < n_blocks) row[i] = carry + inc - v;
    carry += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) cnt2[k] = carry;
}

// ---- one work item: two minimum-phase responses through four complex FFTs -----------------
// Channel 0 / 1 of a periodic item: the periodic and the aperiodic response of one pulse.
// Channel 0 / 1 of a pair item: the aperiodic responses of two non-periodic pulses.
//   C: noise of both channels, n0 + i n1            -> spectra by Hermitian split
//   A: even extensions of the log spectra, L0 + i L1 -> both cepstra (Re, Im)
//   B: folded cepstra (common.cpp:194-206)           -> both analytic log spectra (split)
//   D: inverse of Y0 + i Y1 (Hermitian extensions)   -> both responses (Re, Im)
// C = float2 by default: every quantity here is a log spectrum, a cepstrum, unit-variance
// noise or a response that is accumulated into y, and 2^-24 relative to the largest element
// leaves the resynthesis > 90 dB above the error (tolerance: 60 dB).  Pulse positions, the
// interpolation of sp / ap and the overlap-add itself stay in FP64.
// dynamic shared memory: [ cbuf: cpad_size(N) C | nzb: cpad_size(N) C | red: 96 doubles ]
struct ChanInfo {
  int p, index, noise_raw, noise_size, utt, y_len, fr_floor, fr_ceil;
  bool vuv;
  double interp;
  size_t row0;
  const uint32_t* rn;
};

__device__ __forceinline__ ChanInfo chan_info(int p, const int* __restrict__ p_index, const int* __restrict__ p_utt,
                                              const unsigned char* __restrict__ p_vuv,
                                              const int* __restrict__ pulse_off, const int* __restrict__ pulse_cnt,
                                              const int* __restrict__ f_off, const int* __restrict__ f_len,
                                              const int* __restrict__ y_len_all, const uint32_t* __restrict__ randn_tab,
                                              const SynthConst& c, int N) {
  ChanInfo ci;
  ci.p = p;
  if (p < 0) { ci.index = 0; ci.noise_raw = 0; ci.noise_size = 0; ci.utt = 0; ci.y_len = 0; ci.fr_floor = ci.fr_ceil = 0;
               ci.vuv = false; ci.interp = 0.0; ci.row0 = 0; ci.rn = randn_tab; return ci; }
  const int u = p_utt[p];
  const int first = pulse_off[u], last = first + pulse_cnt[u] - 1;
  ci.utt = u;
  ci.index = p_index[p];
  ci.noise_raw = p_index[min(last, p + 1)] - ci.index;                    // :370-371
  ci.noise_size = min(ci.noise_raw, N);                                   // memory guard
  ci.vuv = p_vuv[p] != 0;
  ci.y_len = y_len_all[u];
  const PulseFrames fr = pulse_frames(ci.index, c.fs, c.frame_period_s, f_len[u]);
  ci.fr_floor = fr.fr_floor; ci.fr_ceil = fr.fr_ceil; ci.interp = fr.interp;
  ci.row0 = (size_t)f_off[u];
  ci.rn = randn_tab + (ci.index - p_index[first]);
  return ci;
}

// interpolated spectral envelope and squared aperiodicity of bin k (:140-178)
__device__ __forceinline__ void envelope_at(const ChanInfo& ci, const double* __restrict__ sp_all,
                                            const double* __restrict__ ap_all, int half, int k, double* s_out,
                                            double* a_out) {
  const size_t r0 = (ci.row0 + ci.fr_floor) * (half + 1) + k, r1 = (ci.row0 + ci.fr_ceil) * (half + 1) + k;
  const double a0 = safe_ap(ap_all[r0]);
  if (ci.fr_floor == ci.fr_ceil) {
    *s_out = fabs(sp_all[r0]);
    *a_out = a0 * a0;
  } else {
    const double a1 = safe_ap(ap_all[r1]);
    *s_out = add_rn(mul_rn(1.0 - ci.interp, fabs(sp_all[r0])), mul_rn(ci.interp, fabs(sp_all[r1])));
    const double m = add_rn(mul_rn(1.0 - ci.interp, a0), mul_rn(ci.interp, a1));
    *a_out = m * m;
  }
}

__device__ __forceinline__ void exp_sincos(double re, double im, double* er, double* ei) {
  const double e = exp(re); double sn, cs; sincos(im, &sn, &cs); *er = e * cs; *ei = e * sn;
}
// FP32 channel: hardware approximations (MUFU.EX2 / SIN / COS / LG2, ~2^-21 relative); |im| is a
// minimum-phase angle of a few radians at most, far inside the accurate range of __sincosf.  The
// resynthesis stays > 100 dB above the error (tests: tolerance 60 dB).
__device__ __forceinline__ void exp_sincos(float re, float im, float* er, float* ei) {
  const float e = __expf(re); float sn, cs; __sincosf(im, &sn, &cs); *er = e * cs; *ei = e * sn;
}
__device__ __forceinline__ double log_of(double v, double) { return log(v); }
__device__ __forceinline__ float log_of(double v, float) { return __logf(static_cast<float>(v)); }

// Overlap-add (W/src/synthesis.cpp:376-383) with a result that does not depend on the order in which the
// work items reach a sample: every response value is rounded once to a multiple of 2^-40 and accumulated
// as a 64-bit integer (integer addition is associative, an atomic double add is not), and one pass at the
// end converts the sums back.  2^-40 = 9e-13 is 20 dB below the rounding of the FP32 transforms even for a
// recording at -120 dB; the sums hold |y| < 2^23 (double samples on a 16-bit scale fit), values beyond
// saturate.
constexpr double kOlaScale = 1099511627776.0;            // 2^40
__device__ __forceinline__ void ola_add(double* y, int i, double r) {
  const double q = fmin(fmax(r * kOlaScale, -4.6e18), 4.6e18);
  atomicAdd(reinterpret_cast<unsigned long long*>(y) + i, static_cast<unsigned long long>(__double2ll_rn(q)));
}
__global__ void synth_ola_finish_kernel(double* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = static_cast<double>(reinterpret_cast<const long long*>(y)[i]) * (1.0 / kOlaScale);
}

#ifndef WB_SYNTH_MAXK
#define WB_SYNTH_MAXK 4
#endif
constexpr int kSynthMaxK = WB_SYNTH_MAXK;      // radix 2^k of the items' transform passes (experiments: build.py --variant)
template <int LOG2N, typename C, int THREADS = 256, int MINB = (sizeof(C) == 8 ? 1024 : 768) / THREADS>      // LOG2N 0: size given at run time (c.log2n)
__global__ void __launch_bounds__(THREADS, MINB)
synth_item_kernel(const double* __restrict__ sp_all, const double* __restrict__ ap_all,
                  const int* __restrict__ f_off, const int* __restrict__ f_len,
                  const long long* __restrict__ y_off, const int* __restrict__ y_len_all,
                  const int* __restrict__ pulse_off, const int* __restrict__ pulse_cnt,
                  const int* __restrict__ p_index, const double* __restrict__ p_shift,
                  const unsigned char* __restrict__ p_vuv, const int* __restrict__ p_utt,
                  const int* __restrict__ list_per, const int* __restrict__ list_aper, int n_per, int n_aper,
                  const uint32_t* __restrict__ randn_tab, const C* __restrict__ tw,
                  const double* __restrict__ dc_remover, SynthConst c, double* __restrict__ y_all) {
  using R = scalar_t<C>;
  constexpr int TWL = LOG2N > 0 ? LOG2N : kTwLog2;       // compact twiddle table of this size, or the master table
  WB_DYN_SMEM(double2, smem2);
  const int log2n = LOG2N > 0 ? LOG2N : c.log2n;
  const int N = 1 << log2n, half = N >> 1;
  constexpr int T = THREADS;
  constexpr int kQ = LOG2N > 0 ? ((1 << (LOG2N > 0 ? LOG2N - 1 : 0)) + THREADS) / THREADS : 2304 / THREADS;   // ceil((N/2 + 1) / T)
  C* cbuf = reinterpret_cast<C*>(smem2);
  C* nzb = cbuf + cpad_size(N);
  double* red = reinterpret_cast<double*>(nzb + cpad_size(N));
  const int tid = threadIdx.x;
  const int item = blockIdx.x;
  const bool per_item = item < n_per;
  int pa, pb;
  if (per_item) { pa = pb = list_per[item]; }
  else {
    const int j = 2 * (item - n_per);
    if (j >= n_aper) return;
    pa = list_aper[j];
    pb = j + 1 < n_aper ? list_aper[j + 1] : -1;
  }
  const ChanInfo c0 = chan_info(pa, p_index, p_utt, p_vuv, pulse_off, pulse_cnt, f_off, f_len, y_len_all, randn_tab, c, N);
  const ChanInfo c1 = per_item ? c0 : chan_info(pb, p_index, p_utt, p_vuv, pulse_off, pulse_cnt, f_off, f_len, y_len_all, randn_tab, c, N);

  // The sp / ap rows of the pulse(s) are needed first; ask L2 for them now and transform the
  // noise (which only needs the randn table) while they are in flight.
  {
    const int lines = ((half + 1) * 8 + 127) / 128;
    const int n_rows = (per_item || pb < 0) ? 4 : 8;
    for (int i = tid; i < n_rows * lines; i += T) {
      const int r = i / lines;
      const ChanInfo& ci = r < 4 ? c0 : c1;
      const size_t row = ci.row0 + ((r & 1) ? ci.fr_ceil : ci.fr_floor);
      const char* ptr = reinterpret_cast<const char*>(((r & 2) ? ap_all : sp_all) + row * (half + 1)) + (size_t)(i % lines) * 128;
#ifndef WB_HOST_EMU
      asm volatile("prefetch.global.L2 [%0];" :: "l"(ptr));      // (an L1 prefetch measured the same)
#endif
    }
  }
  // ---- noise of both channels (:19-33) -----------------------------------------------------------
  {
    double s2[2] = {0.0, 0.0};
    if (!per_item) for (int i = tid; i < c0.noise_size; i += T) s2[0] += randn_from_u32(c0.rn[i]);
    for (int i = tid; i < c1.noise_size; i += T) s2[1] += randn_from_u32(c1.rn[i]);
    block_sum<2>(s2, red);
    const double av0 = s2[0] / c0.noise_raw, av1 = s2[1] / c1.noise_raw;
    for (int i = tid; i < N; i += T) {
      const R n0 = (!per_item && i < c0.noise_size) ? static_cast<R>(randn_from_u32(c0.rn[i]) - av0) : static_cast<R>(0);
      const R n1 = i < c1.noise_size ? static_cast<R>(randn_from_u32(c1.rn[i]) - av1) : static_cast<R>(0);
      nzb[cpadT<C>(brev(i, log2n))] = mk2(n0, n1);
    }
  }
  fft_dit<LOG2N, false, T, kSynthMaxK, TWL>(nzb, log2n, tw);       // C
  // ---- log spectra (:45-51, :115-117), written as the even extension in bit-reversed order ----
  for (int k = tid; k <= half; k += T) {
    R l0 = 0, l1 = 0;
    double s, a;
    envelope_at(c0, sp_all, ap_all, half, k, &s, &a);
    if (per_item) {
      l0 = log_of(s * (1.0 - a) + kMySafeGuardMinimum, R()) / 2;
      l1 = log_of(s * a, R()) / 2;
    } else {
      l0 = log_of(c0.vuv ? s * a : s, R()) / 2;
      if (pb >= 0) {
        envelope_at(c1, sp_all, ap_all, half, k, &s, &a);
        l1 = log_of(c1.vuv ? s * a : s, R()) / 2;
      }
    }
    const C z = mk2(l0, l1);
    cbuf[cpadT<C>(brev(k, log2n))] = z;
    if (k > 0 && k < half) cbuf[cpadT<C>(brev(N - k, log2n))] = z;
  }
  fft_dit<LOG2N, false, T, kSynthMaxK, TWL>(cbuf, log2n, tw);      // A
  // ---- fold the cepstra (common.cpp:194-206) -----------------------------------------------------
  {
    C keep[kQ];
#pragma unroll
    for (int q = 0; q < kQ; ++q) {
      const int i = tid + q * T;
      C z = mk2(static_cast<R>(0), static_cast<R>(0));
      if (i <= half) {
        z = cbuf[cpadT<C>(i)];
        if (i > 0 && i < half) { z.x *= 2; z.y *= 2; }
      }
      keep[q] = z;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kQ; ++q) {
      const int i = tid + q * T;
      if (i <= half) cbuf[cpadT<C>(brev(i, log2n))] = keep[q];
    }
    for (int i = half + 1 + tid; i < N; i += T) cbuf[cpadT<C>(brev(i, log2n))] = mk2(static_cast<R>(0), static_cast<R>(0));
  }
  fft_dit<LOG2N, false, T, kSynthMaxK, TWL>(cbuf, log2n, tw);      // B
  // ---- minimum-phase spectra, time shift / noise product (:56-65, :88-100, :120-131) ------------
  {
    const double coefficient = per_item
        ? div_rn(mul_rn(mul_rn(kTwoPi, p_shift[pa]), (double)c.fs), (double)N) : 0.0;     // :128-129
    const R inv_n = static_cast<R>(1.0 / N), hlf = static_cast<R>(0.5);
    C y0[kQ], y1[kQ];
#pragma unroll
    for (int q = 0; q < kQ; ++q) {
      const int k = tid + q * T;
      y0[q] = mk2(static_cast<R>(0), static_cast<R>(0));
      y1[q] = y0[q];
      if (k > half) continue;
      const int km = (N - k) & (N - 1);
      const C A = cbuf[cpadT<C>(k)], B = cbuf[cpadT<C>(km)];
      const C Zn = nzb[cpadT<C>(k)], Zm = nzb[cpadT<C>(km)];
      // S0 = (A + conj B)/2, S1 = (A - conj B)/(2i); likewise for the noise spectra
      const R s0r = hlf * (A.x + B.x), s0i = hlf * (A.y - B.y);
      const R s1r = hlf * (A.y + B.y), s1i = hlf * (B.x - A.x);
      const C X0 = mk2(hlf * (Zn.x + Zm.x), hlf * (Zn.y - Zm.y));
      const C X1 = mk2(hlf * (Zn.y + Zm.y), hlf * (Zm.x - Zn.x));
      C m0, m1;
      exp_sincos(s0r * inv_n, s0i * inv_n, &m0.x, &m0.y);
      exp_sincos(s1r * inv_n, s1i * inv_n, &m1.x, &m1.y);
      if (per_item) {
        // cos(c k) - j sqrt(1 - cos^2(c k))  (:93-96; the square root is |sin|)
        if constexpr (sizeof(R) == 4) {
          // FP32 channel: the angle c k lies in [0, pi]; in units of pi it is exact to 6e-8 as a float
          // and sincospif needs no range reduction (the phasor error, ~1e-7, is 20 dB below the
          // rounding of the FP32 transforms around it)
          float sn, cs;
          sincospif(static_cast<float>(coefficient * k * (1.0 / kPi)), &sn, &cs);
          y0[q] = cmul(m0, mk2(cs, -fabsf(sn)));
        } else {
          double sn, cs;
          sincos(coefficient * k, &sn, &cs);
          y0[q] = cmul(m0, mk2(static_cast<R>(cs), static_cast<R>(-fabs(sn))));
        }
      } else {
        y0[q] = cmul(m0, X0);
      }
      y1[q] = (per_item || pb >= 0) ? cmul(m1, X1) : mk2(static_cast<R>(0), static_cast<R>(0));
    }
    __syncthreads();
    // D input: Y0 + i Y1 at bin k, conj(Y0) + i conj(Y1) at bin N - k; Im of bins 0 and N/2
    // is ignored by the reference's c2r (W/src/fft.cpp:27-34)
#pragma unroll
    for (int q = 0; q < kQ; ++q) {
      const int k = tid + q * T;
      if (k > half) continue;
      C a = y0[q], b = y1[q];
      if (k == 0 || k == half) { a.y = 0; b.y = 0; }
      cbuf[cpadT<C>(brev(k, log2n))] = mk2(a.x - b.y, a.y + b.x);
      if (k > 0 && k < half) cbuf[cpadT<C>(brev(N - k, log2n))] = mk2(a.x + b.y, b.x - a.y);
    }
  }
  fft_dit<LOG2N, true, T, kSynthMaxK, TWL>(cbuf, log2n, tw);       // D
  // ---- fftshift, RemoveDCComponent (:73-82), mix (:214-217), overlap-add (:376-383) -------------
  if (per_item) {
    double dc[1] = {0.0};
    for (int i = tid; i < half; i += T) dc[0] += cbuf[cpadT<C>(i)].x;        // shifted [N/2, N) = raw [0, N/2)
    block_sum<1>(dc, red);
    const double sqrt_noise = sqrt((double)c0.noise_raw);
    double* __restrict__ y = y_all + y_off[c0.utt];
    for (int jj = tid; jj < N; jj += T) {
      const int raw = jj < half ? jj + half : jj - half;                 // fftshift
      const C v = cbuf[cpadT<C>(raw)];
      const double pr = jj < half ? -dc[0] * dc_remover[jj] : (double)v.x - dc[0] * dc_remover[jj];
      const double r = (pr * sqrt_noise + (double)v.y) / N;
      const int oi = jj + c0.index - half + 1;
      if (oi >= 0 && oi <= c0.y_len - 1) ola_add(y, oi, r);
    }
  } else {
    double* __restrict__ ya = y_all + y_off[c0.utt];
    double* __restrict__ yb = y_all + y_off[c1.utt];
    for (int jj = tid; jj < N; jj += T) {
      const int raw = jj < half ? jj + half : jj - half;
      const C v = cbuf[cpadT<C>(raw)];
      const int oa = jj + c0.index - half + 1;
      if (oa >= 0 && oa <= c0.y_len - 1) ola_add(ya, oa, (double)v.x / N);
      if (pb >= 0) {
        const int ob = jj + c1.index - half + 1;
        if (ob >= 0 && ob <= c1.y_len - 1) ola_add(yb, ob, (double)v.y / N);
      }
    }
  }
}

}  // namespace

#ifndef WB_HOST_EMU      // the launcher; tests/emu has its own
bool synthesis_run(Batch* b, const int* y_len) {
  Context* ctxp = ctx();
  if (!ctxp) return false;
  cudaStream_t st = ctxp->stream;
  const int n_utt = b->n_utt;
  const int N = b->fft_size;
  int log2n = 0;
  while ((1 << log2n) < N) ++log2n;
  if ((1 << log2n) != N || log2n < 5 || log2n > 12) { set_error("Synthesis: unsupported fft_size %d", N); return false; }
  for (int u = 0; u < n_utt; ++u)
    if (b->h_f_len[u] < 2) { set_error("Synthesis: utterance %d has fewer than 2 frames", u); return false; }
  b->h_y_off.resize(n_utt);
  b->h_y_len.assign(y_len, y_len + n_utt);
  long long o = 0;
  for (int u = 0; u < n_utt; ++u) { b->h_y_off[u] = o; o += (y_len[u] + 1) & ~1LL; }
  b->total_y = o;
  if (!b->y.alloc((size_t)o) || !b->y_off.alloc(n_utt) || !b->y_len.alloc(n_utt)) return false;
  if (n_utt == 0) return true;
  if (!write_dev(b->y_off.p, b->h_y_off.data(), n_utt * sizeof(long long))) return false;
  if (!write_dev(b->y_len.p, b->h_y_len.data(), n_utt * sizeof(int))) return false;
  if (!dev_fill(b->y.p, 0, (size_t)o * sizeof(double))) return false;

  SynthConst c;
  c.fs = b->fs;
  c.log2n = log2n;
  c.frame_period_s = b->frame_period / 1000.0;
  c.lowest_f0 = b->fs / N + 1.0;                    // integer division, W/src/synthesis.cpp:359
  c.f0_max_len = b->max_f_len;

  // pulse time base: one pass into per-utterance regions sized by a cheap upper bound
  DevBuf<int> d_cnt, d_poff, d_cap;
  if (!d_cnt.alloc(n_utt) || !d_poff.alloc(n_utt) || !d_cap.alloc(n_utt)) return false;
  synth_pulse_bound_kernel<<<n_utt, 128, 0, st>>>(b->f0.p, b->f_off.p, b->f_len.p, c, d_cap.p);
  WB_LAUNCH_CHECK();
  std::vector<int> h_cap(n_utt), h_cnt(n_utt), h_poff(n_utt);
  if (!read_back(h_cap.data(), d_cap.p, n_utt * sizeof(int))) return false;
  long long total_p = 0;
  int max_y = 0;
  for (int u = 0; u < n_utt; ++u) { h_poff[u] = (int)total_p; total_p += h_cap[u]; max_y = std::max(max_y, y_len[u]); }
  if (total_p > 0x7fffffffLL) { set_error("Synthesis: too many pulses"); return false; }
  if (!ensure_randn((size_t)max_y + 16)) return false;
  DevBuf<int> p_index, p_utt;
  DevBuf<double> p_shift, d_rem;
  DevBuf<unsigned char> p_vuv;
  if (!p_index.alloc(total_p) || !p_utt.alloc(total_p) || !p_shift.alloc(total_p) || !p_vuv.alloc(total_p) || !d_rem.alloc(N)) return false;
  if (!write_dev(d_poff.p, h_poff.data(), n_utt * sizeof(int))) return false;
  if (!dev_fill(p_utt.p, 0xff, (size_t)total_p * sizeof(int))) return false;   // -1 = unused slot
  {
    const int n_chunks_max = (max_y + kTbChunk - 1) / kTbChunk + 1;
    DevBuf<double> d_inc, d_tot;
    DevBuf<unsigned char> d_vuv;
    DevBuf<int> d_counts;
    if (!d_inc.alloc((size_t)b->total_y + 1) || !d_tot.alloc((size_t)b->total_y + 1) || !d_vuv.alloc((size_t)b->total_y + 1) ||
        !d_counts.alloc((size_t)n_utt * n_chunks_max))
      return false;
    KernelTimer kt2("synth_timebase_kernel");               // all five launches of the time base
    synth_inc_kernel<<<dim3((max_y + kTbChunk - 1) / kTbChunk, n_utt), 256, 0, st>>>(b->f0.p, b->f_off.p, b->f_len.p, b->y_len.p, b->y_off.p, c,
                                                                      d_inc.p, d_vuv.p);
    WB_LAUNCH_CHECK();
    if (option("synth_phase4")) synth_phase4_kernel<<<n_utt, kTbChunk / kTbPer, 0, st>>>(d_inc.p, b->y_off.p, b->y_len.p, d_tot.p);
    else synth_phase_kernel<<<n_utt, kTbChunk, 0, st>>>(d_inc.p, b->y_off.p, b->y_len.p, d_tot.p);
    WB_LAUNCH_CHECK();
    synth_pulses_kernel<false><<<dim3(n_chunks_max, n_utt), 256, 0, st>>>(d_tot.p, d_vuv.p, b->y_off.p, b->y_len.p, c, n_chunks_max,
        d_counts.p, d_poff.p, d_cap.p, p_index.p, p_shift.p, p_vuv.p, p_utt.p);
    WB_LAUNCH_CHECK();
    synth_pulse_scan_kernel<<<(n_utt + 127) / 128, 128, 0, st>>>(d_counts.p, n_utt, n_chunks_max, d_cnt.p);
    WB_LAUNCH_CHECK();
    synth_pulses_kernel<true><<<dim3(n_chunks_max, n_utt), 256, 0, st>>>(d_tot.p, d_vuv.p, b->y_off.p, b->y_len.p, c, n_chunks_max,
        d_counts.p, d_poff.p, d_cap.p, p_index.p, p_shift.p, p_vuv.p, p_utt.p);
    WB_LAUNCH_CHECK(); kt2.stop();
    // the scratch is released in stream order when this scope ends (stream-ordered pool)
  }
  // GetDCRemover (:322-334)
  std::vector<double> rem(N);
  double dc_component = 0.0;
  for (int i = 0; i < N / 2; ++i) {
    rem[i] = 0.5 - 0.5 * cos(2.0 * kPi * (i + 1.0) / (1.0 + N));
    rem[N - i - 1] = rem[i];
    dc_component += rem[i] * 2.0;
  }
  for (int i = 0; i < N / 2; ++i) { rem[i] /= dc_component; rem[N - i - 1] = rem[i]; }
  if (!write_dev(d_rem.p, rem.data(), N * sizeof(double))) return false;
  // classify the pulses into periodic / non-periodic work lists
  DevBuf<int> d_cnt2, list_per, list_aper, d_blk;
  const int n_cblocks = (int)((total_p + 255) / 256);
  if (!d_cnt2.alloc(2) || !list_per.alloc(total_p) || !list_aper.alloc(total_p) || !d_blk.alloc(2 * (size_t)std::max(1, n_cblocks))) return false;
  if (!dev_fill(d_cnt2.p, 0, 2 * sizeof(int))) return false;
  if (n_cblocks > 0) {
    synth_classify_kernel<false><<<n_cblocks, 256, 0, st>>>(b->ap.p, b->f_off.p, b->f_len.p, p_index.p, p_vuv.p, p_utt.p, (int)total_p, c,
                                                          d_blk.p, list_per.p, list_aper.p);
    WB_LAUNCH_CHECK();
    synth_classify_scan_kernel<<<1, 64, 0, st>>>(d_blk.p, n_cblocks, d_cnt2.p);
    WB_LAUNCH_CHECK();
    synth_classify_kernel<true><<<n_cblocks, 256, 0, st>>>(b->ap.p, b->f_off.p, b->f_len.p, p_index.p, p_vuv.p, p_utt.p, (int)total_p, c,
                                                         d_blk.p, list_per.p, list_aper.p);
    WB_LAUNCH_CHECK();
  }
  int h_cnt2[2] = {0, 0};
  if (!read_back(h_cnt2, d_cnt2.p, 2 * sizeof(int))) return false;
  if (!read_back(h_cnt.data(), d_cnt.p, n_utt * sizeof(int))) return false;
  for (int u = 0; u < n_utt; ++u)
    if (h_cnt[u] > h_cap[u]) { set_error("Synthesis: utterance %d has %d pulses, bound was %d", u, h_cnt[u], h_cap[u]); return false; }
  const int n_per = h_cnt2[0], n_aper = h_cnt2[1];
  const unsigned n_items = (unsigned)(n_per + (n_aper + 1) / 2);
  if (n_items == 0) return true;
  static const bool fp64 = getenv("WB_SYNTH_FP64") != nullptr;      // debugging aid: all four transforms in FP64
  const size_t smem = 2 * cpad_size(N) * (fp64 ? sizeof(double2) : sizeof(float2)) + 96 * sizeof(double);
  if (N / 2 / 256 + 1 > 9) { set_error("Synthesis: fft_size %d too large for the register fold", N); return false; }
  KernelTimer kt3("synth_pulse_kernel");
#define WB_SP_LAUNCH(L, CT, TW)                                                                                     \
  do {                                                                                                              \
    WB_CUDA_OR_RETURN(cudaFuncSetAttribute(synth_item_kernel<L, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), false); \
    synth_item_kernel<L, CT><<<n_items, 256, smem, st>>>(b->sp.p, b->ap.p, b->f_off.p, b->f_len.p, b->y_off.p, b->y_len.p, d_poff.p, d_cnt.p, \
        p_index.p, p_shift.p, p_vuv.p, p_utt.p, list_per.p, list_aper.p, n_per, n_aper, ctxp->d_randn, TW, d_rem.p, c, b->y.p); \
  } while (0)
  if (fp64) {
    switch (log2n) {
      case 11: WB_SP_LAUNCH(11, double2, ctxp->tw_c(11)); break;
      default: WB_SP_LAUNCH(0, double2, ctxp->d_twiddle); break;
    }
  } else {
    switch (log2n) {
      case 10: WB_SP_LAUNCH(10, float2, ctxp->tw_cf(10)); break;
#ifdef WB_SYNTH_T128      // experiment (build.py --variant): 128-thread CTAs, six per SM
      case 11:
        WB_CUDA_OR_RETURN(cudaFuncSetAttribute(synth_item_kernel<11, float2, 128, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), false);
        synth_item_kernel<11, float2, 128, 6><<<n_items, 128, smem, st>>>(b->sp.p, b->ap.p, b->f_off.p, b->f_len.p, b->y_off.p, b->y_len.p, d_poff.p, d_cnt.p,
            p_index.p, p_shift.p, p_vuv.p, p_utt.p, list_per.p, list_aper.p, n_per, n_aper, ctxp->d_randn, ctxp->tw_cf(11), d_rem.p, c, b->y.p);
        break;
#else
      case 11: WB_SP_LAUNCH(11, float2, ctxp->tw_cf(11)); break;
#endif
      case 12: WB_SP_LAUNCH(12, float2, ctxp->tw_cf(12)); break;
      default: WB_SP_LAUNCH(0, float2, ctxp->d_twiddle_f); break;
    }
  }
#undef WB_SP_LAUNCH
  WB_LAUNCH_CHECK(); kt3.stop();
  synth_ola_finish_kernel<<<148 * 8, 256, 0, st>>>(b->y.p, b->total_y);      // integer sums -> samples
  WB_LAUNCH_CHECK();
  WB_CUDA_OR_RETURN(cudaStreamSynchronize(st), false);
  return true;
}

#endif  // WB_HOST_EMU

}  // namespace wb
