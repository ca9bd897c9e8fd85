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

// Time base: WRITE = false counts pulses, WRITE = true stores them.
template <bool WRITE>
__global__ void __launch_bounds__(512)
synth_timebase_kernel(const double* __restrict__ f0_all, const int* __restrict__ f_off,
                      const int* __restrict__ f_len, const int* __restrict__ y_len_all, SynthConst c,
                      int* __restrict__ pulse_count, const int* __restrict__ pulse_off,
                      int* __restrict__ p_index, double* __restrict__ p_shift,
                      unsigned char* __restrict__ p_vuv, int* __restrict__ p_utt) {
  __shared__ double inc_s[512];
  __shared__ int wcnt[32];
  __shared__ double carry_phase, last_wrap_prev;
  __shared__ int carry_cnt;
  __shared__ double wrap_s[512 + 1];
  __shared__ unsigned char vuv_s[512 + 1];
  const int u = blockIdx.x;
  const double* __restrict__ f0 = f0_all + f_off[u];
  const int n_frames = f_len[u];
  const int y_len = y_len_all[u];
  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = T >> 5;
  const int n_knots = n_frames + 1;
  const double fp = c.frame_period_s;
  if (tid == 0) { carry_phase = 0.0; carry_cnt = 0; last_wrap_prev = 0.0; }
  __syncthreads();
  const int base_out = WRITE ? pulse_off[u] : 0;
  // sample i needs wrap[i] and wrap[i+1]; process chunks of T samples, keeping the previous
  // chunk's last sample for the pair that straddles the chunk boundary.
  for (int base = 0; base < y_len; base += T) {
    const int i = base + tid;
    double inc = 0.0;
    unsigned char vuv = 0;
    if (i < y_len) {
      const double t = (double)i / (double)c.fs;                       // :227-228
      const int k = uniform_segment(t, fp, n_knots);
      const double x0 = mul_rn((double)(k - 1), fp), x1 = mul_rn((double)k, fp);
      const double s = div_rn(add_rn(t, -x0), add_rn(x1, -x0));
      const double fa = coarse_f0_at(f0, n_frames, k - 1, c.lowest_f0), fb = coarse_f0_at(f0, n_frames, k, c.lowest_f0);
      const double va = coarse_vuv_at(f0, n_frames, k - 1, c.lowest_f0), vb = coarse_vuv_at(f0, n_frames, k, c.lowest_f0);
      double fi = add_rn(fa, mul_rn(s, add_rn(fb, -fa)));
      const double vi = add_rn(va, mul_rn(s, add_rn(vb, -va)));
      vuv = vi > 0.5 ? 1 : 0;                                           // :305-309
      if (!vuv) fi = kDefaultF0;
      inc = div_rn(mul_rn(kTwoPi, fi), (double)c.fs);                   // :250,253
    }
    // total_phase[i] = total_phase[i-1] + inc[i] (:252-253) must be accumulated in exactly the
    // reference's order: with fs / 500 Hz an integer (48 kHz, 16 kHz) every unvoiced pulse sits
    // on a wrap-around that is decided by the rounding of this running sum, so a tree-shaped
    // scan moves pulses by one sample.  One thread walks the chunk sequentially (the loads
    // are independent, only the DADD chain is serial: ~T x 8 cycles per chunk, all utterances
    // in flight at once); everything else in this kernel stays parallel.
    inc_s[tid] = inc;
    __syncthreads();
    if (tid == 0) {
      double run = carry_phase;
#pragma unroll 8
      for (int q = 0; q < T; ++q) { run = add_rn(run, inc_s[q]); inc_s[q] = run; }
    }
    __syncthreads();
    const double total = inc_s[tid];
    const double wrap = fmod(total, kTwoPi);                            // :251,254
    wrap_s[tid + 1] = wrap;
    vuv_s[tid + 1] = vuv;
    if (tid == 0) { wrap_s[0] = last_wrap_prev; }
    __syncthreads();
    // pair (j, j+1) with j = base + tid - 1: both wraps are now in shared memory
    const int j = base + tid - 1;
    bool is_pulse = false;
    double y1 = 0.0, y2 = 0.0;
    if (j >= 0 && j + 1 < y_len) {
      y1 = wrap_s[tid];
      y2 = wrap_s[tid + 1];
      is_pulse = fabs(y2 - y1) > kPi;                                   // :255,259
    }
    // compaction
    const unsigned bal = __ballot_sync(0xffffffffu, is_pulse);
    if (lane == 0) wcnt[wid] = __popc(bal);
    __syncthreads();
    int before = carry_cnt;
    for (int w = 0; w < wid; ++w) before += wcnt[w];
    const int pos = before + __popc(bal & ((1u << lane) - 1u));
    if (WRITE && is_pulse) {
      const int o = base_out + pos;
      p_index[o] = j;
      const double yy1 = y1 - kTwoPi;                                   // :271-274
      const double xx = -yy1 / (y2 - yy1);
      p_shift[o] = xx / c.fs;
      p_vuv[o] = vuv_s[tid];                                           // vuv of sample j
      p_utt[o] = u;
    }
    __syncthreads();
    if (tid == T - 1) {
      carry_phase = total;     // == inc_s[T-1], the running sum after this chunk
      last_wrap_prev = wrap;
      int tot = 0;
      for (int w = 0; w < nw; ++w) tot += wcnt[w];
      carry_cnt += tot;
    }
    if (tid == 0) vuv_s[0] = vuv_s[T];   // vuv of the chunk's last sample, for the straddling pair
    __syncthreads();
  }
  if (!WRITE && tid == 0) pulse_count[u] = carry_cnt;
}

// dynamic shared memory: [ cbuf: cpad_size(N) double2 | sa: 2*(N/2+8) doubles (se | ar, later P) |
//                          nz: (N/2+8) double2 | red: 96 doubles ]
template <int LOG2N>      // 0: size given at run time (c.log2n)
__global__ void __launch_bounds__(256)
synth_pulse_kernel(const double* __restrict__ f0_all, const double* __restrict__ sp_all,
                   const double* __restrict__ ap_all, const int* __restrict__ f_off,
                   const int* __restrict__ f_len, const long long* __restrict__ y_off,
                   const int* __restrict__ y_len_all, const int* __restrict__ pulse_off,
                   const int* __restrict__ pulse_cnt, const int* __restrict__ p_index,
                   const double* __restrict__ p_shift, const unsigned char* __restrict__ p_vuv,
                   const int* __restrict__ p_utt, const uint32_t* __restrict__ randn_tab,
                   const double2* __restrict__ tw, const double* __restrict__ dc_remover,
                   SynthConst c, double* __restrict__ y_all) {
  extern __shared__ double2 smem2[];
  const int log2n = LOG2N > 0 ? LOG2N : c.log2n;
  const int N = 1 << log2n, half = N >> 1;
  constexpr int LM = LOG2N > 0 ? LOG2N - 1 : 0;
  double2* cbuf = smem2;
  double* cbufd = reinterpret_cast<double*>(cbuf);
  double* se = reinterpret_cast<double*>(cbuf + cpad_size(N));
  double* ar = se + half + 8;
  double2* Pk = reinterpret_cast<double2*>(se);
  double2* nz = reinterpret_cast<double2*>(ar + half + 8);
  double* red = reinterpret_cast<double*>(nz + half + 8);
  const int tid = threadIdx.x, T = blockDim.x;
  const int p = blockIdx.x;
  const int u = p_utt[p];
  const int first = pulse_off[u], last = first + pulse_cnt[u] - 1;
  const int index = p_index[p];
  const int noise_size_raw = p_index[min(last, p + 1)] - index;           // :370-371
  const int noise_size = min(noise_size_raw, N);                          // memory guard
  const bool vuv = p_vuv[p] != 0;
  const int n_frames = f_len[u];
  const size_t row0 = (size_t)f_off[u];
  const double current_time = (double)index / (double)c.fs;
  // ---- GetSpectralEnvelope / GetAperiodicRatio (:140-178) ------------------------------------
  const double pos_f = current_time / c.frame_period_s;
  const int fr_floor = min(n_frames - 1, static_cast<int>(floor(pos_f)));
  const int fr_ceil = min(n_frames - 1, static_cast<int>(ceil(pos_f)));
  const double interp = pos_f - fr_floor;
  const double* __restrict__ sp0 = sp_all + (row0 + fr_floor) * (half + 1);
  const double* __restrict__ sp1 = sp_all + (row0 + fr_ceil) * (half + 1);
  const double* __restrict__ ap0 = ap_all + (row0 + fr_floor) * (half + 1);
  const double* __restrict__ ap1 = ap_all + (row0 + fr_ceil) * (half + 1);
  for (int k = tid; k <= half; k += T) {
    double s, a;
    const double a0 = fmax(0.001, fmin(0.999999999999, ap0[k]));          // common.h:111-113
    if (fr_floor == fr_ceil) {
      s = fabs(sp0[k]);
      a = a0 * a0;
    } else {
      const double a1 = fmax(0.001, fmin(0.999999999999, ap1[k]));
      s = add_rn(mul_rn(1.0 - interp, fabs(sp0[k])), mul_rn(interp, fabs(sp1[k])));
      const double m = add_rn(mul_rn(1.0 - interp, a0), mul_rn(interp, a1));
      a = m * m;
    }
    se[k] = s;
    ar[k] = a;
  }
  // ---- GetNoiseSpectrum (:19-33): transform C ------------------------------------------------
  {
    const int log2m = log2n - 1;
    const uint32_t* __restrict__ rn = randn_tab + (index - p_index[first]);
    double s1[1] = {0.0};
    for (int i = tid; i < noise_size; i += T) s1[0] += randn_from_u32(rn[i]);
    block_sum<1>(s1, red);
    const double average = s1[0] / noise_size_raw;
    for (int i = tid; i < N; i += T)
      cbufd[rfft_in_slot(i, log2m)] = i < noise_size ? randn_from_u32(rn[i]) - average : 0.0;
    fft_dit<LM, false, 256>(cbuf, log2m, tw);
    for (int k = tid; k <= half; k += T) nz[k] = rfft_bin(cbuf, log2m, k, tw);
  }
  __syncthreads();
  const bool periodic = vuv && !(ar[0] > 0.999);                          // :110
  // ---- transform A: cepstra of the two log spectra -----------------------------------------------
  for (int k = tid; k <= half; k += T) {
    const double s = se[k], a = ar[k];
    se[k] = periodic ? log(s * (1.0 - a) + kMySafeGuardMinimum) / 2.0 : 0.0;   // :115-117
    ar[k] = vuv ? log(s * a) / 2.0 : log(s) / 2.0;                               // :45-51
  }
  __syncthreads();
  for (int i = tid; i < N; i += T) {
    const int k = i <= half ? i : N - i;                                         // even extension
    cbuf[cpad(brev(i, log2n))] = make_double2(se[k], ar[k]);
  }
  fft_dit<LOG2N, false, 256>(cbuf, log2n, tw);
  // fold (common.cpp:194-206) into registers, then transform B
  {
    double2 keep[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int i = tid + q * T;
      double2 z = make_double2(0.0, 0.0);
      if (i <= half) {
        z = cbuf[cpad(i)];
        if (i > 0 && i < half) { z.x *= 2.0; z.y *= 2.0; }
      }
      keep[q] = z;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int i = tid + q * T;
      if (i < N) cbuf[cpad(brev(i, log2n))] = keep[q];
    }
  }
  fft_dit<LOG2N, false, 256>(cbuf, log2n, tw);
  // ---- minimum-phase spectra, time shift, noise product ---------------------------------------
  const double coefficient = div_rn(mul_rn(mul_rn(kTwoPi, p_shift[p]), (double)c.fs), (double)N);   // :128-129
  for (int k = tid; k <= half; k += T) {
    const double2 A = cbuf[cpad(k)];
    const double2 B = cbuf[cpad((N - k) & (N - 1))];
    // S1 = (A + conj B)/2, S2 = (A - conj B)/(2i)
    const double s1r = 0.5 * (A.x + B.x), s1i = 0.5 * (A.y - B.y);
    const double s2r = 0.5 * (A.y + B.y), s2i = 0.5 * (B.x - A.x);
    double2 Pv = make_double2(0.0, 0.0);
    if (periodic) {
      const double e1 = exp(s1r / N);
      double sn, cs;
      sincos(s1i / N, &sn, &cs);
      const double mr = e1 * cs, mi = e1 * sn;
      const double re2 = cos(coefficient * k);
      const double im2 = sqrt(1.0 - re2 * re2);                          // :95 (always >= 0)
      Pv = make_double2(mr * re2 + mi * im2, mi * re2 - mr * im2);
    }
    const double e2 = exp(s2r / N);
    double sn2, cs2;
    sincos(s2i / N, &sn2, &cs2);
    const double2 M2 = make_double2(e2 * cs2, e2 * sn2);
    const double2 Z = nz[k];
    nz[k] = make_double2(M2.x * Z.x - M2.y * Z.y, M2.x * Z.y + M2.y * Z.x);   // :56-65
    Pk[k] = Pv;
  }
  __syncthreads();
  // ---- transform D: inverse of P + i A (Hermitian extensions; Im of bins 0 and N/2 ignored) ---
  for (int k = tid; k < N; k += T) {
    const int kk = k <= half ? k : N - k;
    double2 Pv = Pk[kk], Av = nz[kk];
    if (kk == 0 || kk == half) { Pv.y = 0.0; Av.y = 0.0; }
    if (k > half) { Pv.y = -Pv.y; Av.y = -Av.y; }
    cbuf[cpad(brev(k, log2n))] = make_double2(Pv.x - Av.y, Pv.y + Av.x);
  }
  fft_dit<LOG2N, true, 256>(cbuf, log2n, tw);
  // ---- fftshift, RemoveDCComponent (:73-82), mix (:214-217), overlap-add (:376-383) ----------
  double dc[1] = {0.0};
  if (periodic)
    for (int i = tid; i < half; i += T) dc[0] += cbuf[cpad(i)].x;        // shifted [N/2, N) = raw [0, N/2)
  block_sum<1>(dc, red);
  const double sqrt_noise = sqrt((double)noise_size_raw);
  const int y_len = y_len_all[u];
  double* __restrict__ y = y_all + y_off[u];
  for (int jj = tid; jj < N; jj += T) {
    const int raw = jj < half ? jj + half : jj - half;                   // fftshift
    const double2 v = cbuf[cpad(raw)];
    double pr = 0.0;
    if (periodic) pr = jj < half ? -dc[0] * dc_remover[jj] : v.x - dc[0] * dc_remover[jj];
    const double r = (pr * sqrt_noise + v.y) / N;
    const int oi = jj + index - half + 1;
    if (oi >= 0 && oi <= y_len - 1) atomicAdd(&y[oi], r);
  }
}

}  // namespace

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
  WB_CUDA_OR_RETURN(cudaMemcpyAsync(b->y_off.p, b->h_y_off.data(), n_utt * sizeof(long long), cudaMemcpyHostToDevice, st), false);
  WB_CUDA_OR_RETURN(cudaMemcpyAsync(b->y_len.p, b->h_y_len.data(), n_utt * sizeof(int), cudaMemcpyHostToDevice, st), false);
  WB_CUDA_OR_RETURN(cudaMemsetAsync(b->y.p, 0, (size_t)o * sizeof(double), st), false);

  SynthConst c;
  c.fs = b->fs;
  c.log2n = log2n;
  c.frame_period_s = b->frame_period / 1000.0;
  c.lowest_f0 = b->fs / N + 1.0;                    // integer division, W/src/synthesis.cpp:359
  c.f0_max_len = b->max_f_len;

  DevBuf<int> d_cnt, d_poff;
  if (!d_cnt.alloc(n_utt) || !d_poff.alloc(n_utt)) return false;
  KernelTimer kt1("synth_timebase_kernel");
  synth_timebase_kernel<false><<<n_utt, 512, 0, st>>>(b->f0.p, b->f_off.p, b->f_len.p, b->y_len.p, c, d_cnt.p,
                                                      nullptr, nullptr, nullptr, nullptr, nullptr);
  WB_LAUNCH_CHECK(); kt1.stop();
  std::vector<int> h_cnt(n_utt), h_poff(n_utt);
  WB_CUDA_OR_RETURN(cudaMemcpyAsync(h_cnt.data(), d_cnt.p, n_utt * sizeof(int), cudaMemcpyDeviceToHost, st), false);
  WB_CUDA_OR_RETURN(cudaStreamSynchronize(st), false);
  long long total_p = 0;
  int max_y = 0;
  for (int u = 0; u < n_utt; ++u) { h_poff[u] = (int)total_p; total_p += h_cnt[u]; max_y = std::max(max_y, y_len[u]); }
  if (total_p > 0x7fffffffLL) { set_error("Synthesis: too many pulses"); return false; }
  if (total_p == 0) return true;
  if (!ensure_randn((size_t)max_y + 16)) return false;
  DevBuf<int> p_index, p_utt;
  DevBuf<double> p_shift, d_rem;
  DevBuf<unsigned char> p_vuv;
  if (!p_index.alloc(total_p) || !p_utt.alloc(total_p) || !p_shift.alloc(total_p) || !p_vuv.alloc(total_p) || !d_rem.alloc(N)) return false;
  WB_CUDA_OR_RETURN(cudaMemcpyAsync(d_poff.p, h_poff.data(), n_utt * sizeof(int), cudaMemcpyHostToDevice, st), false);
  KernelTimer kt2("synth_timebase_kernel");
  synth_timebase_kernel<true><<<n_utt, 512, 0, st>>>(b->f0.p, b->f_off.p, b->f_len.p, b->y_len.p, c, nullptr, d_poff.p,
                                                     p_index.p, p_shift.p, p_vuv.p, p_utt.p);
  WB_LAUNCH_CHECK(); kt2.stop();
  // GetDCRemover (:322-334)
  std::vector<double> rem(N);
  double dc_component = 0.0;
  for (int i = 0; i < N / 2; ++i) {
    rem[i] = 0.5 - 0.5 * cos(2.0 * kPi * (i + 1.0) / (1.0 + N));
    rem[N - i - 1] = rem[i];
    dc_component += rem[i] * 2.0;
  }
  for (int i = 0; i < N / 2; ++i) { rem[i] /= dc_component; rem[N - i - 1] = rem[i]; }
  WB_CUDA_OR_RETURN(cudaMemcpyAsync(d_rem.p, rem.data(), N * sizeof(double), cudaMemcpyHostToDevice, st), false);
  const size_t smem = cpad_size(N) * sizeof(double2) + (size_t)(2 * (N / 2 + 8)) * sizeof(double) +
                      (size_t)(N / 2 + 8) * sizeof(double2) + 96 * sizeof(double);
  if (N / 256 > 8) { set_error("Synthesis: fft_size %d too large for the register fold", N); return false; }
  KernelTimer kt3("synth_pulse_kernel");
#define WB_SP_LAUNCH(L)                                                                                             \
  do {                                                                                                              \
    WB_CUDA_OR_RETURN(cudaFuncSetAttribute(synth_pulse_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), false); \
    synth_pulse_kernel<L><<<(unsigned)total_p, 256, smem, st>>>(b->f0.p, b->sp.p, b->ap.p, b->f_off.p, b->f_len.p, b->y_off.p, b->y_len.p, d_poff.p, d_cnt.p, p_index.p, p_shift.p, p_vuv.p, p_utt.p, ctxp->d_randn, ctxp->d_twiddle, d_rem.p, c, b->y.p); \
  } while (0)
  switch (log2n) {
    case 10: WB_SP_LAUNCH(10); break;
    case 11: WB_SP_LAUNCH(11); break;
    case 12: WB_SP_LAUNCH(12); break;
    default: WB_SP_LAUNCH(0); break;
  }
#undef WB_SP_LAUNCH
  WB_LAUNCH_CHECK(); kt3.stop();
  WB_CUDA_OR_RETURN(cudaStreamSynchronize(st), false);
  return true;
}

}  // namespace wb
