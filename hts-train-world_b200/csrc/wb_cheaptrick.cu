// world-b200: CheapTrick spectral envelope, one CTA per frame.
//
// Reference: W/src/cheaptrick.cpp (CheapTrick :200-228, CheapTrickGeneralBody :159-187,
// GetWindowedWaveform :112-142, GetPowerSpectrum :64-82, SmoothingWithRecovery :22-57,
// AddInfinitesimalNoise :147-151) and W/src/common.cpp (DCCorrection :56-75,
// LinearSmoothing :77-111).  The reference walks the frames of one utterance sequentially on
// one core; here every frame of every utterance in the batch is an independent CTA whose
// whole working set (windowed frame, power spectrum, mirrored cumulative spectrum, cepstrum)
// stays in shared memory; HBM sees one read of the window and one write of the output row.
//
// randn: the reference draws (2*hwl+1) + (N/2+1) variates per frame, in frame order
// (SURVEY Appendix A1).  A per-utterance exclusive scan of those counts gives each frame its
// offset into the precomputed randn table, so the dither is bit-identical to the reference.
#include <stdlib.h>
#include "wb_batch.h"
#include "wb_fft.cuh"
#include "wb_spectral.cuh"

namespace wb {

__device__ __forceinline__ double cheaptrick_f0(double f0, double f0_floor) {
  return f0 <= f0_floor ? kDefaultF0 : f0;       // W/src/cheaptrick.cpp:217
}
__device__ __forceinline__ int cheaptrick_hwl(int fs, double f0c) {
  return matlab_round(div_rn(mul_rn(1.5, (double)fs), f0c));   // :114
}

__global__ void cheaptrick_count_kernel(const double* __restrict__ f0, int total_frames, int fs,
                                        int fft_size, double f0_floor,
                                        long long* __restrict__ counts) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= total_frames) return;
  const double f0c = cheaptrick_f0(f0[f], f0_floor);
  counts[f] = 2LL * cheaptrick_hwl(fs, f0c) + 1 + fft_size / 2 + 1;
}

// dynamic shared memory: [ buf: cpad_size(N/2) double2 | aux: N + 16 doubles | red: 96 doubles ]
#ifndef WB_CT_MAXK32
#define WB_CT_MAXK32 4
#endif
constexpr int kCtMaxK32 = WB_CT_MAXK32;      // radix 2^k of the two FP32 liftering transforms (experiments: build.py --variant)
template <int LOG2N, int THREADS>      // LOG2N 0: size given at run time (log2n_rt)
__global__ void __launch_bounds__(THREADS, 768 / THREADS)
cheaptrick_kernel(UttView u, const int* __restrict__ frame_utt, const double* __restrict__ frame_t,
                  const double* __restrict__ f0_in, const long long* __restrict__ rng_off,
                  const uint32_t* __restrict__ randn_tab, const double2* __restrict__ tw,
                  const float2* __restrict__ twf, int fs, int log2n_rt, double q1, double f0_floor, double* __restrict__ sp_out) {
  WB_DYN_SMEM(double2, smem2);
  const int log2n = LOG2N > 0 ? LOG2N : log2n_rt;
  constexpr int LM = LOG2N > 0 ? LOG2N - 1 : 0;
  constexpr int TWL = LOG2N > 0 ? LOG2N : kTwLog2;       // compact twiddle tables of this size, or the master tables
  const int N = 1 << log2n, M = N >> 1, log2m = log2n - 1, half = M;   // half = N/2
  double2* buf = smem2;
  double* bufd = reinterpret_cast<double*>(buf);
  double* aux = reinterpret_cast<double*>(buf + cpad_size(M));
  double* red = aux + N + 16;
  const int tid = threadIdx.x, T = blockDim.x;
  const int f = blockIdx.x;
  const int utt = frame_utt[f];
  const double* __restrict__ x = u.x + u.x_off[utt];
  const int x_len = u.x_len[utt];
  const double t_pos = frame_t[f];
  const double f0c = cheaptrick_f0(f0_in[f], f0_floor);
  const uint32_t* __restrict__ rn = randn_tab + rng_off[f];
  double* __restrict__ out = sp_out + (size_t)f * (half + 1);

  const int hwl = cheaptrick_hwl(fs, f0c);
  const int W = 2 * hwl + 1;
  const int boundary = static_cast<int>(mul_rn(f0c * 2.0 / 3.0, (double)N) / fs) + 1;
  if (W > N || half + 2 * boundary + 1 > N + 16 || !(f0c > 0.0)) {
    // outside the domain the reference supports (it would index out of bounds)
    for (int k = tid; k <= half; k += T) out[k] = __longlong_as_double(0x7ff8000000000000LL);
    return;
  }

  // ---- GetWindowedWaveform (:112-142) ---------------------------------------------------
  // w_i = 0.5 cos(theta_i) + 0.5 with theta_i = (i - hwl) pi f0 / (1.5 fs): one sincos per thread,
  // then angle-addition steps of T samples.  The reference normalises w by sqrt(sum w^2), adds
  // the dither and removes the weighted mean; all four sums come from ONE block reduction:
  //   wave_i = x_i w_i / norm + d_i,   coef = sum(wave) / sum(w / norm) = (Sx + Sd norm) / Sw.
  const int origin = matlab_round(add_rn(mul_rn(t_pos, (double)fs), 0.001));
  {
    // stage the sample window with one TMA bulk copy into aux (where the window coefficients will
    // overwrite it in place); frames whose window crosses an utterance edge gather with clamping
    __shared__ uint64_t mbar;
    int a0 = 0, n_stage = 0;
    const bool staged = bulk_window_range(origin - hwl, W, x_len, N + 16, &a0, &n_stage);
    double* waux = aux + (staged ? (origin - hwl) - a0 : 0);     // sample / coefficient i lives at waux[i]
    if (staged) {
      if (tid == 0) mbar_init(&mbar, 1);
      __syncthreads();
      if (tid == 0) bulk_load_issue(aux, x + a0, (unsigned)n_stage * 8u, &mbar);
    }
    const double turn_step = f0c / (1.5 * fs);                   // angle step in units of pi
    double cs, sn, cs_step, sn_step;
    sincospi((double)(tid - hwl) * turn_step, &sn, &cs);
    sincospi((double)T * turn_step, &sn_step, &cs_step);
    double sums[4] = {0.0, 0.0, 0.0, 0.0};            // Sww, Sw, Sx, Sd
    if (staged) mbar_wait(&mbar, 0);
    for (int i = tid; i < W; i += T) {
      const double w = 0.5 * cs + 0.5;
      {
        const double c2 = cs * cs_step - sn * sn_step;
        sn = sn * cs_step + cs * sn_step;
        cs = c2;
      }
      const double xv = staged ? waux[i] : x[min(x_len - 1, max(0, origin + i - hwl))];
      const double xw = xv * w;
      waux[i] = w;
      bufd[rfft_in_slot(i, log2m)] = xw;
      sums[0] += w * w; sums[1] += w; sums[2] += xw;
      sums[3] += randn_from_u32(rn[i]) * kMySafeGuardMinimum;
    }
    block_sum<4>(sums, red);
    const double inv_norm = 1.0 / sqrt(sums[0]);
    const double coef = (sums[2] + sums[3] / inv_norm) / sums[1];
    for (int i = tid; i < N; i += T) {
      const int slot = rfft_in_slot(i, log2m);
      double wave = 0.0;
      if (i < W)
        wave = bufd[slot] * inv_norm + randn_from_u32(rn[i]) * kMySafeGuardMinimum - waux[i] * inv_norm * coef;
      bufd[slot] = wave;
    }
  }
  // ---- GetPowerSpectrum (:64-82) -----------------------------------------------------------
  fft_dit<LM, false, THREADS, THREADS == 256 ? 4 : 3, TWL>(buf, log2m, tw);
  for (int k = tid; k <= half; k += T) {
    const double2 X = rfft_bin<TWL>(buf, log2m, k, tw);
    aux[k] = X.x * X.x + X.y * X.y;
  }
  __syncthreads();
  // DCCorrection (common.cpp:56-75)
  const double inv_df = (double)N / fs;
  const double inv_n = 1.0 / N;                 // N is a power of two: x * inv_n == x / N exactly
  dc_correction<true>(aux, bufd, f0c, fs, N);
  // ---- LinearSmoothing (common.cpp:77-111), width = f0 * 2 / 3 ------------------------------
  const double width = f0c * 2.0 / 3.0;
  const int len = half + 2 * boundary + 1;
  if (!mirrored_cumsum<9>(aux, bufd, red, half, boundary, fs, inv_n)) {
    for (int i = tid; i < len; i += T) {
      double v;
      if (i < boundary) v = aux[boundary - i];
      else if (i < half + boundary) v = aux[i - boundary];
      else v = aux[half - (i - (half + boundary))];
      bufd[i] = mul_rn(v, (double)fs) * inv_n;
    }
    __syncthreads();
    block_inclusive_scan(bufd, len, red);
  }
  {
    // (cum(k + c1) - cum(k + c0)) / width: the offsets of the two interpolation points are the same for every
    // bin (wb_spectral.cuh, linear_smoothing), so their integer parts and fractions are computed once
    const double c0 = (boundary - 0.5) - 0.5 * width * inv_df;
    const double c1 = c0 + width * inv_df;
    const int i0 = static_cast<int>(c0), i1 = static_cast<int>(c1);
    const double fr0 = c0 - i0, fr1 = c1 - i1;
    const double inv_width = 1.0 / width;
    const uint32_t* __restrict__ rn2 = rn + W;
    for (int k = tid; k <= half; k += T) {
      const int b0 = min(len - 1, k + i0), b1 = min(len - 1, k + i1);
      const double l0 = bufd[b0], h0 = bufd[b1];
      const double low = fma(b0 + 1 < len ? bufd[b0 + 1] - l0 : 0.0, fr0, l0);
      const double high = fma(b1 + 1 < len ? bufd[b1 + 1] - h0 : 0.0, fr1, h0);
      const double sm = (high - low) * inv_width;
      // AddInfinitesimalNoise (:147-151) then log (:38-39)
      // the logarithm feeds the FP32 liftering transforms: FP32 accuracy is all that survives
      aux[k] = logf(static_cast<float>(sm + fabs(randn_from_u32(rn2[k])) * kEps));
    }
  }
  __syncthreads();
  // ---- SmoothingWithRecovery (:22-57) ---------------------------------------------------------
  // Both transforms act on the LOG spectrum (|values| <= ~40) and its cepstrum; in FP32 their
  // error is ~1e-5 nepers, i.e. 1e-4 dB against the 0.01 dB tolerance, so they run in FP32
  // (half the shared-memory traffic, twice the FMA rate).  exp() is taken in FP64.
  {
    float2* fb = reinterpret_cast<float2*>(buf);
    float* fbs = reinterpret_cast<float*>(buf);
    for (int i = tid; i < N; i += T) fbs[rfft_in_slot_f(i, log2m)] = static_cast<float>(aux[i <= half ? i : N - i]);
    fft_dit<LM, false, THREADS, kCtMaxK32, TWL>(fb, log2m, twf);
    float* lif = reinterpret_cast<float*>(aux);         // liftered cepstrum, real
    __syncthreads();                                     // everyone has read aux
    for (int k = tid; k <= half; k += T) {
      const float re = rfft_bin<TWL>(fb, log2m, k, twf).x;
      float lifter = 1.f, comp = 1.f;
      if (k > 0) {
        const float fq = static_cast<float>(f0c * ((double)k / fs));    // f0 * quefrency
        float sn, cs;
        sincospif(fq, &sn, &cs);                         // sin(pi f0 q), cos(pi f0 q)
        lifter = sn / (static_cast<float>(kPi) * fq);
        comp = static_cast<float>(1.0 - 2.0 * q1) + static_cast<float>(2.0 * q1) * (2.f * cs * cs - 1.f);   // cos(2 pi f0 q)
      }
      lif[k] = re * lifter * comp * (1.f / N);
    }
    __syncthreads();
    for (int k = tid; k < half; k += T) {
      const float2 z = c2r_pack<TWL>(make_float2(lif[k], 0.f), make_float2(lif[half - k], 0.f), k, log2m, twf);
      fb[cpadf(brev(k, log2m))] = z;
    }
    fft_dit<LM, true, THREADS, kCtMaxK32, TWL>(fb, log2m, twf);
    // the liftered log spectrum is a float: the single-precision exponential (1 ulp) keeps everything it
    // holds at a quarter of the cost; below e^-80 (digital silence: the dither's spectrum) floats would
    // go subnormal, there the double exponential is used
    for (int k = tid; k <= half; k += T) {
      const float v = fbs[rfft_out_slot_f(k)];
      out[k] = v > -80.f ? static_cast<double>(expf(v)) : exp(static_cast<double>(v));
    }
  }
}

#ifndef WB_HOST_EMU      // the launcher; tests/emu has its own
bool cheaptrick_run(const UttView& u, int fs, int total_frames, const int* frame_utt,
                    const double* frame_t, const double* f0, int fft_size, double q1,
                    double* sp) {
  Context* c = ctx();
  if (!c) return false;
  if (total_frames <= 0) return true;
  int log2n = 0;
  while ((1 << log2n) < fft_size) ++log2n;
  if ((1 << log2n) != fft_size || log2n < 5 || log2n > 14) {
    set_error("CheapTrick: unsupported fft_size %d", fft_size);
    return false;
  }
  const double f0_floor = 3.0 * fs / (fft_size - 3.0);      // GetF0FloorForCheapTrick :196-198
  if (u.max_f_len > 0) flush_deferred_copies();             // no read-back below: deferred bulk copies may start (see stonemask_run)
  DevBuf<long long> counts, offs, totals;
  if (!counts.alloc(total_frames) || !offs.alloc(total_frames) || !totals.alloc(u.n_utt)) return false;
  cudaStream_t st = c->stream;
  cheaptrick_count_kernel<<<(total_frames + 255) / 256, 256, 0, st>>>(f0, total_frames, fs, fft_size, f0_floor, counts.p);
  WB_LAUNCH_CHECK();
  if (!segmented_exclusive_scan(counts.p, u.f_off, u.f_len, u.n_utt, offs.p, totals.p)) return false;
  // The randn table must reach the largest per-utterance total.  With the frame counts known on the host an upper
  // bound does (every frame draws 2 hwl + 1 + fft_size / 2 + 1 variates, hwl <= round(1.5 fs / f0_floor)): no
  // read-back, the stage enqueues without a host round trip.  Unknown frame counts: read the totals.
  long long mx = 0;
  if (u.max_f_len > 0) {
    mx = (long long)u.max_f_len * (2LL * (static_cast<long long>(1.5 * fs / f0_floor + 0.5) + 1) + 1 + fft_size / 2 + 1);
  } else {
    std::vector<long long> h_tot(u.n_utt);
    if (!read_back(h_tot.data(), totals.p, u.n_utt * sizeof(long long))) return false;
    for (long long v : h_tot) mx = v > mx ? v : mx;
  }
  if (!ensure_randn((size_t)mx)) return false;
  const size_t smem = cpad_size(fft_size / 2) * sizeof(double2) + (fft_size + 16 + 128) * sizeof(double);
  KernelTimer kt1("cheaptrick_kernel");
  // 128-thread CTAs: 6 frames per SM instead of 3, half as many warps behind every barrier (9 % faster)
  const bool t128 = true;
#define WB_CT_LAUNCH(L)                                                                                             \
  do {                                                                                                              \
    if (t128) {                                                                                                     \
    WB_CUDA_OR_RETURN(cudaFuncSetAttribute(cheaptrick_kernel<L, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), false); \
    cheaptrick_kernel<L, 128><<<total_frames, 128, smem, st>>>(u, frame_utt, frame_t, f0, offs.p, c->d_randn, L > 0 ? c->tw_c(L > 0 ? L : 4) : c->d_twiddle, L > 0 ? c->tw_cf(L > 0 ? L : 4) : c->d_twiddle_f, fs, log2n, q1, f0_floor, sp); \
    break; }                                                                                                        \
    WB_CUDA_OR_RETURN(cudaFuncSetAttribute(cheaptrick_kernel<L, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), false); \
    cheaptrick_kernel<L, 256><<<total_frames, 256, smem, st>>>(u, frame_utt, frame_t, f0, offs.p, c->d_randn, L > 0 ? c->tw_c(L > 0 ? L : 4) : c->d_twiddle, L > 0 ? c->tw_cf(L > 0 ? L : 4) : c->d_twiddle_f, fs, log2n, q1, f0_floor, sp); \
  } while (0)
  switch (log2n) {
    case 10: WB_CT_LAUNCH(10); break;
    case 11: WB_CT_LAUNCH(11); break;
    case 12: WB_CT_LAUNCH(12); break;
    default: WB_CT_LAUNCH(0); break;
  }
#undef WB_CT_LAUNCH
  WB_LAUNCH_CHECK(); kt1.stop();
  // counts/offs are freed when this returns: make sure the kernel is done with them
  WB_CUDA_OR_RETURN(cudaStreamSynchronize(st), false);
  return true;
}

#endif  // WB_HOST_EMU

}  // namespace wb
