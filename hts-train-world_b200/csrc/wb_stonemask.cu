// world-b200: StoneMask F0 refinement, one CTA per frame.
//
// Reference: W/src/stonemask.cpp — StoneMask :211-217, GetRefinedF0 :184-207, GetMeanF0 :136-178,
// GetBaseIndex :24-28, GetMainWindow :33-43, GetDiffWindow :49-55, GetSpectra :61-91,
// FixF0 :96-117, GetTentativeF0 :122-131.
//
// The reference plans and runs two real FFTs per voiced frame (x.w and x.w').  Here both are
// one packed complex FFT z = x.w + i x.w'; with Z[k] = a+ib and Z[N-k] = c+id the two
// quantities StoneMask needs are
//     power[k]     = |X_w[k]|^2                     = ((a+c)^2 + (b-d)^2) / 4
//     numerator[k] = Re X_w Im X_w' - Im X_w Re X_w' = (|Z[N-k]|^2 - |Z[k]|^2) / 4
// and only the <= 8 harmonic bins that FixF0 reads are ever evaluated.
#include "wb_batch.h"
#include "wb_fft.cuh"

namespace wb {
namespace {

__device__ __forceinline__ bool stonemask_in_range(double f0, int fs) {
  return !(f0 <= kFloorF0StoneMask || f0 > fs / 12.0);          // :186-187
}
__device__ __forceinline__ int stonemask_hwl(double f0, int fs) {
  return static_cast<int>(add_rn(div_rn(mul_rn(1.5, (double)fs), f0), 1.0));   // :188
}
__device__ __forceinline__ int stonemask_log2fft(int hwl) {
  return 2 + (31 - __clz(2 * hwl + 1));                          // :192-193 (2*hwl+1 is odd)
}

__global__ void stonemask_maxfft_kernel(const double* __restrict__ f0, int n, int fs, int* __restrict__ out) {
  int m = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double v = f0[i];
    if (stonemask_in_range(v, fs)) m = max(m, stonemask_log2fft(stonemask_hwl(v, fs)));
  }
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

struct Bins { double power, numer; };

// The packed transform runs in FP32: the harmonic bins FixF0 reads are the strongest of the
// frame, their 2^-24 relative error moves the amplitude-weighted instantaneous frequency by
// < 1e-6 relative (tolerance 1e-4).  Windows and all later arithmetic stay in FP64.
__device__ __forceinline__ Bins stonemask_bin(const float2* cbuf, int nfft, int k) {
  k = max(0, min(nfft / 2, k));                    // memory guard (UB in the reference beyond N/2)
  const float2 Af = cbuf[cpadf(k)];
  const float2 Bf = cbuf[cpadf((nfft - k) & (nfft - 1))];
  const double2 A = make_double2(Af.x, Af.y), B = make_double2(Bf.x, Bf.y);
  Bins r;
  const double re = 0.5 * (A.x + B.x), im = 0.5 * (A.y - B.y);
  const double dre = 0.5 * (A.y + B.y), dim = 0.5 * (B.x - A.x);
  r.power = re * re + im * im;
  r.numer = re * dim - im * dre;
  return r;
}

// FixF0 (:96-117), evaluated by one thread
__device__ double stonemask_fix_f0(const float2* cbuf, int nfft, int fs, double initial_f0, int nh) {
  double numerator = 0.0, denominator = 0.0;
  for (int i = 0; i < nh; ++i) {
    const int index = matlab_round(mul_rn(div_rn(mul_rn(initial_f0, (double)nfft), (double)fs), (double)(i + 1)));
    const Bins b = stonemask_bin(cbuf, nfft, index);
    const double inst = b.power == 0.0 ? 0.0
        : add_rn(div_rn(mul_rn((double)index, (double)fs), (double)nfft),
                 div_rn(div_rn(mul_rn(div_rn(b.numer, b.power), (double)fs), 2.0), kPi));
    const double amp = sqrt(b.power);
    numerator += amp * inst;
    denominator += amp * (i + 1);
  }
  return numerator / (denominator + kMySafeGuardMinimum);
}

// dynamic shared memory: [ cbuf: cpad_size(max fft) float2 | win: max_fft/2 + 8 doubles | idx: max_fft/2 ints ]
__global__ void __launch_bounds__(256)
stonemask_kernel(UttView u, const int* __restrict__ frame_utt, const double* __restrict__ frame_t,
                 const double* __restrict__ f0_in, const float2* __restrict__ tw, int fs,
                 int max_log2fft, double* __restrict__ f0_out) {
  WB_DYN_SMEM(double2, smem2);
  const int f = blockIdx.x;
  const double f0 = f0_in[f];
  if (!stonemask_in_range(f0, fs)) { if (threadIdx.x == 0) f0_out[f] = 0.0; return; }
  float2* cbuf = reinterpret_cast<float2*>(smem2);
  double* win = reinterpret_cast<double*>(cbuf + ((cpad_size(1 << max_log2fft) + 1) & ~1));
  const int tid = threadIdx.x, T = blockDim.x;
  const int utt = frame_utt[f];
  const double* __restrict__ x = u.x + u.x_off[utt];
  const int x_len = u.x_len[utt];
  const double t_pos = frame_t[f];
  const int hwl = stonemask_hwl(f0, fs);
  const int W = 2 * hwl + 1;
  const int log2fft = stonemask_log2fft(hwl);
  const int nfft = 1 << log2fft;
  const double wlen = div_rn(add_rn(mul_rn(2.0, (double)hwl), 1.0), (double)fs);   // :189
  // GetBaseIndex + GetMainWindow.  index_raw decides which sample is read, so it is formed with the
  // reference's own roundings (one exact division per sample) and kept for the second pass; the
  // argument of the cosine only shapes the window, there the two divisions become multiplications
  // by reciprocals (the window moves by ~1e-14 relative).
  int* idx_s = reinterpret_cast<int*>(win + (1 << max_log2fft) / 2 + 8);
  const double inv_fs = 1.0 / fs, two_pi_over_wlen = 2.0 * kPi / wlen;
  for (int i = tid; i < W; i += T) {
    const double base_time = div_rn((double)(i - hwl), (double)fs);
    const int index_raw = matlab_round(mul_rn(add_rn(t_pos, base_time), (double)fs));
    idx_s[i] = max(0, min(x_len - 1, index_raw - 1));
    const double tmp = (index_raw - 1.0) * inv_fs - t_pos;
    const double cs = cos(tmp * two_pi_over_wlen);
    win[i] = 0.42 + 0.5 * cs + 0.08 * (2.0 * cs * cs - 1.0);        // cos(2a) = 2 cos^2(a) - 1
  }
  __syncthreads();
  // GetDiffWindow + GetSpectra, packed
  for (int i = tid; i < nfft; i += T) {
    float2 z = make_float2(0.f, 0.f);
    if (i < W) {
      const double xv = x[idx_s[i]];
      double dw;
      if (i == 0) dw = -win[1] / 2.0;
      else if (i == W - 1) dw = win[W - 2] / 2.0;
      else dw = -(win[i + 1] - win[i - 1]) / 2.0;
      z = make_float2(static_cast<float>(xv * win[i]), static_cast<float>(xv * dw));
    }
    cbuf[cpadf(brev(i, log2fft))] = z;
  }
  fft_dit_rt_compact<false, 256, 4>(cbuf, log2fft, tw);      // tw: base of the compact FP32 tables
  if (tid == 0) {
    // GetTentativeF0 (:122-131) and the 20 % sanity check of GetRefinedF0 (:203-204)
    double mean_f0 = 0.0;
    const double tentative = stonemask_fix_f0(cbuf, nfft, fs, f0, 2);
    if (!(tentative <= 0.0 || tentative > f0 * 2)) mean_f0 = stonemask_fix_f0(cbuf, nfft, fs, tentative, 6);
    if (fabs(mean_f0 - f0) / f0 > 0.2) mean_f0 = f0;
    f0_out[f] = mean_f0;
  }
}


// ---- direct evaluation of the harmonic bins (default path) ------------------------------------------
// FixF0 reads at most 2 + 6 bins of the two spectra, so no transform is needed at all: one WARP owns a
// frame and evaluates  X_w[k] = sum_n x_n w_n e^{-2 pi i k n / N}  and the same sum with the
// differentiated window for the bins FixF0 asks for -- lane l takes samples l, l + 32, ...; the
// phasor of every bin advances by a fixed rotation (32 samples) per iteration, its start and step come
// from the compact twiddle table of size N (exact to 1 ulp); everything is FP64.  ~80 FP64 operations
// per sample for the two passes (2 bins, then 6 bins around the tentative F0) against ~170 for the
// packed FFT plus one libm cos per sample, and no shared memory or block barrier.
//
// Window phase (GetMainWindow :33-43): a_n = 2 pi ((index_raw[n] - 1) / fs - t) / wlen with
// index_raw[n] = round((t + (n - hwl) / fs) fs) (:24-28).  When t fs is not within 1e-6 of a
// half-integer every index_raw[n] is r0 + n (the rounding cannot flip), a_n is linear in n and one
// sincospi per lane plus angle-addition steps of 32 samples give the whole window (|error| < 1e-14);
// the +-1 neighbours of the differentiated window (:49-55) are one more angle addition.  Otherwise
// (44.1 / 22.05 kHz: t fs is a half-integer on every other frame) each sample's own index_raw and
// those of its two neighbours are formed with the reference's roundings and the phase is taken
// from them directly.
template <int NH, bool linear>
__device__ __forceinline__ double stonemask_fix_f0_dft(const double* __restrict__ x, int x_len, int fs, double t_pos,
                                                       int hwl, int log2fft, double wlen, int r0,
                                                       double initial_f0, int lane,
                                                       const double2* __restrict__ tw_c_base) {
  const int W = 2 * hwl + 1, nfft = 1 << log2fft, nhalf = nfft >> 1;
  int bins[NH];
  double pc[NH], ps[NH], qc[NH], qs[NH], acc[NH][4];
  {
    const double2* __restrict__ tw = tw_c_base + Context::tw_c_offset(log2fft);
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      bins[h] = matlab_round(mul_rn(div_rn(mul_rn(initial_f0, (double)nfft), (double)fs), (double)(h + 1)));   // :103
      const int k = max(0, min(nhalf, bins[h]));                 // memory guard (UB in the reference beyond N/2)
      const int m0 = (int)(((long long)k * lane) & (nfft - 1));
      const int m1 = (int)(((long long)k * 32) & (nfft - 1));
      double2 a = __ldg(&tw[m0 & (nhalf - 1)]), b = __ldg(&tw[m1 & (nhalf - 1)]);
      if (m0 & nhalf) { a.x = -a.x; a.y = -a.y; }
      if (m1 & nhalf) { b.x = -b.x; b.y = -b.y; }
      pc[h] = a.x; ps[h] = a.y; qc[h] = b.x; qs[h] = b.y;
      acc[h][0] = acc[h][1] = acc[h][2] = acc[h][3] = 0.0;
    }
  }
  auto blackman = [](double cv) { return 0.42 + 0.5 * cv + 0.08 * (2.0 * cv * cv - 1.0); };
  const double turn_per_sample = 2.0 / (wlen * fs);               // phase step in units of pi
  auto index_raw = [&](int n) {                                     // :24-28
    return matlab_round(mul_rn(add_rn(t_pos, div_rn((double)(n - hwl), (double)fs)), (double)fs));
  };
  auto window_of_index = [&](int ir) {                              // :38-42 for one sample (slow path)
    const double tmp = add_rn(div_rn(ir - 1.0, (double)fs), -t_pos);
    return blackman(cospi(2.0 * tmp / wlen));
  };
  double cd = 1.0, sd = 0.0, c32 = 1.0, s32 = 0.0, cs = 1.0, sn = 0.0;
  if (linear) {
    sincospi(turn_per_sample, &sd, &cd);
    sincospi(32.0 * turn_per_sample, &s32, &c32);
    sincospi(2.0 * add_rn(div_rn(r0 + lane - 1.0, (double)fs), -t_pos) / wlen, &sn, &cs);
  }
  for (int n = lane; n < W; n += 32) {
    double w, w_next, w_prev;
    int ir;
    if (linear) {
      ir = r0 + n;
      w = blackman(cs);
      w_next = blackman(cs * cd - sn * sd);
      w_prev = blackman(cs * cd + sn * sd);
      const double t = cs * c32 - sn * s32;
      sn = sn * c32 + cs * s32;
      cs = t;
    } else {
      ir = index_raw(n);
      w = window_of_index(ir);
      w_next = n + 1 < W ? window_of_index(index_raw(n + 1)) : 0.0;
      w_prev = n > 0 ? window_of_index(index_raw(n - 1)) : 0.0;
    }
    double dw;                                                      // GetDiffWindow (:49-55)
    if (n == 0) dw = -w_next / 2.0;
    else if (n == W - 1) dw = w_prev / 2.0;
    else dw = -(w_next - w_prev) / 2.0;
    const double xv = x[max(0, min(x_len - 1, ir - 1))];            // GetSpectra (:67-70)
    const double xm = xv * w, xd = xv * dw;
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      acc[h][0] += xm * pc[h]; acc[h][1] += xm * ps[h];
      acc[h][2] += xd * pc[h]; acc[h][3] += xd * ps[h];
      const double t = pc[h] * qc[h] - ps[h] * qs[h];
      ps[h] = ps[h] * qc[h] + pc[h] * qs[h];
      pc[h] = t;
    }
  }
  double numerator = 0.0, denominator = 0.0;                         // FixF0 (:96-117)
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    const double re = warp_sum(acc[h][0]), im = warp_sum(acc[h][1]);
    const double dre = warp_sum(acc[h][2]), dim = warp_sum(acc[h][3]);
    const double power = re * re + im * im;
    const double numer = re * dim - im * dre;
    const double inst = power == 0.0 ? 0.0
        : add_rn(div_rn(mul_rn((double)bins[h], (double)fs), (double)nfft),
                 div_rn(div_rn(mul_rn(div_rn(numer, power), (double)fs), 2.0), kPi));
    const double amp = sqrt(power);
    numerator += amp * inst;
    denominator += amp * (h + 1);
  }
  return numerator / (denominator + kMySafeGuardMinimum);
}

// frames whose t fs sits on a half-integer (rare except at 44.1 / 22.05 kHz): out of line, so that its
// registers do not count against the common path
__device__ __noinline__ double stonemask_exact_indices(const double* __restrict__ x, int x_len, int fs, double t_pos,
                                                       int hwl, int log2fft, double wlen, double f0, int lane,
                                                       const double2* __restrict__ tw_c_base) {
  const double tentative = stonemask_fix_f0_dft<2, false>(x, x_len, fs, t_pos, hwl, log2fft, wlen, 0, f0, lane, tw_c_base);
  if (tentative <= 0.0 || tentative > f0 * 2) return 0.0;
  return stonemask_fix_f0_dft<6, false>(x, x_len, fs, t_pos, hwl, log2fft, wlen, 0, tentative, lane, tw_c_base);
}

constexpr int kSmWarps = 8;      // frames per CTA
__global__ void __launch_bounds__(kSmWarps * 32, 2)
stonemask_dft_kernel(UttView u, const int* __restrict__ frame_utt, const double* __restrict__ frame_t,
                     const double* __restrict__ f0_in, const double2* __restrict__ tw_c_base, int fs,
                     int total_frames, double* __restrict__ f0_out) {
  const int f = blockIdx.x * kSmWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (f >= total_frames) return;
  const double f0 = f0_in[f];
  if (!stonemask_in_range(f0, fs)) { if (lane == 0) f0_out[f] = 0.0; return; }
  const int utt = frame_utt[f];
  const double* __restrict__ x = u.x + u.x_off[utt];
  const int x_len = u.x_len[utt];
  const double t_pos = frame_t[f];
  const int hwl = stonemask_hwl(f0, fs);
  const int log2fft = stonemask_log2fft(hwl);
  const double wlen = div_rn(add_rn(mul_rn(2.0, (double)hwl), 1.0), (double)fs);   // :189
  // is every index_raw[n] = round(t fs) - hwl + n?  (t + (n - hwl) / fs) fs differs from t fs + (n - hwl)
  // by a few ulps of t fs (< 1e-9 for any utterance length): the rounding can only flip within that
  // distance of a half-integer
  const double p = mul_rn(t_pos, (double)fs);
  const double fr = p - floor(p);
  const bool linear = fabs(fr - 0.5) > 1e-6;
  const int r0 = matlab_round(p) - hwl;
  double mean_f0 = 0.0;
  // GetTentativeF0 (:122-131) and the 20 % sanity check of GetRefinedF0 (:203-204)
  if (linear) {
    const double tentative = stonemask_fix_f0_dft<2, true>(x, x_len, fs, t_pos, hwl, log2fft, wlen, r0, f0, lane, tw_c_base);
    if (!(tentative <= 0.0 || tentative > f0 * 2))
      mean_f0 = stonemask_fix_f0_dft<6, true>(x, x_len, fs, t_pos, hwl, log2fft, wlen, r0, tentative, lane, tw_c_base);
  } else {
    mean_f0 = stonemask_exact_indices(x, x_len, fs, t_pos, hwl, log2fft, wlen, f0, lane, tw_c_base);
  }
  if (fabs(mean_f0 - f0) / f0 > 0.2) mean_f0 = f0;
  if (lane == 0) f0_out[f] = mean_f0;
}

}  // namespace

#ifndef WB_HOST_EMU      // the launcher; tests/emu has its own
bool stonemask_run(const UttView& u, int fs, int total_frames, const int* frame_utt,
                   const double* frame_t, const double* f0_in, double* f0_out) {
  Context* c = ctx();
  if (!c) return false;
  if (total_frames <= 0) return true;
  cudaStream_t st = c->stream;
  if (option("stonemask_dft")) {
    // StoneMask -> CheapTrick -> D4C is the longest stretch of a pass without a single host round trip (~60 % of
    // it): the deferred bulk copies of a pipelined caller start here (wb200_set_copy_deferral, wb_api.cu)
    flush_deferred_copies();
    // largest transform size the bins refer to: f0 just above 40 Hz -> hwl = 1.5 fs / 40 + 1
    const int w_max = 2 * static_cast<int>(1.5 * fs / kFloorF0StoneMask + 1.0) + 1;
    int l2 = 0; while ((1 << (l2 + 1)) <= w_max) ++l2;
    if (2 + l2 > kTwLog2) { set_error("StoneMask: sampling rate %d not supported (bin table 2^%d)", fs, 2 + l2); return false; }
    KernelTimer kt("stonemask_kernel");
    stonemask_dft_kernel<<<(total_frames + kSmWarps - 1) / kSmWarps, kSmWarps * 32, 0, st>>>(u, frame_utt, frame_t, f0_in, c->d_twiddle_c, fs,
                                                                                         total_frames, f0_out);
    WB_LAUNCH_CHECK(); kt.stop();
    return true;
  }
  DevBuf<int> d_max;
  if (!d_max.alloc(1)) return false;
  if (!dev_fill(d_max.p, 0, sizeof(int))) return false;
  stonemask_maxfft_kernel<<<std::min(1024, (total_frames + 255) / 256), 256, 0, st>>>(f0_in, total_frames, fs, d_max.p);
  WB_LAUNCH_CHECK();
  int h_max = 0;
  if (!read_back(&h_max, d_max.p, sizeof(int))) return false;
  if (h_max < 3) h_max = 3;
  if (h_max > 13) { set_error("StoneMask: FFT size 2^%d not supported", h_max); return false; }
  const size_t smem = ((cpad_size(1 << h_max) + 1) & ~1) * sizeof(float2) + ((size_t)(1 << h_max) / 2 + 8) * sizeof(double) + ((size_t)(1 << h_max) / 2) * sizeof(int);
  if (smem > c->smem_optin) { set_error("StoneMask: needs %zu bytes of shared memory", smem); return false; }
  WB_CUDA_OR_RETURN(cudaFuncSetAttribute(stonemask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), false);
  KernelTimer kt1("stonemask_kernel");
  stonemask_kernel<<<total_frames, 256, smem, st>>>(u, frame_utt, frame_t, f0_in, c->d_twiddle_cf, fs, h_max, f0_out);
  WB_LAUNCH_CHECK(); kt1.stop();
  return true;
}

#endif  // WB_HOST_EMU

}  // namespace wb
