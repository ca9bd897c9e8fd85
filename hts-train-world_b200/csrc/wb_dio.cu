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
#include "wb_zerocross.cuh"

namespace wb {
namespace {

constexpr int kMaxDioBands = 16;

struct DioFilterBank {
  int nb = 0;
  int bn = 0, log2bn = 0;            // overlap-save block size (real points)
  int hN = 0;                        // low-cut half length
  int hal[kMaxDioBands];             // Nuttall half_average_length per band
  int D = 0, V = 0;                  // block pre-roll and valid outputs per block
  DevBuf<double2> G;                 // [nb][bn/2 + 1] spectra of the causal combined filters
  double boundary_f0[kMaxDioBands];
};

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
  // one CTA per utterance: keep eight loads in flight per thread, the latency of a dependent
  // load-add chain is what this kernel costs (the mean only shifts the DC the low-cut filter removes)
  double a[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  const int T = blockDim.x;
  int i = threadIdx.x;
  for (; i + 7 * T < n; i += 8 * T) {
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] += x[i + q * T];
  }
  for (; i < n; i += T) a[0] += x[i];
  double v[1] = {((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]))};
  block_sum<1>(v, red);
  if (threadIdx.x == 0) mean[u] = v[0] / y_len[u];
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
               int total_frames, double* tmp1, double* tmp2,          // exchanged between threads: not __restrict__
               int* __restrict__ pos_idx, int* __restrict__ neg_idx, double* f0_out) {
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

#ifndef WB_HOST_EMU      // the launcher; tests/emu has its own
bool dio_run(Batch* b, const DioParams& p, double* d_f0_out) {
  Context* ctxp = ctx();
  if (!ctxp) return false;
  cudaStream_t st = ctxp->stream;
  const int n_utt = b->n_utt;
  if (n_utt == 0) return true;
  const int ratio = std::max(std::min(p.speed, 12), 1);
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
  if (!write_dev(d_ylen.p, h_ylen.data(), n_utt * sizeof(int))) return false;
  if (!write_dev(d_mask.p, h_mask.data(), n_utt * sizeof(int))) return false;
  if (!dev_fill(d_f0_out, 0, (size_t)TF * sizeof(double))) return false;
  // the signal the bands are computed from: x itself, or its decimated copy when speed > 1 (:69-71)
  const double* xin = b->x.p;
  const long long* xin_off = b->x_off.p;
  const int* xin_len = b->x_len.p;
  DevBuf<double> d_dec;
  DevBuf<long long> d_dec_off;
  DevBuf<int> d_dec_len;
  if (ratio != 1) {
    std::vector<int> got;
    if (!decimate_run(b, ratio, h_ylen, &d_dec, &d_dec_off, &d_dec_len, &got)) return false;
    xin = d_dec.p; xin_off = d_dec_off.p; xin_len = d_dec_len.p;
  }
  dio_mean_kernel<<<n_utt, 1024, 0, st>>>(xin, xin_off, xin_len, d_ylen.p, d_mean.p);
  WB_LAUNCH_CHECK();

  const size_t smem = 2 * cpad_size(c.bn / 2) * sizeof(double2);
  WB_CUDA_OR_RETURN(cudaFuncSetAttribute(ols_filter_kernel<13, 512, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), false);
  WB_CUDA_OR_RETURN(cudaFuncSetAttribute(ols_filter_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), false);
  WB_CUDA_OR_RETURN(cudaFuncSetAttribute(ols_filter_zc_kernel<13, 512, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), false);
  WB_CUDA_OR_RETURN(cudaFuncSetAttribute(ols_filter_zc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), false);

  OlsConst oc = {c.nb, c.bn, c.log2bn, c.D, c.V};
  std::vector<int> h_shift(c.nb);
  for (int i = 0; i < c.nb; ++i) h_shift[i] = c.D + 2 * c.hal[i] + c.hN;
  DevBuf<int> d_shift;
  if (!d_shift.alloc(c.nb)) return false;
  if (!write_dev(d_shift.p, h_shift.data(), c.nb * sizeof(int))) return false;
  WB_CUDA_OR_RETURN(cudaStreamSynchronize(st), false);

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
      if (u1 > u0 && (tot + need > kMaxScratchDoubles || (long long)(u1 - u0 + 1) * c.nb > 65535)) break;   // grid.y of the per-(utterance, band) kernels
      h_foff.push_back((long long)tot);
      tot += need;
      max_y = std::max(max_y, h_ylen[u1]);
      ++u1;
    }
    const int nu = u1 - u0;
    DevBuf<double> d_F, d_edges;
    DevBuf<long long> d_foff, d_loff;
    DevBuf<int> d_counts, d_ltot;
    const int n_lists = nu * c.nb * 4;
    if (!d_ltot.alloc(n_lists + 1) || !d_loff.alloc(n_lists)) return false;
    std::vector<int> h_ltot(n_lists + 1);
    std::vector<long long> h_loff(n_lists);
    bool fused_done = false;
    if (option("dio_fused") && c.V > 64) {
      // filter + zero crossings in one kernel: the band signals never leave shared memory (wb_zerocross.cuh)
      OlsConst ocz = oc;
      ocz.V = c.V - 2;                                        // two-sample halo for the first differences
      const int n_blocks = (max_y - 1 + ocz.V - 1) / ocz.V;
      DevBuf<int> d_segcnt, d_segoff;
      DevBuf<double> d_seg;
      if (!d_segcnt.alloc((size_t)n_lists * n_blocks) || !d_segoff.alloc((size_t)n_lists * n_blocks) ||
          !d_seg.alloc((size_t)n_lists * n_blocks * kZcSegCap))
        return false;
      if (!dev_fill(d_segcnt.p, 0, (size_t)n_lists * n_blocks * sizeof(int))) return false;
      if (!dev_fill(d_ltot.p + n_lists, 0, sizeof(int))) return false;
      {
        KernelTimer kt1("dio_filter_kernel");
        if (c.log2bn == 13)
          ols_filter_zc_kernel<13, 512, 4><<<dim3(n_blocks, nu), 512, smem, st>>>(xin, xin_off, xin_len, d_ylen.p, d_mask.p, d_mean.p, fb->G.p,
                                                                                 ctxp->tw_c(13), ocz, d_shift.p, u0, n_blocks, kZcSegCap, d_segcnt.p, d_seg.p);
        else
          ols_filter_zc_kernel<0><<<dim3(n_blocks, nu), 256, smem, st>>>(xin, xin_off, xin_len, d_ylen.p, d_mask.p, d_mean.p, fb->G.p,
                                                                        ctxp->d_twiddle, ocz, d_shift.p, u0, n_blocks, kZcSegCap, d_segcnt.p, d_seg.p);
        WB_LAUNCH_CHECK(); kt1.stop();
      }
      zc_seg_scan_kernel<<<(n_lists + 127) / 128, 128, 0, st>>>(d_segcnt.p, n_lists, n_blocks, kZcSegCap, d_segoff.p, d_ltot.p, d_ltot.p + n_lists);
      WB_LAUNCH_CHECK();
      if (!read_back(h_ltot.data(), d_ltot.p, (n_lists + 1) * sizeof(int))) return false;
      if (h_ltot[n_lists] == 0) {                              // no segment overflowed
        long long etot = 0;
        for (int l = 0; l < n_lists; ++l) { h_loff[l] = etot; etot += h_ltot[l]; }
        if (!d_edges.alloc((size_t)etot + 2)) return false;
        if (!write_dev(d_loff.p, h_loff.data(), n_lists * sizeof(long long))) return false;
        KernelTimer kt3("dio_zc_kernel");
        zc_seg_gather_kernel<<<n_lists, 128, 0, st>>>(d_segcnt.p, d_segoff.p, d_seg.p, n_blocks, kZcSegCap, d_loff.p, d_edges.p);
        WB_LAUNCH_CHECK(); kt3.stop();
        WB_CUDA_OR_RETURN(cudaStreamSynchronize(st), false);  // h_loff and the segment buffers die with this scope
        fused_done = true;
      }
    }
    if (!fused_done) {
      const int n_blocks = (max_y + c.V - 1) / c.V;
      const int n_chunks = (max_y + kZcChunk - 1) / kZcChunk;
      if (!d_F.alloc(tot) || !d_foff.alloc(nu) || !d_counts.alloc((size_t)n_lists * n_chunks)) return false;
      if (!write_dev(d_foff.p, h_foff.data(), nu * sizeof(long long))) return false;
      KernelTimer kt1("dio_filter_kernel");
      // 150 KB of shared memory per block leave one CTA per SM: 512 threads (16 warps) hide the latency of
      // the multiply / pack / store sweeps even though only 256 of them own a radix-16 group (-16 %)
      if (c.log2bn == 13)
        ols_filter_kernel<13, 512, 4><<<dim3(n_blocks, nu), 512, smem, st>>>(xin, xin_off, xin_len, d_ylen.p, d_mask.p, d_mean.p,
                                                              d_foff.p, fb->G.p, ctxp->tw_c(13), oc, d_shift.p, u0, d_F.p);
      else
        ols_filter_kernel<0><<<dim3(n_blocks, nu), 256, smem, st>>>(xin, xin_off, xin_len, d_ylen.p, d_mask.p, d_mean.p,
                                                              d_foff.p, fb->G.p, ctxp->d_twiddle, oc, d_shift.p, u0, d_F.p);
      WB_LAUNCH_CHECK(); kt1.stop();
      KernelTimer kt2("dio_zc_kernel");
      zc_kernel<false><<<dim3(n_chunks, nu * c.nb), 256, 0, st>>>(d_F.p, d_foff.p, d_ylen.p, c.nb, u0, n_chunks, d_counts.p, nullptr, nullptr);
      WB_LAUNCH_CHECK(); kt2.stop();
      zc_scan_kernel<<<(n_lists + 127) / 128, 128, 0, st>>>(d_counts.p, n_lists, n_chunks, d_ltot.p);
      WB_LAUNCH_CHECK();
      if (!read_back(h_ltot.data(), d_ltot.p, n_lists * sizeof(int))) return false;
      long long etot = 0;
      for (int l = 0; l < n_lists; ++l) { h_loff[l] = etot; etot += h_ltot[l]; }
      if (!d_edges.alloc((size_t)etot + 2)) return false;
      if (!write_dev(d_loff.p, h_loff.data(), n_lists * sizeof(long long))) return false;
      KernelTimer kt3("dio_zc_kernel");
      zc_kernel<true><<<dim3(n_chunks, nu * c.nb), 256, 0, st>>>(d_F.p, d_foff.p, d_ylen.p, c.nb, u0, n_chunks, d_counts.p, d_loff.p, d_edges.p);
      WB_LAUNCH_CHECK(); kt3.stop();
    }
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

#endif  // WB_HOST_EMU

}  // namespace wb
