// world-b200: block-cooperative spectral helpers shared by CheapTrick and D4C.
// Reference: W/src/common.cpp DCCorrection :56-75, LinearSmoothing :77-111 (+ :27-46).
#pragma once
#include "wb_common.cuh"

namespace wb {

// In-place DC correction of spec[0..N/2] held in shared memory (common.cpp:56-75): the bins below f0 receive
// the mirror image of the bins above it, spec[i] += interp1Q(spec)(f0 - i df) for i < upper_limit - 1.
// The <= ~70 bins involved (f0 N / fs + 1) are the work of ONE warp: warp 0 computes its values into
// registers, orders itself with a warp barrier and adds them; the other warps pass straight through, so a
// caller whose next step does not read the low bins (D4C's centroid, whose consumer is three phases away)
// pays nothing, and one that does pays one block barrier instead of two plus a round trip through scratch.
// `spec` must be complete on entry (caller syncs).  Does NOT end with a barrier.  Returns false (nothing
// done) when the range exceeds what a warp holds; the caller then uses dc_correction_block.
__device__ __forceinline__ bool dc_correction_warp(double* spec, double f0, int fs, int N) {
  const double inv_df = (double)N / fs;
  const double inv_n = 1.0 / N;                 // N is a power of two: x * inv_n == x / N exactly
  const int upper_limit = min(N / 2 - 1, 2 + static_cast<int>(mul_rn(f0, (double)N) / fs));
  const int n = upper_limit - 1;
  if (n > 128) return false;
  if (threadIdx.x < 32) {
    double v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = (int)threadIdx.x + 32 * j;
      v[j] = i < n ? interp1q_at(f0, -inv_df, spec, upper_limit + 1, mul_rn((double)i, (double)fs) * inv_n) : 0.0;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = (int)threadIdx.x + 32 * j;
      if (i < n) spec[i] += v[j];
    }
  }
  return true;
}
// The same by the whole block through scratch.  tmp: >= 2 + f0*N/fs doubles.  Ends with __syncthreads().
__device__ __forceinline__ void dc_correction_block(double* spec, double* tmp, double f0, int fs, int N) {
  const int T = blockDim.x, tid = threadIdx.x;
  const double inv_df = (double)N / fs;
  const double inv_n = 1.0 / N;
  const int upper_limit = min(N / 2 - 1, 2 + static_cast<int>(mul_rn(f0, (double)N) / fs));
  for (int i = tid; i < upper_limit - 1; i += T)
    tmp[i] = interp1q_at(f0, -inv_df, spec, upper_limit + 1, mul_rn((double)i, (double)fs) * inv_n);
  __syncthreads();
  for (int i = tid; i < upper_limit - 1; i += T) spec[i] += tmp[i];
  __syncthreads();
}
// BARRIER_AFTER: the corrected bins are read by other warps right away
template <bool BARRIER_AFTER>
__device__ __forceinline__ void dc_correction(double* spec, double* tmp, double f0, int fs, int N) {
  if (dc_correction_warp(spec, f0, fs, N)) {          // block-uniform
    if (BARRIER_AFTER) __syncthreads();
  } else {
    dc_correction_block(spec, tmp, f0, fs, N);
  }
}

__device__ __forceinline__ int smoothing_boundary(double width, int fs, int N) {
  return static_cast<int>(mul_rn(width, (double)N) / fs) + 1;
}

// cum[0..len) = running sum of the mirrored, df-scaled spectrum (LinearSmoothing :86-97):
// element i of the mirrored axis is in[boundary - i], in[i - boundary] or in[half - (i - (half +
// boundary))].  Every thread gathers its contiguous chunk (<= PER elements) straight from `in`
// into registers, sums it there, the chunk totals are scanned with warp shuffles and the
// finished values are stored once -- no mirrored copy and no second read-modify-write pass.
// Returns false (nothing written) when len > blockDim.x * PER; the caller then takes the generic
// path.  `in` must be complete on entry -- i.e. the caller has passed a block barrier since the last use
// of `red` as well, which is why no barrier protects `red` here; ends with __syncthreads().
template <int PER>
__device__ __forceinline__ bool mirrored_cumsum(const double* in, double* cum, double* red, int half,
                                                int boundary, int fs, double inv_n, int len_limit = 0x7fffffff) {
  const int T = blockDim.x, tid = threadIdx.x;
  const int len = min(half + 2 * boundary + 1, len_limit);      // a prefix of the running sum is all a caller with n_out < half needs
  const int per = (len + T - 1) / T;
  if (per > PER) return false;
  const int lo = tid * per;
  double v[PER];
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = lo + j;
    if (j < per && i < len) {
      const int src = i < boundary ? boundary - i : i < half + boundary ? i - boundary : half - (i - (half + boundary));
      s += mul_rn(in[src], (double)fs) * inv_n;
    }
    v[j] = s;
  }
  const int lane = tid & 31, wid = tid >> 5;
  double inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) red[wid] = inc;
  __syncthreads();
  double offset = inc - s;               // exclusive within the warp
  for (int w = 0; w < wid; ++w) offset += red[w];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = lo + j;
    if (j < per && i < len) cum[i] = v[j] + offset;
  }
  __syncthreads();
  return true;
}

// Rectangular smoothing of in[0..N/2] with the given width (Hz) -> out[0..n_out]; out may
// alias in.  cum: scratch of >= N/2 + 2*boundary + 1 doubles; red: >= 33 doubles.
// out[k] is a difference of the running sum of the mirrored input at k + boundary - 1/2 -+ width / (2 df),
// i.e. it reads the sum up to element k + 3 boundary / 2 + 1 and the input up to k + boundary / 2 + 1: a
// caller that needs only the bins up to n_out < N/2 (D4C: nothing above the highest band is ever used)
// gets the same values from a prefix -- smoothing_input_need(n_out, boundary) input bins.
// Must be entered by all threads with `in` complete (caller syncs); ends with __syncthreads().
// Not inlined: D4C calls it three times per frame and the kernels are instruction-cache bound.
__host__ __device__ __forceinline__ int smoothing_input_need(int n_out, int boundary) { return n_out + boundary + 3; }
static __device__ __noinline__ void linear_smoothing(const double* in, double* out, double* cum,
                                              double* red, double width, int fs, int N, int n_out) {
  const int T = blockDim.x, tid = threadIdx.x;
  const int half = N / 2;
  const int boundary = smoothing_boundary(width, fs, N);
  n_out = min(n_out, half);
  const int len = min(half + 2 * boundary + 1, n_out + 2 * boundary + 3);
  const double inv_df = (double)N / fs;
  const double inv_n = 1.0 / N;                 // N is a power of two: x * inv_n == x / N exactly
  if (!mirrored_cumsum<9>(in, cum, red, half, boundary, fs, inv_n, len)) {
    for (int i = tid; i < len; i += T) {
      double v;
      if (i < boundary) v = in[boundary - i];
      else if (i < half + boundary) v = in[i - boundary];
      else v = in[half - (i - (half + boundary))];
      cum[i] = mul_rn(v, (double)fs) * inv_n;
    }
    __syncthreads();
    block_inclusive_scan(cum, len, red);
  }
  // out[k] = (cum(k + c1) - cum(k + c0)) / width with cum() the linear interpolant of the running sum:
  // the reference evaluates interp1Q at fa = k df - width / 2 on an axis that starts at -(boundary - 1/2) df,
  // i.e. at the position k + boundary - 1/2 - width / (2 df) -- the SAME offset for every k.  Its integer
  // part and fraction are therefore computed once per call instead of twice per bin (two double -> int
  // conversions, two int -> double, four more FP64 operations each); the reference's own rounding of the
  // position moves the fraction by ~1e-13, and where that crosses an integer the interpolant is continuous.
  const double c0 = (boundary - 0.5) - 0.5 * width * inv_df;
  const double c1 = c0 + width * inv_df;
  const int i0 = static_cast<int>(c0), i1 = static_cast<int>(c1);      // c0 > 0: boundary > width / df
  const double f0 = c0 - i0, f1 = c1 - i1;
  const double inv_width = 1.0 / width;
  for (int k = tid; k <= n_out; k += T) {
    const int b0 = min(len - 1, k + i0), b1 = min(len - 1, k + i1);   // memory guards (as interp1q_at)
    const double l0 = cum[b0], h0 = cum[b1];
    const double low = fma(b0 + 1 < len ? cum[b0 + 1] - l0 : 0.0, f0, l0);
    const double high = fma(b1 + 1 < len ? cum[b1 + 1] - h0 : 0.0, f1, h0);
    out[k] = (high - low) * inv_width;
  }
  __syncthreads();
}

}  // namespace wb
