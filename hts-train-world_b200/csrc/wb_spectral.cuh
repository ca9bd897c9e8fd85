// world-b200: block-cooperative spectral helpers shared by CheapTrick and D4C.
// Reference: W/src/common.cpp DCCorrection :56-75, LinearSmoothing :77-111 (+ :27-46).
#pragma once
#include "wb_common.cuh"

namespace wb {

// In-place DC correction of spec[0..N/2] held in shared memory.
// tmp: scratch of >= 2 + f0*N/fs doubles.  Ends with __syncthreads().
__device__ __forceinline__ void dc_correction(double* spec, double* tmp, double f0, int fs, int N) {
  const int T = blockDim.x, tid = threadIdx.x;
  const double inv_df = (double)N / fs;
  const double inv_n = 1.0 / N;                 // N is a power of two: x * inv_n == x / N exactly
  const int upper_limit = min(N / 2 - 1, 2 + static_cast<int>(mul_rn(f0, (double)N) / fs));
  for (int i = tid; i < upper_limit - 1; i += T)
    tmp[i] = interp1q_at(f0, -inv_df, spec, upper_limit + 1, mul_rn((double)i, (double)fs) * inv_n);
  __syncthreads();
  for (int i = tid; i < upper_limit - 1; i += T) spec[i] += tmp[i];
  __syncthreads();
}

__device__ __forceinline__ int smoothing_boundary(double width, int fs, int N) {
  return static_cast<int>(mul_rn(width, (double)N) / fs) + 1;
}

// Rectangular smoothing of in[0..N/2] with the given width (Hz) -> out[0..N/2]; out may
// alias in.  cum: scratch of >= N/2 + 2*boundary + 1 doubles; red: >= 33 doubles.
// Must be entered by all threads with `in` complete (caller syncs); ends with __syncthreads().
__device__ __forceinline__ void linear_smoothing(const double* in, double* out, double* cum,
                                                 double* red, double width, int fs, int N) {
  const int T = blockDim.x, tid = threadIdx.x;
  const int half = N / 2;
  const int boundary = smoothing_boundary(width, fs, N);
  const int len = half + 2 * boundary + 1;
  const double inv_df = (double)N / fs;
  const double inv_n = 1.0 / N;                 // N is a power of two: x * inv_n == x / N exactly
  for (int i = tid; i < len; i += T) {
    double v;
    if (i < boundary) v = in[boundary - i];
    else if (i < half + boundary) v = in[i - boundary];
    else v = in[half - (i - (half + boundary))];
    cum[i] = mul_rn(v, (double)fs) * inv_n;
  }
  __syncthreads();
  block_inclusive_scan(cum, len, red);
  const double origin_axis = -(boundary - 0.5) * fs / N;
  const double inv_width = 1.0 / width;
  for (int k = tid; k <= half; k += T) {
    const double fa = add_rn(mul_rn((double)k * inv_n, (double)fs), -width / 2.0);
    const double low = interp1q_at(origin_axis, inv_df, cum, len, fa);
    const double high = interp1q_at(origin_axis, inv_df, cum, len, add_rn(fa, width));
    out[k] = (high - low) * inv_width;
  }
  __syncthreads();
}

}  // namespace wb
