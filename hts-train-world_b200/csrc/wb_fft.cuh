// world-b200: FP64 shared-memory FFT for one CTA (sizes 2^3 .. 2^13 complex points).
//
// The reference computes every transform with Ooura's split-radix code on one CPU thread
// (W/src/fft.cpp).  Only its conventions matter here (W/src/fft.cpp:26-74):
//   r2c : X[k] = sum_n x[n] exp(-2 pi i k n / N), k = 0..N/2
//   c2r : x[n] = sum_{k=0}^{N-1} X[k] exp(+2 pi i k n / N)  (unnormalised; Im X[0], Im X[N/2] ignored)
//
// Design: decimation in time, in place.  The caller stores element c at slot brev(c); after
// fft_dit() slot k holds X[k] in natural order.  Each pass keeps 8 (or 4 / 2) points per
// thread in registers and performs three (two / one) radix-2 stages before touching shared
// memory again, so a 2048-point transform makes 4 round trips through shared memory instead
// of 11.  Slots are padded by one double2 every 8 (cpad) so that the stride-8 accesses of
// the first pass and the stride-1 accesses of later passes are both bank-conflict-free for
// 16-byte elements.  Twiddles come from one global table (L1-resident, read-only path).
#pragma once
#include "wb_common.cuh"

namespace wb {

__host__ __device__ __forceinline__ constexpr int cpad(int c) { return c + (c >> 3); }
__host__ __device__ constexpr int cpad_size(int n) { return n + (n >> 3) + 2; }
__device__ __forceinline__ int brev(int c, int log2n) {
  return static_cast<int>(__brev(static_cast<unsigned>(c)) >> (32 - log2n));
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }

template <int K, bool INV>
__device__ __forceinline__ void fft_pass(double2* __restrict__ s, int log2n, int stage,
                                         const double2* __restrict__ tw) {
  constexpr int R = 1 << K;
  const int h_mask = (1 << stage) - 1;
  const int nb = 1 << (log2n - K);
  for (int b = threadIdx.x; b < nb; b += blockDim.x) {
    const int j = b & h_mask;
    const int base = ((b >> stage) << (stage + K)) + j;
    double2 v[R];
#pragma unroll
    for (int m = 0; m < R; ++m) v[m] = s[cpad(base + (m << stage))];
#pragma unroll
    for (int t = 0; t < K; ++t) {
      constexpr int dummy = 0; (void)dummy;
      const int span = 1 << t;
      const int sh = kTwLog2 - stage - t - 1;
#pragma unroll
      for (int m = 0; m < R; ++m) {
        if (m & span) continue;
        const int e = ((m & (span - 1)) << stage) + j;
        double2 w = __ldg(&tw[e << sh]);
        if (INV) w.y = -w.y;
        const double2 a = v[m];
        const double2 wb = cmul(w, v[m + span]);
        v[m] = cadd(a, wb);
        v[m + span] = csub(a, wb);
      }
    }
#pragma unroll
    for (int m = 0; m < R; ++m) s[cpad(base + (m << stage))] = v[m];
  }
}

// In-place complex FFT of 2^log2n points held in shared memory at padded slots.
// Input: element c stored at slot brev(c, log2n).  Output: slot k = X[k].
// Starts and ends with __syncthreads().
template <bool INV>
__device__ __forceinline__ void fft_dit(double2* s, int log2n, const double2* __restrict__ tw) {
  __syncthreads();
  int stage = 0;
  const int rem = log2n % 3;
  if (rem == 1) { fft_pass<1, INV>(s, log2n, 0, tw); stage = 1; __syncthreads(); }
  else if (rem == 2) { fft_pass<2, INV>(s, log2n, 0, tw); stage = 2; __syncthreads(); }
  for (; stage < log2n; stage += 3) {
    fft_pass<3, INV>(s, log2n, stage, tw);
    __syncthreads();
  }
}

// ---- real transforms on top of a half-size complex FFT -----------------------------------
// Forward: pack x[2n] + i x[2n+1] into element n (slot brev(n)), run fft_dit<false> with
// log2m = log2(N) - 1, then rfft_bin(k) returns X[k] for k in [0, N/2].
__device__ __forceinline__ double2 rfft_bin(const double2* s, int log2m, int k,
                                            const double2* __restrict__ tw) {
  const int M = 1 << log2m;
  if (k == 0 || k == M) {
    const double2 z0 = s[0];
    return make_double2(k == 0 ? z0.x + z0.y : z0.x - z0.y, 0.0);
  }
  const double2 A = s[cpad(k)];
  const double2 B = cconj(s[cpad(M - k)]);
  const double2 E = make_double2(0.5 * (A.x + B.x), 0.5 * (A.y + B.y));
  const double2 O = make_double2(0.5 * (A.x - B.x), 0.5 * (A.y - B.y));
  const double2 w = __ldg(&tw[k << (kTwLog2 - log2m - 1)]);
  const double2 t = cmul(w, O);
  return make_double2(E.x + t.y, E.y - t.x);     // E - i w O
}

// Where the real sample with index i (0 <= i < N) lives (as an index into the shared array
// viewed as doubles) before a forward real transform / after an inverse one.
__device__ __forceinline__ int rfft_in_slot(int i, int log2m) {
  return 2 * cpad(brev(i >> 1, log2m)) + (i & 1);
}
__device__ __forceinline__ int rfft_out_slot(int i) {   // natural order after c2r
  return 2 * cpad(i >> 1) + (i & 1);
}

// Inverse (c2r): for k in [0, N/2) compute the packed element from X[k] and X[N/2 - k]
// and store it at slot brev(k); then fft_dit<true>; real sample i is at rfft_out_slot(i).
__device__ __forceinline__ double2 c2r_pack(double2 Xk, double2 XMk, int k, int log2m,
                                            const double2* __restrict__ tw) {
  if (k == 0)   // Xk = X[0], XMk = X[N/2]; imaginary parts ignored like W/src/fft.cpp:27-29
    return make_double2(Xk.x + XMk.x, Xk.x - XMk.x);
  const double2 B = cconj(XMk);
  const double2 S = cadd(Xk, B);
  const double2 D = csub(Xk, B);
  double2 w = __ldg(&tw[k << (kTwLog2 - log2m - 1)]);
  w.y = -w.y;                                     // w^{-k}
  const double2 t = cmul(w, D);
  return make_double2(S.x - t.y, S.y + t.x);      // S + i w^{-k} D
}

}  // namespace wb
