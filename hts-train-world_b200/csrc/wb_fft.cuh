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
// of 11.  Slots are padded (cpad) so that the butterfly passes and the bit-reversed input /
// output permutations are bank-conflict-free for 16-byte elements.  Twiddles come from one
// global table (L1-resident, read-only path), three loads per radix-8 group.
#pragma once
#include "wb_common.cuh"

namespace wb {

// Slot padding: one extra double2 every 8, 64 and 512 elements.  A quarter-warp (8 threads x 16
// bytes = one 128-byte shared-memory wavefront) is conflict-free when its 8 slots differ
// modulo 8; with the three skew terms that holds for the unit- and 8-stride accesses of the
// butterfly passes AND for the bit-reversed scatter / gather of the input and output
// permutations (strides 2^(log2n-3) * {0,4,2,6,1,5,3,7}) for every size from 2^6 to 2^13.
__host__ __device__ __forceinline__ constexpr int cpad(int c) { return c + (c >> 3) + (c >> 6) + (c >> 9); }
__host__ __device__ constexpr int cpad_size(int n) { return n + (n >> 3) + (n >> 6) + (n >> 9) + 4; }
__device__ __forceinline__ int brev(int c, int log2n) {
  return static_cast<int>(__brev(static_cast<unsigned>(c)) >> (32 - log2n));
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }

// v * exp(-/+ 2 pi i e8 / 8) for e8 in {0, 1, 2, 3} (forward: minus sign; INV: plus sign).
// e8 is a compile-time constant after unrolling, so only one branch survives.
template <bool INV>
__device__ __forceinline__ double2 rot8(double2 v, int e8) {
  constexpr double kH = 0.70710678118654752440;
  if (e8 == 0) return v;
  if (e8 == 2) return INV ? make_double2(-v.y, v.x) : make_double2(v.y, -v.x);
  if (e8 == 1) return INV ? make_double2((v.x - v.y) * kH, (v.x + v.y) * kH)
                          : make_double2((v.x + v.y) * kH, (v.y - v.x) * kH);
  return INV ? make_double2(-(v.x + v.y) * kH, (v.x - v.y) * kH)
             : make_double2((v.y - v.x) * kH, -(v.x + v.y) * kH);
}

// One pass of K radix-2 stages on 2^K points held in registers.  The K twiddles of the group
// (W_{2^(stage+t+1)}^j, t < K) are loaded once; the other butterflies of a sub-stage use the same
// twiddle times a multiple of 45 degrees, which costs at most two multiplies.  The first pass
// (stage 0) has j = 0 for every group and needs no twiddle at all.
template <int K, bool INV, bool FIRST>
__device__ __forceinline__ void fft_pass(double2* __restrict__ s, int log2n, int stage,
                                         const double2* __restrict__ tw) {
  constexpr int R = 1 << K;
  const int h_mask = (1 << stage) - 1;
  const int nb = 1 << (log2n - K);
  for (int b = threadIdx.x; b < nb; b += blockDim.x) {
    const int j = FIRST ? 0 : (b & h_mask);
    const int base = FIRST ? (b << K) : (((b >> stage) << (stage + K)) + j);
    double2 v[R];
#pragma unroll
    for (int m = 0; m < R; ++m) v[m] = s[cpad(base + (m << stage))];
    double2 w[K];
    if (!FIRST) {
#pragma unroll
      for (int t = 0; t < K; ++t) {
        w[t] = __ldg(&tw[j << (kTwLog2 - stage - t - 1)]);
        if (INV) w[t].y = -w[t].y;
      }
    }
#pragma unroll
    for (int t = 0; t < K; ++t) {
      const int span = 1 << t;
#pragma unroll
      for (int m = 0; m < R; ++m) {
        if (m & span) continue;
        double2 x = v[m + span];
        if (!FIRST) x = cmul(w[t], x);
        x = rot8<INV>(x, (m & (span - 1)) * (4 >> t));
        const double2 a = v[m];
        v[m] = cadd(a, x);
        v[m + span] = csub(a, x);
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
  int stage;
  const int rem = log2n % 3;
  if (rem == 1) { fft_pass<1, INV, true>(s, log2n, 0, tw); stage = 1; }
  else if (rem == 2) { fft_pass<2, INV, true>(s, log2n, 0, tw); stage = 2; }
  else { fft_pass<3, INV, true>(s, log2n, 0, tw); stage = 3; }
  __syncthreads();
  for (; stage < log2n; stage += 3) {
    fft_pass<3, INV, false>(s, log2n, stage, tw);
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
