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

// One pass of K radix-2 stages on 2^K points held in registers.  Size, stage and block size
// are compile-time constants, so every shared-memory offset inside a group is an immediate and
// the groups of one thread are unrolled (their loads overlap).  The K twiddles of the group
// (W_{2^(STAGE+t+1)}^j, t < K) are loaded once; the other butterflies of a sub-stage use the same
// twiddle times a multiple of 45 degrees, which costs at most two multiplies.  The first pass
// (STAGE 0) has j = 0 for every group and needs no twiddle at all.
//
// cpad() is additive over non-overlapping bit fields: the group base has zeros where
// (m << STAGE) lives, hence cpad(base + (m << STAGE)) = cpad(base) + cpad(m << STAGE).
template <int K, bool INV, int LOG2N, int STAGE, int THREADS>
__device__ __forceinline__ void fft_pass(double2* __restrict__ s, const double2* __restrict__ tw) {
  constexpr int R = 1 << K;
  constexpr bool FIRST = STAGE == 0;
  constexpr int NB = 1 << (LOG2N - K);
  constexpr int ITERS = (NB + THREADS - 1) / THREADS;
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int b = threadIdx.x + it * THREADS;
    if (NB % THREADS != 0 && b >= NB) break;
    const int j = FIRST ? 0 : (b & ((1 << STAGE) - 1));
    const int base = FIRST ? (b << K) : (((b >> STAGE) << (STAGE + K)) + j);
    double2* __restrict__ sb = s + cpad(base);
    double2 v[R];
#pragma unroll
    for (int m = 0; m < R; ++m) v[m] = sb[cpad(m << STAGE)];
    double2 w[K];
    if (!FIRST) {
#pragma unroll
      for (int t = 0; t < K; ++t) {
        w[t] = __ldg(&tw[j << (kTwLog2 - STAGE - t - 1)]);
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
    for (int m = 0; m < R; ++m) sb[cpad(m << STAGE)] = v[m];
  }
}

template <int LOG2N, int STAGE, bool INV, int THREADS>
__device__ __forceinline__ void fft_later_passes(double2* s, const double2* __restrict__ tw) {
  if constexpr (STAGE < LOG2N) {
    fft_pass<3, INV, LOG2N, STAGE, THREADS>(s, tw);
    __syncthreads();
    fft_later_passes<LOG2N, STAGE + 3, INV, THREADS>(s, tw);
  }
}

// In-place complex FFT of 2^LOG2N points held in shared memory at padded slots.
// Input: element c stored at slot brev(c, LOG2N).  Output: slot k = X[k].
// Starts and ends with __syncthreads().
template <int LOG2N, bool INV, int THREADS>
__device__ __forceinline__ void fft_dit_fixed(double2* s, const double2* __restrict__ tw) {
  constexpr int K0 = LOG2N % 3 == 0 ? 3 : LOG2N % 3;
  __syncthreads();
  fft_pass<K0, INV, LOG2N, 0, THREADS>(s, tw);
  __syncthreads();
  fft_later_passes<LOG2N, K0, INV, THREADS>(s, tw);
}

// Size chosen at run time (block-uniform): sizes 2^3 .. 2^13.
template <bool INV, int THREADS>
__device__ __noinline__ void fft_dit_rt(double2* s, int log2n, const double2* __restrict__ tw) {
  switch (log2n) {
    case 3: fft_dit_fixed<3, INV, THREADS>(s, tw); break;
    case 4: fft_dit_fixed<4, INV, THREADS>(s, tw); break;
    case 5: fft_dit_fixed<5, INV, THREADS>(s, tw); break;
    case 6: fft_dit_fixed<6, INV, THREADS>(s, tw); break;
    case 7: fft_dit_fixed<7, INV, THREADS>(s, tw); break;
    case 8: fft_dit_fixed<8, INV, THREADS>(s, tw); break;
    case 9: fft_dit_fixed<9, INV, THREADS>(s, tw); break;
    case 10: fft_dit_fixed<10, INV, THREADS>(s, tw); break;
    case 11: fft_dit_fixed<11, INV, THREADS>(s, tw); break;
    case 12: fft_dit_fixed<12, INV, THREADS>(s, tw); break;
    case 13: fft_dit_fixed<13, INV, THREADS>(s, tw); break;
    default: break;
  }
}

// LOG2N > 0: compile-time size; LOG2N == 0: the run-time value log2n_rt.
template <int LOG2N, bool INV, int THREADS>
__device__ __forceinline__ void fft_dit(double2* s, int log2n_rt, const double2* __restrict__ tw) {
  if constexpr (LOG2N > 0) fft_dit_fixed<LOG2N, INV, THREADS>(s, tw);
  else fft_dit_rt<INV, THREADS>(s, log2n_rt, tw);
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
