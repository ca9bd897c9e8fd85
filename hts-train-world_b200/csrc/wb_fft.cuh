// world-b200: shared-memory FFT for one CTA (sizes 2^3 .. 2^13 complex points), FP64 or FP32.
//
// The reference computes every transform with Ooura's split-radix code on one CPU thread
// (W/src/fft.cpp).  Only its conventions matter here (W/src/fft.cpp:26-74):
//   r2c : X[k] = sum_n x[n] exp(-2 pi i k n / N), k = 0..N/2
//   c2r : x[n] = sum_{k=0}^{N-1} X[k] exp(+2 pi i k n / N)  (unnormalised; Im X[0], Im X[N/2] ignored)
//
// Design: decimation in time, in place.  The caller stores element c at slot brev(c); after
// fft_dit() slot k holds X[k] in natural order.  Each pass keeps 2^K points (K <= 4) per
// thread in registers and performs K radix-2 stages before touching shared memory again, so
// a 4096-point transform makes 3 (radix-16) or 4 (radix-8) round trips through shared memory
// instead of 12.  Slots are padded (cpad) so that the butterfly passes and the bit-reversed
// input / output permutations are bank-conflict-free.  Twiddles come from one global table
// per precision (L1-resident, read-only path), K loads per group.  Size, pass schedule and
// block size are template parameters: all offsets inside a group are immediates.
//
// The element type C is double2 (everything whose result feeds a cancellation: power
// spectra, cumulative sums, group delays) or float2 (transforms of log spectra, cepstra,
// noise and band-limited group-delay slices, where 2^-24 relative to the largest element is
// far inside the parity tolerances; DESIGN.md section 4 lists which transform uses which).
#pragma once
#include "wb_common.cuh"

namespace wb {

// Slot padding: one extra element every 8, 64 and 512 elements.  For 16-byte elements a
// quarter-warp (8 threads = one 128-byte shared-memory wavefront) is conflict-free when its 8
// slots differ modulo 8; with the three skew terms that holds for the unit- and 8-stride
// accesses of the butterfly passes AND for the bit-reversed scatter / gather of the input and
// output permutations (strides 2^(log2n-3) * {0,4,2,6,1,5,3,7}) for every size 2^6 .. 2^13.
__host__ __device__ __forceinline__ constexpr int cpad(int c) { return c + (c >> 3) + (c >> 6) + (c >> 9); }
__host__ __device__ constexpr int cpad_size(int n) { return n + (n >> 3) + (n >> 6) + (n >> 9) + 4; }
// 8-byte elements (float2): 16 slots per 128-byte wavefront, so the skew is one element every
// 16, 256 and 4096; conflict-free for the radix-16 / radix-8 first pass, for later passes whose
// stage is >= 4 (16 consecutive lanes) and for the bit-reversed permutations.
__host__ __device__ __forceinline__ constexpr int cpadf(int c) { return c + (c >> 4) + (c >> 8) + (c >> 12); }
template <typename C> __host__ __device__ __forceinline__ constexpr int cpadT(int c) {
  return sizeof(C) == 16 ? cpad(c) : cpadf(c);
}
__device__ __forceinline__ int brev(int c, int log2n) {
  return static_cast<int>(__brev(static_cast<unsigned>(c)) >> (32 - log2n));
}

template <typename C> struct scalar_of;
template <> struct scalar_of<double2> { using type = double; };
template <> struct scalar_of<float2> { using type = float; };
template <typename C> using scalar_t = typename scalar_of<C>::type;

__device__ __forceinline__ double2 mk2(double x, double y) { return make_double2(x, y); }
__device__ __forceinline__ float2 mk2(float x, float y) { return make_float2(x, y); }

template <typename C> __device__ __forceinline__ C cmul(C a, C b) { return mk2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <typename C> __device__ __forceinline__ C cadd(C a, C b) { return mk2(a.x + b.x, a.y + b.y); }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { return mk2(a.x - b.x, a.y - b.y); }
template <typename C> __device__ __forceinline__ C cconj(C a) { return mk2(a.x, -a.y); }

// v * exp(-/+ 2 pi i e16 / 16) for e16 in [0, 8) (forward: minus sign; INV: plus sign).
// e16 is a compile-time constant after unrolling, so only one branch survives.
template <bool INV, typename C>
__device__ __forceinline__ C rot16(C v, int e16) {
  using R = scalar_t<C>;
  constexpr R kH = static_cast<R>(0.70710678118654752440);
  constexpr R kC = static_cast<R>(0.92387953251128675613);   // cos(pi/8)
  constexpr R kS = static_cast<R>(0.38268343236508977173);   // sin(pi/8)
  switch (e16) {
    case 0: return v;
    case 4: return INV ? mk2(-v.y, v.x) : mk2(v.y, -v.x);
    case 2: return INV ? mk2((v.x - v.y) * kH, (v.x + v.y) * kH) : mk2((v.x + v.y) * kH, (v.y - v.x) * kH);
    case 6: return INV ? mk2(-(v.x + v.y) * kH, (v.x - v.y) * kH) : mk2((v.y - v.x) * kH, -(v.x + v.y) * kH);
    default: {
      // exp(-i e16 pi / 8) = (cr, -ci) forward, (cr, +ci) inverse
      const R cr = e16 == 1 ? kC : e16 == 3 ? kS : e16 == 5 ? -kS : -kC;
      const R ci = (e16 == 1 || e16 == 7) ? kS : kC;
      const R sy = INV ? ci : -ci;
      return mk2(v.x * cr - v.y * sy, v.x * sy + v.y * cr);
    }
  }
}

// ---- butterflies in fused multiply-add form ------------------------------------------------------
// a <- a + W b, b <- a - W b with SIX fused multiply-adds (Linzer & Feig): the sum is two FMAs per
// component and the difference is 2a - (a + W b), one more.  The plain form (complex product, then
// add and subtract) is eight instructions, four of them additions the FMA pipe cannot fuse; in FP64
// every instruction is two issue cycles of the half-rate pipe, so this is -25 % on the arithmetic of
// every twiddled pass.  The rounding of the difference is relative to |a + W b| instead of |a - W b|:
// the same absolute bound (one ulp of the larger of the two) the usual FFT error analysis assumes.
template <typename C>
__device__ __forceinline__ void bfly_w(C& a, C& b, const C W) {
  using R = scalar_t<C>;
  const R px = fma(W.x, b.x, fma(-W.y, b.y, a.x));
  const R py = fma(W.x, b.y, fma(W.y, b.x, a.y));
  b = mk2(fma(static_cast<R>(2), a.x, -px), fma(static_cast<R>(2), a.y, -py));
  a = mk2(px, py);
}
// the same with W = exp(-/+ 2 pi i e16 / 16), a compile-time constant: multiples of a quarter turn
// are plain additions, everything else the six-FMA form with constant operands
template <bool INV, typename C>
__device__ __forceinline__ void bfly_const(C& a, C& b, int e16) {
  using R = scalar_t<C>;
  if (e16 == 0 || e16 == 4) {
    const C x = rot16<INV>(b, e16);
    b = csub(a, x);
    a = cadd(a, x);
  } else {
    bfly_w(a, b, rot16<INV>(mk2(static_cast<R>(1), static_cast<R>(0)), e16));
  }
}

// Stage T_ of a pass on 2^K points in registers.  Butterfly (m, m + 2^T_) multiplies its second input
// by W_j = w[T_] exp(-/+ 2 pi i j / 2^(T_+1)), j = m mod 2^T_: the group's twiddle times a multiple of
// 22.5 degrees.  The W_j of a stage are formed ONCE (a constant rotation of the twiddle: 4 operations,
// and none at all for the second half, which is the first half times -/+ i) and each butterfly is
// then six FMAs; the first pass has w = 1 and its W_j are constants.
template <int T_, int K, bool INV, bool FIRST, typename C>
__device__ __forceinline__ void fft_stage(C (&v)[1 << K], const C* w) {
  constexpr int R = 1 << K, span = 1 << T_;
  if constexpr (FIRST) {
#pragma unroll
    for (int m = 0; m < R; ++m) {
      if (m & span) continue;
      bfly_const<INV>(v[m], v[m + span], (m & (span - 1)) * (8 >> T_));
    }
  } else {
    C W[span];
    W[0] = w[T_];
#pragma unroll
    for (int j = 1; j < span; ++j) {
      const int e16 = j * (8 >> T_);
      W[j] = e16 >= 4 ? rot16<INV>(W[j - span / 2 > 0 ? j - span / 2 : 0], 4) : rot16<INV>(w[T_], e16);
    }
#pragma unroll
    for (int m = 0; m < R; ++m) {
      if (m & span) continue;
      bfly_w(v[m], v[m + span], W[m & (span - 1)]);
    }
  }
}
template <int K, bool INV, bool FIRST, int T_ = 0, typename C>
__device__ __forceinline__ void fft_stages(C (&v)[1 << K], const C* w) {
  if constexpr (T_ < K) {
    fft_stage<T_, K, INV, FIRST>(v, w);
    fft_stages<K, INV, FIRST, T_ + 1>(v, w);
  }
}

// One pass of K radix-2 stages on 2^K points held in registers.  The finest of the K twiddles of the
// group (W_{2^(STAGE+t+1)}^j, t < K) is loaded, the coarser ones are its squares; the other butterflies
// of a sub-stage use the same twiddle times a multiple of 22.5 degrees (fft_stage).  The first pass
// (STAGE 0) has j = 0 for every group and needs no table twiddle at all.
//
// cpad() is additive over non-overlapping bit fields: the group base has zeros where
// (m << STAGE) lives, hence cpad(base + (m << STAGE)) = cpad(base) + cpad(m << STAGE).
// The K twiddles of group b of the pass at STAGE: one table load, w[K-1] = exp(-2 pi i j / 2^(STAGE+K)) with
// j = b mod 2^STAGE is the finest, and every coarser one is the square of the next finer one -- 4 flops instead
// of a load whose latency the whole group waits for (with ~28 KB of L1 left beside the shared memory the
// tables miss; measured -7 % on the step).  The rounding error doubles per squaring: <= 2^(K-1) ulp.
// They depend on the thread index only, NOT on the data, so fft_run_passes forms the twiddles of pass p + 1
// between the stores of pass p and the barrier.  (Measured: no gain -- ptxas sinks the load behind the
// barrier again, also when it is a volatile asm; the stall samples that sit on this load are the barrier
// wait itself, which ncu attributes to the first instruction behind BAR.SYNC.)
template <int K, bool INV, int STAGE, int TWL, typename C>
__device__ __forceinline__ void fft_twiddles(const C* __restrict__ tw, int b, C (&w)[4]) {
  if constexpr (STAGE > 0) {
    const int j = b & ((1 << STAGE) - 1);
    w[K - 1] = __ldg(&tw[j << (TWL - STAGE - K)]);
    if (INV) w[K - 1].y = -w[K - 1].y;
#pragma unroll
    for (int t = K - 2; t >= 0; --t)
      w[t] = mk2((w[t + 1].x - w[t + 1].y) * (w[t + 1].x + w[t + 1].y), (w[t + 1].x + w[t + 1].x) * w[t + 1].y);
  }
}

// w0: the twiddles of this thread's first group (fft_twiddles with b = threadIdx.x), formed by the caller.
template <int K, bool INV, int LOG2N, int STAGE, int THREADS, int TWL = kTwLog2, typename C>
__device__ __forceinline__ void fft_pass(C* __restrict__ s, const C* __restrict__ tw, const C (&w0)[4]) {
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
    C* __restrict__ sb = s + cpadT<C>(base);
    C v[R];
#pragma unroll
    for (int m = 0; m < R; ++m) v[m] = sb[cpadT<C>(m << STAGE)];
    if (it == 0) {
      fft_stages<K, INV, FIRST>(v, w0);
    } else {
      C w[4];
      fft_twiddles<K, INV, STAGE, TWL>(tw, b, w);
      fft_stages<K, INV, FIRST>(v, w);
    }
#pragma unroll
    for (int m = 0; m < R; ++m) sb[cpadT<C>(m << STAGE)] = v[m];
  }
}

// First pass with the input produced on the fly: load(slot) returns the element that belongs in
// `slot` (element index brev(slot)), e.g. a normalised / zero-padded value computed from what the
// slot currently holds -- a separate sweep over the whole buffer and its barrier disappear.  The
// group of a thread is its own 2^K consecutive slots, so reading and writing them needs no
// synchronisation beyond "the producers of the slots are done".
template <int K, bool INV, int LOG2N, int THREADS, typename C, typename LOAD>
__device__ __forceinline__ void fft_first_pass_from(C* __restrict__ s, LOAD load) {
  constexpr int R = 1 << K;
  constexpr int NB = 1 << (LOG2N - K);
  constexpr int ITERS = (NB + THREADS - 1) / THREADS;
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int b = threadIdx.x + it * THREADS;
    if (NB % THREADS != 0 && b >= NB) break;
    const int base = b << K;
    C* __restrict__ sb = s + cpadT<C>(base);
    C v[R];
#pragma unroll
    for (int m = 0; m < R; ++m) v[m] = load(base, m);
    fft_stages<K, INV, true>(v, static_cast<const C*>(nullptr));
#pragma unroll
    for (int m = 0; m < R; ++m) sb[cpadT<C>(m)] = v[m];
  }
}

// Pass schedule: P = ceil(LOG2N / MAXK) passes, the first LOG2N % P of them one stage larger.
template <int LOG2N, int MAXK> struct fft_plan {
  static constexpr int P = (LOG2N + MAXK - 1) / MAXK;
  static constexpr int base = LOG2N / P, extra = LOG2N % P;
  __host__ __device__ static constexpr int k_of(int p) { return base + (p < extra ? 1 : 0); }
  __host__ __device__ static constexpr int stage_of(int p) { return p * base + (p < extra ? p : extra); }
};

// The barrier behind pass PASS.  Passes 0 and 1 of equal radix 2^K touch the same slots FROM THE SAME WARP:
// thread b owns slots [b 2^K, (b + 1) 2^K) in pass 0 and column (b mod 2^K) of the 2^2K-slot block b >> K
// in pass 1, so the 32 threads of a warp cover the 32 * 2^K consecutive slots starting at 32 w 2^K both times
// (every iteration of a multi-iteration pass likewise): a warp barrier orders them, the block barrier
// (and the wait for the slowest warp of the CTA) is not needed there.
template <int LOG2N, int MAXK, int PASS, int THREADS>
__device__ __forceinline__ void fft_sync_after() {
  using plan = fft_plan<LOG2N, MAXK>;
  if constexpr (PASS == 0 && plan::P > 1 && plan::k_of(0) == plan::k_of(plan::P > 1 ? 1 : 0) && plan::k_of(0) <= 5 && THREADS % 32 == 0) {
#ifdef WB_FFT_BLOCK_SYNC
    __syncthreads();
#else
    __syncwarp();
#endif
  } else {
    __syncthreads();
  }
}

// passes PASS .. P-1; w = the twiddles of pass PASS for this thread's first group (ignored by pass 0)
template <int LOG2N, int MAXK, int PASS, bool INV, int THREADS, int TWL = kTwLog2, typename C>
__device__ __forceinline__ void fft_run_passes_w(C* s, const C* __restrict__ tw, const C (&w)[4]) {
  using plan = fft_plan<LOG2N, MAXK>;
  if constexpr (PASS < plan::P) {
    fft_pass<plan::k_of(PASS), INV, LOG2N, plan::stage_of(PASS), THREADS, TWL>(s, tw, w);
    C wn[4];
    if constexpr (PASS + 1 < plan::P)
      fft_twiddles<plan::k_of(PASS + 1 < plan::P ? PASS + 1 : PASS), INV, plan::stage_of(PASS + 1 < plan::P ? PASS + 1 : PASS), TWL>(tw, threadIdx.x, wn);
    fft_sync_after<LOG2N, MAXK, PASS, THREADS>();
    fft_run_passes_w<LOG2N, MAXK, PASS + 1, INV, THREADS, TWL>(s, tw, wn);
  }
}
template <int LOG2N, int MAXK, int PASS, bool INV, int THREADS, int TWL = kTwLog2, typename C>
__device__ __forceinline__ void fft_run_passes(C* s, const C* __restrict__ tw) {
  using plan = fft_plan<LOG2N, MAXK>;
  C w[4];
  if constexpr (PASS > 0 && PASS < plan::P)
    fft_twiddles<plan::k_of(PASS < plan::P ? PASS : 0), INV, plan::stage_of(PASS < plan::P ? PASS : 0), TWL>(tw, threadIdx.x, w);
  fft_run_passes_w<LOG2N, MAXK, PASS, INV, THREADS, TWL>(s, tw, w);
}
// The passes after a first pass the caller ran itself (fft_first_pass_from, or a pruned first pass of its own):
// the twiddles of pass 1 are formed BEFORE the barrier behind pass 0.  WARP_LOCAL: the caller's first pass
// used the thread <-> slot mapping of fft_pass (fft_sync_after may relax the barrier to a warp barrier).
template <int LOG2N, int MAXK, bool INV, int THREADS, int TWL = kTwLog2, bool WARP_LOCAL = true, typename C>
__device__ __forceinline__ void fft_finish_after_first_pass(C* s, const C* __restrict__ tw) {
  using plan = fft_plan<LOG2N, MAXK>;
  C w[4];
  if constexpr (plan::P > 1)
    fft_twiddles<plan::k_of(plan::P > 1 ? 1 : 0), INV, plan::stage_of(plan::P > 1 ? 1 : 0), TWL>(tw, threadIdx.x, w);
  if constexpr (WARP_LOCAL) fft_sync_after<LOG2N, MAXK, 0, THREADS>();
  else __syncthreads();
  fft_run_passes_w<LOG2N, MAXK, 1, INV, THREADS, TWL>(s, tw, w);
}

// In-place complex FFT of 2^LOG2N points held in shared memory at padded slots.
// Input: element c stored at slot brev(c, LOG2N).  Output: slot k = X[k].
// Starts and ends with __syncthreads().
template <int LOG2N, bool INV, int THREADS, int MAXK, int TWL = kTwLog2, typename C>
__device__ __forceinline__ void fft_dit_fixed(C* s, const C* __restrict__ tw) {
  __syncthreads();
  fft_run_passes<LOG2N, MAXK, 0, INV, THREADS, TWL>(s, tw);
}

// Size chosen at run time (block-uniform): sizes 2^3 .. 2^13.
template <bool INV, int THREADS, int MAXK, typename C>
__device__ __noinline__ void fft_dit_rt(C* s, int log2n, const C* __restrict__ tw) {
  switch (log2n) {
    case 3: fft_dit_fixed<3, INV, THREADS, MAXK>(s, tw); break;
    case 4: fft_dit_fixed<4, INV, THREADS, MAXK>(s, tw); break;
    case 5: fft_dit_fixed<5, INV, THREADS, MAXK>(s, tw); break;
    case 6: fft_dit_fixed<6, INV, THREADS, MAXK>(s, tw); break;
    case 7: fft_dit_fixed<7, INV, THREADS, MAXK>(s, tw); break;
    case 8: fft_dit_fixed<8, INV, THREADS, MAXK>(s, tw); break;
    case 9: fft_dit_fixed<9, INV, THREADS, MAXK>(s, tw); break;
    case 10: fft_dit_fixed<10, INV, THREADS, MAXK>(s, tw); break;
    case 11: fft_dit_fixed<11, INV, THREADS, MAXK>(s, tw); break;
    case 12: fft_dit_fixed<12, INV, THREADS, MAXK>(s, tw); break;
    case 13: fft_dit_fixed<13, INV, THREADS, MAXK>(s, tw); break;
    default: break;
  }
}

// Run-time size with the compact tables: `tw_c_base` is the start of the concatenated per-size
// tables (Context::d_twiddle_c / d_twiddle_cf); every case picks its own table.
template <bool INV, int THREADS, int MAXK, typename C>
__device__ __noinline__ void fft_dit_rt_compact(C* s, int log2n, const C* __restrict__ tw_c_base) {
#define WB_FFT_CASE(L) case L: fft_dit_fixed<L, INV, THREADS, MAXK, L>(s, tw_c_base + Context::tw_c_offset(L)); break;
  switch (log2n) {
    WB_FFT_CASE(4) WB_FFT_CASE(5) WB_FFT_CASE(6) WB_FFT_CASE(7) WB_FFT_CASE(8) WB_FFT_CASE(9)
    WB_FFT_CASE(10) WB_FFT_CASE(11) WB_FFT_CASE(12) WB_FFT_CASE(13)
    default: break;
  }
#undef WB_FFT_CASE
}

// LOG2N > 0: compile-time size; LOG2N == 0: the run-time value log2n_rt (master table only).
// TWL: `tw` holds exp(-2 pi i k / 2^TWL) -- the master table (kTwLog2) or a compact one.
template <int LOG2N, bool INV, int THREADS, int MAXK = 3, int TWL = kTwLog2, typename C>
__device__ __forceinline__ void fft_dit(C* s, int log2n_rt, const C* __restrict__ tw) {
  if constexpr (LOG2N > 0) fft_dit_fixed<LOG2N, INV, THREADS, MAXK, TWL>(s, tw);
  else { static_assert(LOG2N > 0 || TWL == kTwLog2, "run-time sizes use the master table"); fft_dit_rt<INV, THREADS, MAXK>(s, log2n_rt, tw); }
}

// ---- real transforms on top of a half-size complex FFT -----------------------------------
// Forward: pack x[2n] + i x[2n+1] into element n (slot brev(n)), run fft_dit<false> with
// log2m = log2(N) - 1, then rfft_bin(k) returns X[k] for k in [0, N/2].
template <int TWL = kTwLog2, typename C>
__device__ __forceinline__ C rfft_bin(const C* s, int log2m, int k, const C* __restrict__ tw) {
  using R = scalar_t<C>;
  const int M = 1 << log2m;
  if (k == 0 || k == M) {
    const C z0 = s[0];
    return mk2(k == 0 ? z0.x + z0.y : z0.x - z0.y, static_cast<R>(0));
  }
  const C A = s[cpadT<C>(k)];
  const C B = cconj(s[cpadT<C>(M - k)]);
  const R h = static_cast<R>(0.5);
  const C E = mk2(h * (A.x + B.x), h * (A.y + B.y));
  const C O = mk2(h * (A.x - B.x), h * (A.y - B.y));
  const C w = __ldg(&tw[k << (TWL - log2m - 1)]);
  const C t = cmul(w, O);
  return mk2(E.x + t.y, E.y - t.x);     // E - i w O
}

// Where the real sample with index i (0 <= i < N) lives (as an index into the shared array
// viewed as scalars) before a forward real transform / after an inverse one.
__device__ __forceinline__ int rfft_in_slot(int i, int log2m) {
  return 2 * cpad(brev(i >> 1, log2m)) + (i & 1);
}
__device__ __forceinline__ int rfft_out_slot(int i) {   // natural order after c2r
  return 2 * cpad(i >> 1) + (i & 1);
}
// the same for float2 buffers
__device__ __forceinline__ int rfft_in_slot_f(int i, int log2m) { return 2 * cpadf(brev(i >> 1, log2m)) + (i & 1); }
__device__ __forceinline__ int rfft_out_slot_f(int i) { return 2 * cpadf(i >> 1) + (i & 1); }

// Inverse (c2r): for k in [0, N/2) compute the packed element from X[k] and X[N/2 - k]
// and store it at slot brev(k); then fft_dit<true>; real sample i is at rfft_out_slot(i).
template <int TWL = kTwLog2, typename C>
__device__ __forceinline__ C c2r_pack(C Xk, C XMk, int k, int log2m, const C* __restrict__ tw) {
  if (k == 0)   // Xk = X[0], XMk = X[N/2]; imaginary parts ignored like W/src/fft.cpp:27-29
    return mk2(Xk.x + XMk.x, Xk.x - XMk.x);
  const C B = cconj(XMk);
  const C S = cadd(Xk, B);
  const C D = csub(Xk, B);
  C w = __ldg(&tw[k << (TWL - log2m - 1)]);
  w.y = -w.y;                                     // w^{-k}
  const C t = cmul(w, D);
  return mk2(S.x - t.y, S.y + t.x);      // S + i w^{-k} D
}

}  // namespace wb
