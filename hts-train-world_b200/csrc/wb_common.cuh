// world-b200: shared host/device declarations for the sm_100a WORLD kernels.
//
// Layout conventions:
//   * a "batch" is a set of utterances resident in HBM; samples of all utterances are
//     concatenated in one double array, frames of all utterances are concatenated in one
//     frame table (utterance id, time position, f0), spectrogram / aperiodicity are dense
//     [total_frames][fft_size/2+1] double matrices.
//   * every per-frame (or per-pulse) kernel runs one CTA per frame (pulse) with the frame
//     staged in shared memory; FFTs are FP64 radix-8 shared-memory transforms (wb_fft.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

// dynamic shared memory of a kernel (tests/emu runs the kernels on the CPU, one CTA at a time)
#ifdef WB_HOST_EMU
#define WB_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(::wbemu::dyn_smem)
#else
#define WB_DYN_SMEM(T, name) extern __shared__ T name[]
#endif

namespace wb {

// ---- constants of the reference (W/src/world/constantnumbers.h) -------------------------
constexpr double kPi = 3.1415926535897932384;
constexpr double kMySafeGuardMinimum = 0.000000000001;
constexpr double kEps = 0.00000000000000022204460492503131;
constexpr double kFloorF0 = 71.0;
constexpr double kCeilF0 = 800.0;
constexpr double kDefaultF0 = 500.0;
constexpr double kLog2 = 0.69314718055994529;
constexpr double kMaximumValue = 100000.0;
constexpr double kFloorF0StoneMask = 40.0;
constexpr double kFrequencyInterval = 3000.0;
constexpr double kUpperLimit = 15000.0;
constexpr double kThreshold = 0.85;
constexpr double kFloorF0D4C = 47.0;

// twiddle table: tw[k] = exp(-2 pi i k / kTwN), k in [0, kTwN/2]
constexpr int kTwLog2 = 15;
constexpr int kTwN = 1 << kTwLog2;

// ---- error channel (the WORLD API is void) ----------------------------------
void set_error(const char* fmt, ...);
const char* last_error();
bool check_cuda(cudaError_t e, const char* what, const char* file, int line);
#define WB_CUDA(x) ::wb::check_cuda((x), #x, __FILE__, __LINE__)
#define WB_CUDA_OR_RETURN(x, ret) do { if (!WB_CUDA(x)) return ret; } while (0)

extern unsigned long long g_launch_count;   // kernels launched by this library
#define WB_LAUNCH_CHECK() do { ++::wb::g_launch_count; ::wb::check_cuda(cudaGetLastError(), "kernel launch", __FILE__, __LINE__); } while (0)

// ---- device context ---------------------------------------------------------------------
struct Context {
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;    // device-to-host copies that overlap later stages (lazily created)
  cudaStream_t upload_stream = nullptr;  // host-to-device copies of the NEXT batch while this one computes
  cudaEvent_t copy_event = nullptr;
  double2* d_twiddle = nullptr;          // [kTwN/2 + 1]
  float2* d_twiddle_f = nullptr;         // the same table rounded to FP32
  // compact copies of the same values for one transform size each: table L holds
  // exp(-2 pi i k / 2^L), k = 0 .. 2^(L-1), contiguous, so the entries one FFT touches share
  // cache lines (the master table spreads them 2^(15-L) entries apart and they fall out of the
  // small L1 that is left beside 220 KB of shared memory).  L = 4 .. kTwLog2.
  double2* d_twiddle_c = nullptr;
  float2* d_twiddle_cf = nullptr;
  __host__ __device__ static size_t tw_c_offset(int L) { return ((size_t)1 << (L - 1)) + 2 * (size_t)L; }   // entries before table L
  const double2* tw_c(int L) const { return d_twiddle_c + tw_c_offset(L); }
  const float2* tw_cf(int L) const { return d_twiddle_cf + tw_c_offset(L); }
  uint32_t* d_randn = nullptr;           // randn table: variate k = d_randn[k] / 2^28 - 6
  size_t randn_count = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
};
Context* ctx();                           // lazily initialised; nullptr on failure
// Every exported entry point that touches the device holds one of these for its whole duration: the
// library has ONE context (stream, scratch pool, tables), so concurrent callers are serialised -- the
// reference is not reentrant at all (global RNG state); this makes the replacement safe to call from any
// number of threads.  Recursive: the drop-in entry points call the batched ones.  It also
// binds the calling thread to the context's device (CUDA's current device is per thread and defaults to 0).
struct ApiGuard {
  ApiGuard();
  ~ApiGuard();
  ApiGuard(const ApiGuard&) = delete;
  ApiGuard& operator=(const ApiGuard&) = delete;
};
cudaMemPool_t scratch_pool();             // the library's own stream-ordered pool (DevBuf)
bool trim_pool();                         // give the pool's cached memory back to the driver
bool ensure_randn(size_t count);          // grow the randn table to >= count variates
// Small device -> host read-back on the library stream, complete on return (per-utterance totals,
// counters, statistics partials: the values the host needs to size the next launch).  bytes is a
// multiple of 4.  A cudaMemcpyAsync here would queue on the device-to-host copy engine BEHIND any
// bulk result copy another stream has in flight (hundreds of MB of features / waveform in a
// pipelined run) and stall the compute stream for that copy's whole PCIe time; instead a small
// kernel stores the words into mapped pinned host memory, which involves no copy engine.
bool read_back(void* h_dst, const void* d_src, size_t bytes);
// Small host -> device copy on the library stream without a copy engine (mapped pinned ring + a copy kernel):
// the tables a stage sends ahead of its kernels must not queue behind a bulk upload on the upload stream.
// The source may be reused on return.  d_dst 4-byte aligned.
bool write_dev(void* d_dst, const void* h_src, size_t bytes);
// cudaMemsetAsync on the library stream as a kernel (never a copy engine, see wb_context.cu); 4-byte granularity.
bool dev_fill(void* d_ptr, int byte_value, size_t bytes);
// Issues the bulk copies a caller queued under wb200_set_copy_deferral (wb_api.cu): a stage calls it right before
// a long kernel that needs no host interaction, so the copies run under it instead of beside a read-back.
void flush_deferred_copies();

// Optional per-kernel device timing (CUDA events on the library stream around one launch).
// Off by default; bench.py switches it on to measure the dominant kernel live.
struct KernelTimer {
  explicit KernelTimer(const char* name);
  ~KernelTimer() { stop(); }
  void stop();
  const char* name_;
  cudaEvent_t e0_ = nullptr, e1_ = nullptr;
};
int option(const char* name);           // run-time switch (wb_context.cu), 0 / 1
double measure_fma_peak(bool fp64);       // TFLOP/s of the CUDA-core FMA pipe, measured
void set_stream(cudaStream_t s);          // run everything on a caller-owned stream
void kernel_timing_enable(bool on);
bool kernel_time_query(const char* name, double* ms_total, long long* launches);
void kernel_times_reset();

// Stream-ordered device buffer.  Memory comes from a CUDA memory pool of the library's own
// (cudaMallocFromPoolAsync on the library stream; release threshold "never"), so the per-stage scratch
// buffers -- gigabytes for a corpus-sized batch -- are recycled between stages and calls instead of going
// through cudaMalloc / cudaFree every time, and the device's default pool (PyTorch, other libraries)
// is left alone.  wb200_trim() hands the cached memory back to the driver.
cudaStream_t pool_stream();
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  bool alloc(size_t count) {
    if (count <= n && p) return true;
    release();
    if (count == 0) count = 1;
#ifdef WB_HOST_EMU      // tests/emu: "device" memory is host memory
    p = static_cast<T*>(calloc(count, sizeof(T)));
    n = p ? count : 0;
    return p != nullptr;
  }
  void release() { free(p); p = nullptr; n = 0; }
#else
    // Sizes of 1 MiB and more are rounded up to four significant bits (< 12.5 % more): the data-dependent
    // scratch sizes of consecutive batches then repeat, and a freed block of the pool fits the next request
    // instead of the pool growing by a slightly larger block every call.
    size_t bytes = count * sizeof(T);
    if (bytes >= ((size_t)1 << 20)) {
      int msb = 63;
      while (!((bytes >> msb) & 1)) --msb;
      const size_t q = (size_t)1 << (msb - 3);
      bytes = (bytes + q - 1) & ~(q - 1);
    }
    if (!WB_CUDA(cudaMallocFromPoolAsync((void**)&p, bytes, scratch_pool(), pool_stream()))) { p = nullptr; n = 0; return false; }
    n = bytes / sizeof(T);
    return true;
  }
  void release() { if (p) cudaFreeAsync(p, pool_stream()); p = nullptr; n = 0; }
#endif
  ~DevBuf() { release(); }
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

#if defined(__CUDACC__) || defined(WB_HOST_EMU)
// ---- small device helpers ----------------------------------------------------------------
// matlab_round (W/src/matlabfunctions.cpp:212-214): half away from zero via truncation.
__host__ __device__ __forceinline__ int matlab_round(double x) {
  return x > 0 ? static_cast<int>(x + 0.5) : static_cast<int>(x - 0.5);
}

// Non-contracted arithmetic for index-determining expressions (the reference is compiled
// without FMA contraction; an fma() here could move a rounding across an integer boundary).
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

// W/src/matlabfunctions.cpp:276: (double)v / 268435456.0 - 6.0.  Both operations are exact in
// double (v < 2^32), so is this form: the integer is placed in the mantissa of 2^52 + v and one
// FMA removes the offset, scales by 2^-28 and subtracts 6 -- no 64-bit integer conversion.
__device__ __forceinline__ double randn_from_u32(uint32_t v) {
  const double biased = __hiloint2double(0x43300000, static_cast<int>(v));      // 2^52 + v
  return fma(biased, 3.7252902984619140625e-09, -(16777216.0 + 6.0));          // 2^-28, 2^52 * 2^-28 = 2^24
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of up to 3 values; every thread gets the totals. `red` = >= 3*32 doubles
// of shared scratch.  Contains two __syncthreads().
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();          // protect `red` from a previous use
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[i * 32 + wid] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double t = lane < nw ? red[i * 32 + lane] : 0.0;
    v[i] = warp_sum(t);
  }
}

// In-place inclusive prefix sum of a[0..n) in shared memory (FP64), blockDim.x threads.
// Each thread scans a contiguous chunk, chunk totals are scanned with warp shuffles.
// `red` = >= 33 doubles of shared scratch.  Ends with a __syncthreads().
__device__ __forceinline__ void block_inclusive_scan(double* a, int n, double* red) {
  const int T = blockDim.x, tid = threadIdx.x;
  const int per = (n + T - 1) / T;
  const int lo = min(n, tid * per), hi = min(n, lo + per);
  double s = 0.0;
  for (int i = lo; i < hi; ++i) { s += a[i]; a[i] = s; }
  // exclusive scan of s across threads
  const int lane = tid & 31, wid = tid >> 5, nw = (T + 31) >> 5;
  double inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) red[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    double w = lane < nw ? red[lane] : 0.0;
    double winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    red[lane] = winc - w;    // exclusive warp offsets
  }
  __syncthreads();
  const double offset = red[wid] + (inc - s);
  if (offset != 0.0)
    for (int i = lo; i < hi; ++i) a[i] += offset;
  __syncthreads();
}

// ---- TMA bulk copy (cp.async.bulk; UBLKCP in SASS) ------------------------------------------------
// A frame's sample window is a contiguous range of the utterance: one elected thread asks the
// TMA unit to copy it into shared memory and arms an mbarrier with the byte count; the other
// threads meanwhile compute window coefficients and wait on the barrier only when they need the
// samples.  Source and destination must be 16-byte aligned and the size a multiple of 16 bytes.
#ifdef WB_HOST_EMU
// CPU emulation: the copy is done by the issuing thread, the barrier word counts completed phases
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned) { __atomic_store_n(bar, 0ull, __ATOMIC_SEQ_CST); }
__device__ __forceinline__ void bulk_load_issue(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
  memcpy(smem_dst, gsrc, bytes);
  __atomic_fetch_add(bar, 1ull, __ATOMIC_SEQ_CST);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  while ((__atomic_load_n(bar, __ATOMIC_SEQ_CST) & 1ull) == parity) ::wbemu::yield();
}
#else
__device__ __forceinline__ unsigned smem_addr(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load_issue(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
  // order earlier generic-proxy accesses of the destination before the async-proxy write
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
#endif  // WB_HOST_EMU
// Window [g0, g0 + W) of an utterance of x_len samples whose base pointer is 16-byte aligned:
// can it be staged with one bulk copy into `cap` doubles?  If so returns true and the aligned
// range [*a0, *a0 + *n); sample g0 + i then sits at staging[(g0 - *a0) + i].
__device__ __forceinline__ bool bulk_window_range(int g0, int W, int x_len, int cap, int* a0, int* n) {
  if (g0 < 0 || g0 + W > x_len) return false;      // the reference clamps indices at the edges: gather path
  *a0 = g0 & ~1;
  *n = ((g0 + W + 1) & ~1) - *a0;                 // may include the (allocated) pad sample after x_len
  return *n <= cap;
}

// interp1Q (W/src/matlabfunctions.cpp:220-241) for one query: uniform grid starting at x0
// with step dx, n samples in y; delta_y[n-1] is defined as 0 by the reference.
// The base index is found by multiplying with 1 / dx instead of dividing: when that moves the
// truncation across an integer the fraction is ~0 or ~1 and the interpolant is continuous, so
// the result changes by rounding noise only.
__device__ __forceinline__ double interp1q_at(double x0, double inv_dx, const double* y, int n,
                                              double xi) {
  const double r = add_rn(xi, -x0) * inv_dx;
  int base = static_cast<int>(r);
  const double frac = r - base;
  base = max(0, min(n - 1, base));                 // memory guard only (UB in the reference)
  const double y0 = y[base];
  const double dy = base + 1 < n ? y[base + 1] - y0 : 0.0;
  return add_rn(y0, mul_rn(dy, frac));
}
#endif  // __CUDACC__

}  // namespace wb
