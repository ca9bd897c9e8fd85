// world-b200: device context, error channel, twiddle table, randn table, batch layout.
#include <math.h>
#include <stdarg.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "wb_batch.h"

namespace wb {

// ---------------------------------------------------------------------------------------------
// error channel: the WORLD API returns void, so failures are recorded here and
// the caller's outputs are NaN-filled by the API layer.
// ---------------------------------------------------------------------------------------------
static std::mutex g_err_mutex;
static std::string g_err;
unsigned long long g_launch_count = 0;
StageTimes g_times;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  std::lock_guard<std::mutex> lock(g_err_mutex);
  g_err = buf;
  fprintf(stderr, "[world_b200] error: %s\n", buf);
}
const char* last_error() {
  static thread_local std::string snapshot;   // the caller's pointer stays valid if another thread reports meanwhile
  std::lock_guard<std::mutex> lock(g_err_mutex);
  snapshot = g_err;
  return snapshot.c_str();
}
bool check_cuda(cudaError_t e, const char* what, const char* file, int line) {
  if (e == cudaSuccess) return true;
  set_error("%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
  return false;
}

// ---------------------------------------------------------------------------------------------
// randn table.  The reference's randn() (W/src/matlabfunctions.cpp:247-277) is a xorshift128
// stream re-seeded at the top of CheapTrick / D4C / Synthesis, so variate k of a call is a
// data-independent constant.  We materialise the stream once in HBM (uint32 sum of the twelve
// (w >> 4) terms; value = sum / 2^28 - 6) and every kernel indexes it.  The table is generated
// on the GPU: the host only computes, by GF(2) matrix powers of the xorshift step, the
// generator state at the start of every chunk of kRandnChunk variates.
// ---------------------------------------------------------------------------------------------
struct XState { uint32_t x, y, z, w; };
constexpr int kRandnChunk = 1024;

__host__ __device__ inline void xorshift_step(XState& s) {
  const uint32_t t = s.x ^ (s.x << 11);
  s.x = s.y; s.y = s.z; s.z = s.w;
  s.w = (s.w ^ (s.w >> 19)) ^ (t ^ (t >> 8));
}

struct GF2Mat { XState col[128]; };   // column i = image of basis vector i

static XState gf2_apply(const GF2Mat& m, const XState& s) {
  XState r = {0, 0, 0, 0};
  const uint32_t w[4] = {s.x, s.y, s.z, s.w};
  for (int i = 0; i < 128; ++i)
    if ((w[i >> 5] >> (i & 31)) & 1u) {
      r.x ^= m.col[i].x; r.y ^= m.col[i].y; r.z ^= m.col[i].z; r.w ^= m.col[i].w;
    }
  return r;
}
static void gf2_mul(const GF2Mat& a, const GF2Mat& b, GF2Mat* out) {   // out = a * b
  GF2Mat r;
  for (int i = 0; i < 128; ++i) r.col[i] = gf2_apply(a, b.col[i]);
  *out = r;
}
static void gf2_step_matrix(GF2Mat* m) {
  for (int i = 0; i < 128; ++i) {
    uint32_t w[4] = {0, 0, 0, 0};
    w[i >> 5] = 1u << (i & 31);
    XState s = {w[0], w[1], w[2], w[3]};
    xorshift_step(s);
    m->col[i] = s;
  }
}
static void gf2_pow(const GF2Mat& base, unsigned long long e, GF2Mat* out) {
  GF2Mat result, b = base;
  for (int i = 0; i < 128; ++i) {          // identity
    uint32_t w[4] = {0, 0, 0, 0};
    w[i >> 5] = 1u << (i & 31);
    result.col[i] = XState{w[0], w[1], w[2], w[3]};
  }
  while (e) {
    if (e & 1ull) gf2_mul(b, result, &result);
    gf2_mul(b, b, &b);
    e >>= 1;
  }
  *out = result;
}

__global__ void randn_table_kernel(const XState* __restrict__ starts, uint32_t* __restrict__ out,
                                   size_t n_chunks) {
  const size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  XState s = starts[c];
  uint32_t* o = out + c * kRandnChunk;
  for (int k = 0; k < kRandnChunk; ++k) {
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < 12; ++j) { xorshift_step(s); acc += s.w >> 4; }
    o[k] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// per-kernel timing: events are queued at launch and resolved when queried
// ---------------------------------------------------------------------------------------------
struct PendingTime { std::string name; cudaEvent_t e0, e1; };
struct TimeAcc { double ms = 0.0; long long n = 0; };
static std::mutex g_kt_mutex;
static bool g_kt_on = false;
static std::vector<PendingTime> g_kt_pending;
static std::map<std::string, TimeAcc> g_kt_acc;

void kernel_timing_enable(bool on) { std::lock_guard<std::mutex> l(g_kt_mutex); g_kt_on = on; }
KernelTimer::KernelTimer(const char* name) : name_(name) {
  if (!g_kt_on) return;
  Context* c = ctx();
  if (!c) return;
  if (cudaEventCreate(&e0_) != cudaSuccess || cudaEventCreate(&e1_) != cudaSuccess) { e0_ = e1_ = nullptr; return; }
  cudaEventRecord(e0_, c->stream);
}
void KernelTimer::stop() {
  if (!e0_ || !e1_) return;
  cudaEventRecord(e1_, ctx()->stream);
  std::lock_guard<std::mutex> l(g_kt_mutex);
  g_kt_pending.push_back(PendingTime{name_, e0_, e1_});
  e0_ = e1_ = nullptr;
}
static void kernel_times_resolve() {
  for (auto& p : g_kt_pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(p.e1) == cudaSuccess && cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
      TimeAcc& a = g_kt_acc[p.name];
      a.ms += ms; a.n += 1;
    }
    cudaEventDestroy(p.e0); cudaEventDestroy(p.e1);
  }
  g_kt_pending.clear();
}
bool kernel_time_query(const char* name, double* ms_total, long long* launches) {
  std::lock_guard<std::mutex> l(g_kt_mutex);
  kernel_times_resolve();
  auto it = g_kt_acc.find(name);
  if (it == g_kt_acc.end()) { *ms_total = 0.0; *launches = 0; return false; }
  *ms_total = it->second.ms; *launches = it->second.n;
  return true;
}
void kernel_times_reset() {
  std::lock_guard<std::mutex> l(g_kt_mutex);
  kernel_times_resolve();
  g_kt_acc.clear();
}

// ---------------------------------------------------------------------------------------------
// CUDA-core peak micro-benchmark: dependent-free FMA chains, 8
// accumulators per thread, 8 CTAs of 256 threads per SM.  Returns TFLOP/s (2 flops per FMA).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters, T a, T b) {
  T v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = static_cast<T>(threadIdx.x + j);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = v[j] * a + b;
  }
  T s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
  if (s == static_cast<T>(-1.2345)) out[0] = s;     // never true; keeps the chain alive
}

template <typename T>
static double fma_peak(Context* c) {
  T* d = nullptr;
  if (cudaMalloc((void**)&d, sizeof(T)) != cudaSuccess) return 0.0;
  const int blocks = c->sm_count * 8, iters = 1 << 14;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, c->stream);
    fma_peak_kernel<T><<<blocks, 256, 0, c->stream>>>(d, iters, static_cast<T>(0.999999), static_cast<T>(1e-7));
    cudaEventRecord(e1, c->stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 32.0 * (double)iters * 256.0 * blocks;
    if (rep > 0 && ms > 0.f) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  return best;
}
// Run-time switches (read once from the environment; `1` / `0` override the default).  They select
// between kernels that compute the same result -- A/B measurements, and bench.py's roofline needs to
// know which precision ran.
int option(const char* name) {
  struct Opt { const char* name; const char* env; int def; int val; };
  static Opt opts[] = {
      {"lovetrain_fp32", "WB_D4C_LT32", 1, -1},        // LoveTrain's transform in FP32
      {"synth_phase4", "WB_SYNTH_PHASE4", 1, -1},      // Synthesis time base: four samples per thread (0: one sample per thread, 1 024-thread CTAs)
      {"dio_fused", "WB_DIO_FUSED", 1, -1},            // Dio: zero crossings inside the filter kernel (0: band signals through HBM)
      {"harvest_fused", "WB_HARVEST_FUSED", 1, -1},
      {"harvest_fix_warp", "WB_HARVEST_FIX_WARP", 1, -1},     // Harvest contour logic spread over a warp (0: lane 0 walks)
      {"harvest_refine_thread", "WB_HARVEST_REFINE_THREAD", 1, -1},   // Harvest refinement: one thread per candidate (0: one warp)    // Harvest: the same for its 152 band-pass channels
      {"stonemask_dft", "WB_STONEMASK_DFT", 1, -1},    // StoneMask: direct evaluation of the <= 8 bins (0: packed FP32 FFT)
  };
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  for (Opt& o : opts) {
    if (strcmp(o.name, name) != 0) continue;
    if (o.val < 0) {
      const char* e = getenv(o.env);
      o.val = e ? (e[0] != '0' && e[0] != '\0') : o.def;
    }
    return o.val;
  }
  return 0;
}

double measure_fma_peak(bool fp64) {
  Context* c = ctx();
  if (!c) return 0.0;
  return fp64 ? fma_peak<double>(c) : fma_peak<float>(c);
}

static std::mutex g_ctx_mutex;
static Context g_ctx;
static bool g_ctx_ok = false, g_ctx_failed = false;
static cudaMemPool_t g_pool = nullptr;
static std::recursive_mutex g_api_mutex;

ApiGuard::ApiGuard() {
  g_api_mutex.lock();
  if (g_ctx_ok) {                          // the context lives on one device; make it this thread's current one
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != g_ctx.device) cudaSetDevice(g_ctx.device);
  }
}
ApiGuard::~ApiGuard() { g_api_mutex.unlock(); }

cudaMemPool_t scratch_pool() { return g_pool; }
bool trim_pool() {
  ApiGuard guard;
  if (!g_ctx_ok || !g_pool) return true;
  return WB_CUDA(cudaStreamSynchronize(g_ctx.stream)) && WB_CUDA(cudaMemPoolTrimTo(g_pool, 0));
}

cudaStream_t pool_stream() { return g_ctx_ok ? g_ctx.stream : nullptr; }

void set_stream(cudaStream_t s) {
  Context* c = ctx();
  if (c) { cudaStreamSynchronize(c->stream); c->stream = s; }
}

Context* ctx() {
  std::lock_guard<std::mutex> lock(g_ctx_mutex);
  if (g_ctx_ok) return &g_ctx;
  if (g_ctx_failed) return nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: world_b200 has no CPU fallback");
    g_ctx_failed = true;
    return nullptr;
  }
  int dev = 0;
  if (!WB_CUDA(cudaGetDevice(&dev))) { g_ctx_failed = true; return nullptr; }
  g_ctx.device = dev;
  cudaDeviceProp prop;
  if (!WB_CUDA(cudaGetDeviceProperties(&prop, dev))) { g_ctx_failed = true; return nullptr; }
  g_ctx.sm_count = prop.multiProcessorCount;
  g_ctx.smem_optin = prop.sharedMemPerBlockOptin;
  if (!WB_CUDA(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking))) {
    g_ctx_failed = true; return nullptr;
  }
  {   // the library's own pool: freed scratch stays in it instead of returning to the driver at every sync
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    unsigned long long never = ~0ull;
    if (!WB_CUDA(cudaMemPoolCreate(&g_pool, &props)) ||
        !WB_CUDA(cudaMemPoolSetAttribute(g_pool, cudaMemPoolAttrReleaseThreshold, &never))) {
      g_ctx_failed = true; return nullptr;
    }
  }
  // twiddles, rounded from long double
  std::vector<double2> tw(kTwN / 2 + 1);
  for (int k = 0; k <= kTwN / 2; ++k) {
    const long double a = -2.0L * 3.14159265358979323846264338327950288L * k / kTwN;
    tw[k] = make_double2((double)cosl(a), (double)sinl(a));
  }
  std::vector<float2> twf(tw.size());
  for (size_t k = 0; k < tw.size(); ++k) twf[k] = make_float2((float)tw[k].x, (float)tw[k].y);
  if (!WB_CUDA(cudaMalloc((void**)&g_ctx.d_twiddle, tw.size() * sizeof(double2))) ||
      !WB_CUDA(cudaMemcpy(g_ctx.d_twiddle, tw.data(), tw.size() * sizeof(double2),
                          cudaMemcpyHostToDevice)) ||
      !WB_CUDA(cudaMalloc((void**)&g_ctx.d_twiddle_f, twf.size() * sizeof(float2))) ||
      !WB_CUDA(cudaMemcpy(g_ctx.d_twiddle_f, twf.data(), twf.size() * sizeof(float2),
                          cudaMemcpyHostToDevice))) {
    g_ctx_failed = true; return nullptr;
  }
  {   // compact per-size tables: table L = every 2^(kTwLog2-L)-th entry of the master table
    std::vector<double2> twc(Context::tw_c_offset(kTwLog2 + 1));
    std::vector<float2> twcf(twc.size());
    for (int L = 4; L <= kTwLog2; ++L) {
      const size_t off = Context::tw_c_offset(L);
      for (int k = 0; k <= (1 << (L - 1)); ++k) {
        twc[off + k] = tw[(size_t)k << (kTwLog2 - L)];
        twcf[off + k] = twf[(size_t)k << (kTwLog2 - L)];
      }
    }
    if (!WB_CUDA(cudaMalloc((void**)&g_ctx.d_twiddle_c, twc.size() * sizeof(double2))) ||
        !WB_CUDA(cudaMemcpy(g_ctx.d_twiddle_c, twc.data(), twc.size() * sizeof(double2), cudaMemcpyHostToDevice)) ||
        !WB_CUDA(cudaMalloc((void**)&g_ctx.d_twiddle_cf, twcf.size() * sizeof(float2))) ||
        !WB_CUDA(cudaMemcpy(g_ctx.d_twiddle_cf, twcf.data(), twcf.size() * sizeof(float2), cudaMemcpyHostToDevice))) {
      g_ctx_failed = true; return nullptr;
    }
  }
  g_ctx_ok = true;
  return &g_ctx;
}

bool ensure_randn(size_t count) {
  ApiGuard guard;                                      // the table swap below must not race with another caller
  Context* c = ctx();
  if (!c) return false;
  if (count <= c->randn_count) return true;
  size_t n_chunks = (count + kRandnChunk - 1) / kRandnChunk;
  const size_t min_chunks = (size_t)4096;               // 4 Mi variates at least
  if (n_chunks < min_chunks) n_chunks = min_chunks;
  n_chunks = (n_chunks * 5 / 4 + 1023) / 1024 * 1024;  // head-room, fewer regenerations
  static GF2Mat jump;
  static bool have_jump = false;
  if (!have_jump) {
    GF2Mat step;
    gf2_step_matrix(&step);
    gf2_pow(step, 12ull * kRandnChunk, &jump);
    have_jump = true;
  }
  std::vector<XState> starts(n_chunks);
  XState s = {123456789u, 362436069u, 521288629u, 88675123u};   // randn_reseed()
  for (size_t i = 0; i < n_chunks; ++i) { starts[i] = s; s = gf2_apply(jump, s); }
  XState* d_starts = nullptr;
  uint32_t* d_tab = nullptr;
  if (!WB_CUDA(cudaMalloc((void**)&d_starts, n_chunks * sizeof(XState)))) return false;
  if (!WB_CUDA(cudaMalloc((void**)&d_tab, n_chunks * kRandnChunk * sizeof(uint32_t)))) {
    cudaFree(d_starts); return false;
  }
  bool ok = WB_CUDA(cudaMemcpyAsync(d_starts, starts.data(), n_chunks * sizeof(XState),
                                    cudaMemcpyHostToDevice, c->stream));
  if (ok) {
    randn_table_kernel<<<(unsigned)((n_chunks + 127) / 128), 128, 0, c->stream>>>(d_starts, d_tab, n_chunks);
    WB_LAUNCH_CHECK();
    ok = WB_CUDA(cudaStreamSynchronize(c->stream));
  }
  cudaFree(d_starts);
  if (!ok) { cudaFree(d_tab); return false; }
  if (c->d_randn) cudaFree(c->d_randn);
  c->d_randn = d_tab;
  c->randn_count = n_chunks * kRandnChunk;
  return true;
}

// ---------------------------------------------------------------------------------------------
// read_back: device words -> mapped pinned host memory by a kernel (see wb_common.cuh)
// ---------------------------------------------------------------------------------------------
__global__ void read_back_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, size_t n_words) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

static std::mutex g_rb_mutex;
static void* g_rb_host = nullptr;          // cudaHostAlloc(Mapped)
static uint32_t* g_rb_dev = nullptr;       // the same memory as the device sees it
static size_t g_rb_cap = 0;

bool read_back(void* h_dst, const void* d_src, size_t bytes) {
  Context* c = ctx();
  if (!c) return false;
  if (bytes & 3) { set_error("read_back: %zu bytes is not a multiple of 4", bytes); return false; }
  std::lock_guard<std::mutex> lock(g_rb_mutex);
  if (bytes > g_rb_cap) {
    size_t cap = (size_t)1 << 16;
    while (cap < bytes) cap <<= 1;
    if (g_rb_host) { cudaFreeHost(g_rb_host); g_rb_host = nullptr; g_rb_dev = nullptr; g_rb_cap = 0; }
    if (!WB_CUDA(cudaHostAlloc(&g_rb_host, cap, cudaHostAllocMapped))) { g_rb_host = nullptr; return false; }
    if (!WB_CUDA(cudaHostGetDevicePointer((void**)&g_rb_dev, g_rb_host, 0))) {
      cudaFreeHost(g_rb_host); g_rb_host = nullptr; g_rb_dev = nullptr;
      return false;
    }
    g_rb_cap = cap;
  }
  if (bytes > 0) {
    const size_t n_words = bytes / 4;
    const unsigned blocks = (unsigned)std::min<size_t>((n_words + 255) / 256, (size_t)c->sm_count);
    read_back_kernel<<<blocks, 256, 0, c->stream>>>(static_cast<const uint32_t*>(d_src), g_rb_dev, n_words);
    WB_LAUNCH_CHECK();
  }
  if (!WB_CUDA(cudaStreamSynchronize(c->stream))) return false;
  if (bytes > 0) memcpy(h_dst, g_rb_host, bytes);
  return true;
}

// cudaMemsetAsync as a kernel.  The driver may execute a memset on a copy engine; when that engine is busy with
// a bulk copy of another stream (the coded features leaving while Synthesis starts by zeroing its 1.6 GB of
// accumulators) the memset -- and the compute stream behind it -- waits for the whole copy: the features cost
// the end-to-end leg their full PCIe time.  ptr and bytes must be multiples of 4.
__global__ void dev_fill_kernel(uint32_t* __restrict__ p, uint32_t v, size_t n_words) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    uint4* p4 = reinterpret_cast<uint4*>(p);
    const size_t n4 = n_words / 4;
    for (size_t k = i; k < n4; k += stride) p4[k] = make_uint4(v, v, v, v);
    for (size_t k = n4 * 4 + i; k < n_words; k += stride) p[k] = v;
  } else {
    for (; i < n_words; i += stride) p[i] = v;
  }
}
bool dev_fill(void* d_ptr, int byte_value, size_t bytes) {
  Context* c = ctx();
  if (!c) return false;
  if (bytes == 0) return true;
  if ((reinterpret_cast<uintptr_t>(d_ptr) & 3) || (bytes & 3)) { set_error("dev_fill: pointer / size not a multiple of 4"); return false; }
  const uint32_t b = (uint32_t)(byte_value & 0xff), v = b | (b << 8) | (b << 16) | (b << 24);
  const size_t n_words = bytes / 4;
  const unsigned blocks = (unsigned)std::max<size_t>(1, std::min<size_t>((n_words / 4 + 255) / 256, (size_t)c->sm_count * 8));
  dev_fill_kernel<<<blocks, 256, 0, c->stream>>>(static_cast<uint32_t*>(d_ptr), v, n_words);
  WB_LAUNCH_CHECK();
  return true;
}

// The opposite direction for the small host -> device tables a stage sends ahead of its kernels (offsets,
// lengths, window coefficients): a cudaMemcpyAsync would queue on the host-to-device copy engine BEHIND the bulk
// upload of the next batch that a pipelined caller has in flight on the upload stream (hundreds of MB: the
// stage that runs beside an upload -- Dio -- lost the upload's whole PCIe time, ~4 ms per sub-batch of the
// end-to-end leg).  The bytes go into a ring of mapped pinned host memory and a small kernel on the library
// stream copies them from there; no copy engine is involved.  Like cudaMemcpyAsync from pageable memory the
// source may be reused as soon as the call returns.  bytes need not be a multiple of 4 (the ring is padded).
static void* g_wd_host = nullptr;
static uint32_t* g_wd_dev = nullptr;
static size_t g_wd_cap = 0, g_wd_off = 0;
__global__ void write_dev_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, size_t n_words, size_t tail_bytes) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_words; i += stride) dst[i] = src[i];
  if (tail_bytes && blockIdx.x == 0 && threadIdx.x < tail_bytes)
    reinterpret_cast<unsigned char*>(dst + n_words)[threadIdx.x] = reinterpret_cast<const unsigned char*>(src + n_words)[threadIdx.x];
}
bool write_dev(void* d_dst, const void* h_src, size_t bytes) {
  Context* c = ctx();
  if (!c) return false;
  if (bytes == 0) return true;
  if ((reinterpret_cast<uintptr_t>(d_dst) & 3) != 0) { set_error("write_dev: destination is not 4-byte aligned"); return false; }
  std::lock_guard<std::mutex> lock(g_rb_mutex);
  const size_t need = (bytes + 15) & ~(size_t)15;
  if (need > g_wd_cap) {
    size_t cap = (size_t)16 << 20;
    while (cap < 2 * need) cap <<= 1;
    if (g_wd_host) { cudaStreamSynchronize(c->stream); cudaFreeHost(g_wd_host); g_wd_host = nullptr; g_wd_dev = nullptr; g_wd_cap = 0; }
    if (!WB_CUDA(cudaHostAlloc(&g_wd_host, cap, cudaHostAllocMapped))) { g_wd_host = nullptr; return false; }
    if (!WB_CUDA(cudaHostGetDevicePointer((void**)&g_wd_dev, g_wd_host, 0))) {
      cudaFreeHost(g_wd_host); g_wd_host = nullptr; g_wd_dev = nullptr;
      return false;
    }
    g_wd_cap = cap; g_wd_off = 0;
  }
  if (g_wd_off + need > g_wd_cap) {                  // the ring wraps: everything queued so far must have read its slice
    if (!WB_CUDA(cudaStreamSynchronize(c->stream))) return false;
    g_wd_off = 0;
  }
  memcpy(static_cast<char*>(g_wd_host) + g_wd_off, h_src, bytes);
  const size_t n_words = bytes / 4;
  const unsigned blocks = (unsigned)std::max<size_t>(1, std::min<size_t>((n_words + 255) / 256, (size_t)c->sm_count));
  write_dev_kernel<<<blocks, 256, 0, c->stream>>>(g_wd_dev + g_wd_off / 4, static_cast<uint32_t*>(d_dst), n_words, bytes & 3);
  WB_LAUNCH_CHECK();
  g_wd_off += need;
  return true;
}

// ---------------------------------------------------------------------------------------------
// segmented exclusive scan over the frames of each utterance (one CTA per utterance)
// ---------------------------------------------------------------------------------------------
__global__ void seg_scan_kernel(const long long* __restrict__ counts, const int* __restrict__ f_off,
                                const int* __restrict__ f_len, long long* __restrict__ out,
                                long long* __restrict__ totals) {
  __shared__ long long wsum[32];
  __shared__ long long carry_s;
  const int u = blockIdx.x;
  const int off = f_off[u], len = f_len[u];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < len; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const long long v = i < len ? counts[off + i] : 0;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      const long long w = lane < nw ? wsum[lane] : 0;
      long long winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      wsum[lane] = winc - w;
      if (lane == 31) wsum[31] = winc - w;   // keep exclusive; total handled below
    }
    __syncthreads();
    const long long carry = carry_s;
    const long long excl = carry + wsum[wid] + (inc - v);
    if (i < len) out[off + i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0 && totals) totals[u] = carry_s;
}

bool segmented_exclusive_scan(const long long* counts, const int* f_off, const int* f_len,
                              int n_utt, long long* out, long long* totals) {
  Context* c = ctx();
  if (!c) return false;
  if (n_utt <= 0) return true;
  seg_scan_kernel<<<n_utt, 256, 0, c->stream>>>(counts, f_off, f_len, out, totals);
  WB_LAUNCH_CHECK();
  return true;
}

// ---------------------------------------------------------------------------------------------
// batch layout
// ---------------------------------------------------------------------------------------------
__global__ void default_frames_kernel(const int* __restrict__ f_off, const int* __restrict__ f_len,
                                      double frame_period, int* __restrict__ frame_utt,
                                      double* __restrict__ frame_t) {
  const int u = blockIdx.x;
  const int off = f_off[u], len = f_len[u];
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    frame_utt[off + i] = u;
    frame_t[off + i] = __ddiv_rn(__dmul_rn((double)i, frame_period), 1000.0);  // i * fp / 1000.0
  }
}

bool batch_layout(Batch* b, int fs, double frame_period, int n_utt, const int* x_len,
                  const int* f_len) {
  Context* c = ctx();
  if (!c) return false;
  b->fs = fs; b->frame_period = frame_period; b->n_utt = n_utt;
  b->h_x_off.resize(n_utt); b->h_x_len.assign(x_len, x_len + n_utt);
  b->h_f_off.resize(n_utt); b->h_f_len.assign(f_len, f_len + n_utt);
  long long so = 0; long long fo = 0;
  b->max_x_len = 0; b->max_f_len = 0;
  for (int u = 0; u < n_utt; ++u) {
    b->h_x_off[u] = so; so += (x_len[u] + 1) & ~1LL;    // keep every utterance 16-byte aligned
    b->h_f_off[u] = (int)fo; fo += f_len[u];
    if (x_len[u] > b->max_x_len) b->max_x_len = x_len[u];
    if (f_len[u] > b->max_f_len) b->max_f_len = f_len[u];
  }
  if (fo > 0x7fffffffLL) { set_error("batch has too many frames (%lld)", fo); return false; }
  b->total_samples = so; b->total_frames = (int)fo;
  if (!b->x.alloc((size_t)so) || !b->x_off.alloc(n_utt) || !b->x_len.alloc(n_utt) ||
      !b->f_off.alloc(n_utt) || !b->f_len.alloc(n_utt) || !b->frame_utt.alloc(fo) ||
      !b->frame_t.alloc(fo) || !b->f0_raw.alloc(fo) || !b->f0.alloc(fo))
    return false;
  cudaStream_t st = c->stream;
  bool ok = true;
  if (n_utt > 0) {
    ok = ok && WB_CUDA(cudaMemcpyAsync(b->x_off.p, b->h_x_off.data(), n_utt * sizeof(long long), cudaMemcpyHostToDevice, st));
    ok = ok && WB_CUDA(cudaMemcpyAsync(b->x_len.p, b->h_x_len.data(), n_utt * sizeof(int), cudaMemcpyHostToDevice, st));
    ok = ok && WB_CUDA(cudaMemcpyAsync(b->f_off.p, b->h_f_off.data(), n_utt * sizeof(int), cudaMemcpyHostToDevice, st));
    ok = ok && WB_CUDA(cudaMemcpyAsync(b->f_len.p, b->h_f_len.data(), n_utt * sizeof(int), cudaMemcpyHostToDevice, st));
    ok = ok && WB_CUDA(cudaStreamSynchronize(st));     // the host vectors may be reallocated later
  }
  return ok;
}

bool batch_default_frames(Batch* b) {
  Context* c = ctx();
  if (!c) return false;
  if (b->n_utt <= 0) return true;
  default_frames_kernel<<<b->n_utt, 256, 0, c->stream>>>(b->f_off.p, b->f_len.p, b->frame_period,
                                                        b->frame_utt.p, b->frame_t.p);
  WB_LAUNCH_CHECK();
  return true;
}

}  // namespace wb
