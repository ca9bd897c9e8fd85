// TEST INFRASTRUCTURE: a minimal CUDA-on-CPU shim, enough to run the D4C kernels of
// hts-train-world_b200/csrc/wb_d4c.cu one CTA at a time (every CUDA thread is an OS thread,
// __syncthreads() a barrier, warp shuffles an exchange through a per-warp slot array).  It finds
// index / layout / logic errors of a kernel before it has seen a GPU; it says nothing about
// races, bank conflicts or speed.  g++ -std=c++20 -DWB_HOST_EMU.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <barrier>
#include <memory>
#include <thread>
#include <vector>

namespace wbemu {
struct Dim { unsigned x = 1, y = 1, z = 1; };
inline thread_local Dim t_idx;
inline Dim b_idx, b_dim, g_dim;
alignas(16) inline unsigned char dyn_smem[256 * 1024];
inline std::unique_ptr<std::barrier<>> block_bar;
inline std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
inline uint64_t xchg[64][32];
inline void yield() { std::this_thread::yield(); }
// WBEMU_SKIP_BARRIER=n drops the n-th __syncthreads() of every CTA (all threads skip the same one):
// the self-test of the race check (tests/emu/README.md) -- ThreadSanitizer must then report a race.
inline thread_local int barrier_no = 0;
inline int skip_barrier = getenv("WBEMU_SKIP_BARRIER") ? atoi(getenv("WBEMU_SKIP_BARRIER")) : -1;
inline void syncthreads() {
  if (barrier_no++ == skip_barrier) return;
  block_bar->arrive_and_wait();
}

template <typename T, typename SRC>
inline T exchange(T v, SRC src_of_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of <= 8 bytes");
  const int w = t_idx.x >> 5, lane = t_idx.x & 31;
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  xchg[w][lane] = raw;
  warp_bar[w]->arrive_and_wait();
  const int src = src_of_lane(lane);
  T r = v;
  if (src >= 0 && src < 32) memcpy(&r, &xchg[w][src], sizeof(T));
  warp_bar[w]->arrive_and_wait();
  return r;
}
template <typename T> inline T shfl_xor(T v, int o) { return exchange(v, [o](int l) { return l ^ o; }); }
template <typename T> inline T shfl_up(T v, int o) { return exchange(v, [o](int l) { return l - o; }); }
template <typename T> inline T shfl_down(T v, int o) { return exchange(v, [o](int l) { return l + o; }); }
template <typename T> inline T shfl_idx(T v, int src) { return exchange(v, [src](int) { return src; }); }
template <typename T, typename OP> inline T reduce(T v, OP op) {
  for (int o = 16; o > 0; o >>= 1) v = op(v, shfl_xor(v, o));
  return v;
}

// run `body()` as a grid of CTAs of `threads` threads, one CTA at a time.  smem_bytes = the dynamic
// shared memory the launcher would request: the 4 KB behind it hold a canary, and a kernel that
// writes past its allocation is counted in smem_overruns.
inline int smem_overruns = 0;
template <typename F>
inline void launch(const std::vector<int>& blocks, int grid, int threads, size_t smem_bytes, F body) {
  b_dim.x = threads;
  g_dim.x = grid;
  block_bar = std::make_unique<std::barrier<>>(threads);
  warp_bar.clear();
  for (int w = 0; w < (threads + 31) / 32; ++w) warp_bar.push_back(std::make_unique<std::barrier<>>(32));
  // the CTA's threads are created once per launch and walk the list of CTAs together (thread 0
  // switches blockIdx and checks the canary between two rendezvous)
  std::barrier<> cta_bar(threads);
  auto worker = [&](int i) {
    t_idx.x = i;
    for (size_t k = 0; k < blocks.size(); ++k) {
      if (i == 0) {
        b_idx.x = blocks[k];
        memset(dyn_smem + smem_bytes, 0xA5, 4096);
      }
      cta_bar.arrive_and_wait();
      barrier_no = 0;
      body();
      cta_bar.arrive_and_wait();
      if (i == 0)
        for (int q = 0; q < 4096; ++q)
          if (dyn_smem[smem_bytes + q] != 0xA5) { ++smem_overruns; break; }
    }
  };
  std::vector<std::thread> th;
  for (int i = 0; i < threads; ++i) th.emplace_back(worker, i);
  for (auto& t : th) t.join();
}
// a whole 2-D grid, x fastest
template <typename F>
inline void launch_grid(int gx, int gy, int threads, size_t smem_bytes, F body) {
  g_dim.y = gy;
  for (int y = 0; y < gy; ++y) {
    b_idx.y = y;
    std::vector<int> xs(gx);
    for (int i = 0; i < gx; ++i) xs[i] = i;
    launch(xs, gx, threads, smem_bytes, body);
  }
  b_idx.y = 0;
  g_dim.y = 1;
}
}  // namespace wbemu

#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __noinline__
#define __noinline__
#undef __forceinline__
#define __forceinline__ inline
#define threadIdx (::wbemu::t_idx)
#define blockIdx (::wbemu::b_idx)
#define blockDim (::wbemu::b_dim)
#define gridDim (::wbemu::g_dim)
#define __syncthreads() ::wbemu::syncthreads()
#define __syncwarp() ::wbemu::warp_bar[::wbemu::t_idx.x >> 5]->arrive_and_wait()
#define __shfl_xor_sync(m, v, o) ::wbemu::shfl_xor((v), (o))
#define __shfl_up_sync(m, v, o) ::wbemu::shfl_up((v), (o))
#define __shfl_down_sync(m, v, o) ::wbemu::shfl_down((v), (o))
#define __shfl_sync(m, v, src) ::wbemu::shfl_idx((v), (src))
#define __reduce_add_sync(m, v) ::wbemu::reduce((v), [](auto a, auto b) { return a + b; })
#define __reduce_max_sync(m, v) ::wbemu::reduce((v), [](auto a, auto b) { return a > b ? a : b; })
#define __ldg(p) (*(p))
inline unsigned __brev(unsigned v) {
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
  return r;
}
inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
inline int __double2hiint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(u >> 32); }
inline double __hiloint2double(int hi, int lo) {
  const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double d; memcpy(&d, &u, 8); return d;
}
inline void sincospi(double a, double* s, double* c) {
  const long double x = 3.14159265358979323846264338327950288L * (long double)a;
  *s = (double)sinl(x); *c = (double)cosl(x);
}
inline double cospi(double a) { double s, c; sincospi(a, &s, &c); return c; }
inline void sincospif(float a, float* s, float* c) {
  const double x = 3.14159265358979323846 * (double)a;
  *s = (float)sin(x); *c = (float)cos(x);
}
#define __any_sync(m, pred) (::wbemu::reduce((int)((pred) ? 1 : 0), [](int a, int b) { return a | b; }) != 0)
#define __ballot_sync(m, pred) ::wbemu::reduce((unsigned)((pred) ? (1u << (::wbemu::t_idx.x & 31)) : 0u), [](unsigned a, unsigned b) { return a | b; })
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline void __sincosf(float a, float* s, float* c) { *s = sinf(a); *c = cosf(a); }
inline float __logf(float a) { return logf(a); }
inline float __expf(float a) { return expf(a); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline long long __double2ll_rn(double v) { return llrint(v); }
inline double atomicAdd(double* p, double v) {
  uint64_t old = __atomic_load_n(reinterpret_cast<uint64_t*>(p), __ATOMIC_SEQ_CST), neu;
  double d;
  do {
    memcpy(&d, &old, 8);
    const double sum = d + v;
    memcpy(&neu, &sum, 8);
  } while (!__atomic_compare_exchange_n(reinterpret_cast<uint64_t*>(p), &old, neu, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
  return d;
}
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAnd(int* p, int v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
inline int atomicMax(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
using std::max;
using std::min;
