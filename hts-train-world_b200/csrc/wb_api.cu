// world-b200: the C ABI.  (1) the unchanged WORLD entry points of the reference
// (W/src/world/*.h) as one-utterance batches, (2) the batched extension API of
// include/world_b200.h.  Host code only stages data; all arithmetic runs in the kernels.
#include <math.h>
#include <string.h>
#include <functional>
#include <limits>
#include <string>
#include "../../include/world_b200.h"
#include "wb_batch.h"

using namespace wb;

namespace wb {
bool batch_compose_cmp(Batch* b, const wb200_cmp_stream* streams, int n_streams);   // wb_cmp.cu
bool batch_cmp_stats(Batch* b, double* h_out);
}

struct wb200_batch {
  Batch b;
  DevBuf<int16_t> pcm_stage, pcm_out;
  DevBuf<long long> src_off, out_off;   // per-utterance offsets inside the packed 16-bit input / output
  std::vector<int> out_layout;          // the y_len[] that out_off / pcm_out were built for
  cudaEvent_t upload_done = nullptr;     // recorded on the upload stream by wb200_batch_upload_pcm16_async
  bool upload_pending = false;
  // recorded on the download stream after the asynchronous result copies of this batch: the next
  // pass over the same batch object must not overwrite lf0 / mgc / bap / the 16-bit staging before
  // those copies have read them (a caller that pipelines batches back to back never calls wb200_sync)
  cudaEvent_t download_done[2] = {nullptr, nullptr};     // [0] coded features (lf0 / mgc / bap), [1] 16-bit waveform staging
  bool download_pending[2] = {false, false};
  ~wb200_batch() {
    if (upload_done) cudaEventDestroy(upload_done);
    for (cudaEvent_t e : download_done) if (e) cudaEventDestroy(e);
  }
};

namespace {
bool ensure_copy_streams(Context* c) {
  if (!c->copy_stream) {
    if (!WB_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) ||
        !WB_CUDA(cudaStreamCreateWithFlags(&c->upload_stream, cudaStreamNonBlocking)) ||
        !WB_CUDA(cudaEventCreateWithFlags(&c->copy_event, cudaEventDisableTiming)))
      return false;
  }
  return true;
}
enum { kDlCoded = 0, kDlWave = 1 };
// A bulk copy between pinned host memory and the device as a sequence of 8 MiB pieces.  A small transfer that the
// compute stream is waiting for (read_back's words, written by a kernel into mapped host memory) completed only
// when a bulk copy in flight on another stream did -- the features' 114 MB cost the end-to-end leg their whole
// PCIe time although nothing depended on them; between two pieces the small transfer gets through.
static bool bulk_copy_async(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st) {
  const size_t piece = (size_t)8 << 20;
  for (size_t o = 0; o < bytes; o += piece) {
    const size_t n = bytes - o < piece ? bytes - o : piece;
    if (!WB_CUDA(cudaMemcpyAsync(static_cast<char*>(dst) + o, static_cast<const char*>(src) + o, n, kind, st))) return false;
  }
  return true;
}
// ---- deferred bulk copies (wb200_set_copy_deferral) ---------------------------------------------------------
// Measured on the end-to-end leg: while a bulk device-to-host copy is in flight, a stage's small read-backs
// (list totals, pulse counts: a kernel writes them into mapped host memory and the host waits for it) complete
// only when the copy has drained, so a copy queued beside Dio or Synthesis costs its whole PCIe time -- twice that
// on an 8-GPU box whose GPUs share PCIe switches.  With deferral on, the asynchronous uploads / downloads are
// only RECORDED when the caller asks for them and are issued at the next safe point: at the start of StoneMask
// (or CheapTrick, or right before D4C's main kernel, whichever comes first: StoneMask -> CheapTrick -> D4C is
// ~60 ms per sub-batch of compute without a single read-back), or at the latest when
// something waits for them (wb200_batch_wait_downloads, wb200_sync, the next pass over the same batch, the stage
// that needs the uploaded samples).  Host buffers must stay valid until then, as they must for any async copy.
struct DeferredCopy { wb200_batch* h; std::function<bool()> issue; };
std::vector<DeferredCopy> g_deferred;
bool g_defer = false, g_in_flush = false;
bool flush_deferred(wb200_batch* only) {
  if (g_in_flush) return true;
  g_in_flush = true;
  bool ok = true;
  for (size_t i = 0; i < g_deferred.size();) {
    if (only && g_deferred[i].h != only) { ++i; continue; }
    std::function<bool()> fn = std::move(g_deferred[i].issue);
    g_deferred.erase(g_deferred.begin() + i);
    ok = fn() && ok;
  }
  g_in_flush = false;
  return ok;
}
bool defer_now() { return g_defer && !g_in_flush; }
bool wait_downloads(wb200_batch* h, int which) {        // the buffer `which` is about to be rewritten
  if (!flush_deferred(h)) return false;
  if (!h->download_pending[which]) return true;
  h->download_pending[which] = false;
  return WB_CUDA(cudaStreamWaitEvent(ctx()->stream, h->download_done[which], 0));
}
bool mark_downloads(wb200_batch* h, int which) {
  if (!h->download_done[which] && !WB_CUDA(cudaEventCreateWithFlags(&h->download_done[which], cudaEventDisableTiming))) return false;
  h->download_pending[which] = true;
  return WB_CUDA(cudaEventRecord(h->download_done[which], ctx()->copy_stream));
}
// every stage that reads the samples first waits (on the device) for an asynchronous upload
bool wait_upload(wb200_batch* h) {
  if (!flush_deferred(h)) return false;
  if (!h->upload_pending) return true;
  h->upload_pending = false;
  return WB_CUDA(cudaStreamWaitEvent(ctx()->stream, h->upload_done, 0));
}
}  // namespace

namespace wb {
void flush_deferred_copies() { flush_deferred(nullptr); }      // StoneMask / CheapTrick / D4C call this where no read-back follows
}
int wb200_set_copy_deferral(int on) {
  ApiGuard api_guard;
  g_defer = on != 0;
  return g_defer || flush_deferred(nullptr) ? 0 : 1;
}

namespace {

const double kNaN = std::numeric_limits<double>::quiet_NaN();

struct StageTimer {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  float* dst;
  cudaStream_t st;
  explicit StageTimer(float* d) : dst(d) {
    Context* c = ctx();
    st = c ? c->stream : nullptr;
    if (c && cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess)
      cudaEventRecord(e0, st);
  }
  ~StageTimer() {
    if (e0 && e1) {
      cudaEventRecord(e1, st);
      cudaEventSynchronize(e1);
      float ms = 0;
      if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) *dst = ms;
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  }
};

int samples_for_dio(int fs, int x_length, double frame_period) {
  return static_cast<int>(1000.0 * x_length / fs / frame_period) + 1;   // W/src/dio.cpp:638-640
}

__global__ void pcm16_to_double_kernel(const int16_t* __restrict__ pcm, const long long* __restrict__ src_off,
                                       const long long* __restrict__ x_off, const int* __restrict__ x_len,
                                       double* __restrict__ x) {
  const int u = blockIdx.y;
  const int n = x_len[u];
  const int16_t* s = pcm + src_off[u];
  double* d = x + x_off[u];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    d[i] = static_cast<double>(s[i]) / 32768.0;
}

// y (utterances 16-byte aligned) -> 16-bit samples, utterances back to back (one D2H copy later)
__global__ void y_to_pcm16_kernel(const double* __restrict__ y, const long long* __restrict__ y_off,
                                  const long long* __restrict__ out_off, const int* __restrict__ y_len,
                                  int16_t* __restrict__ out) {
  const int u = blockIdx.y;
  const int n = y_len[u];
  const double* __restrict__ src = y + y_off[u];
  int16_t* __restrict__ dst = out + out_off[u];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    // W/test/audioio.cpp:115-170: (short)(MyMax(-32768, MyMin(32767, (int)(x * 32767))))
    const double v = src[i] * 32767.0;
    int iv = (v != v) ? 0 : (v > 2147483000.0 ? 2147483000 : (v < -2147483000.0 ? -2147483000 : (int)v));
    iv = max(-32768, min(32767, iv));
    dst[i] = (int16_t)iv;
  }
}

__global__ void lf0_stats_kernel(const double* __restrict__ f0, int n, double* __restrict__ out3) {
  __shared__ double red[96];
  double v[3] = {0.0, 0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double f = f0[i];
    if (f > 0.0) { const double l = log(f); v[0] += 1.0; v[1] += l; v[2] += l * l; }
  }
  block_sum<3>(v, red);
  if (threadIdx.x == 0) { atomicAdd(&out3[0], v[0]); atomicAdd(&out3[1], v[1]); atomicAdd(&out3[2], v[2]); }
}

bool single_utt_batch(Batch* b, const double* x, int x_length, int fs, double frame_period,
                      int f0_length) {
  if (!batch_layout(b, fs, frame_period, 1, &x_length, &f0_length)) return false;
  Context* c = ctx();
  if (x_length > 0 && x)
    if (!WB_CUDA(cudaMemcpyAsync(b->x.p, x, (size_t)x_length * sizeof(double), cudaMemcpyHostToDevice, c->stream)))
      return false;
  return true;
}

bool upload_frames(Batch* b, const double* tpos, const double* f0, int n) {
  Context* c = ctx();
  std::vector<int> zeros(n > 0 ? n : 1, 0);
  if (n <= 0) return true;
  return WB_CUDA(cudaMemcpyAsync(b->frame_utt.p, zeros.data(), n * sizeof(int), cudaMemcpyHostToDevice, c->stream)) &&
         WB_CUDA(cudaMemcpyAsync(b->frame_t.p, tpos, n * sizeof(double), cudaMemcpyHostToDevice, c->stream)) &&
         WB_CUDA(cudaMemcpyAsync(b->f0.p, f0, n * sizeof(double), cudaMemcpyHostToDevice, c->stream)) &&
         WB_CUDA(cudaStreamSynchronize(c->stream));
}

void fill_nan(double* p, long long n) { for (long long i = 0; i < n; ++i) p[i] = kNaN; }

}  // namespace

extern "C" {

// =============================================================================================
// the WORLD API (drop-in)
// =============================================================================================
void InitializeDioOption(DioOption* option) {      // W/src/dio.cpp:649-665
  option->channels_in_octave = 2.0;
  option->f0_ceil = kCeilF0;
  option->f0_floor = kFloorF0;
  option->frame_period = 5;
  option->speed = 1;
  option->allowed_range = 0.1;
}
int GetSamplesForDIO(int fs, int x_length, double frame_period) {
  return samples_for_dio(fs, x_length, frame_period);
}
void Dio(const double* x, int x_length, int fs, const DioOption* option,
         double* temporal_positions, double* f0) {
  ApiGuard api_guard;
  const int n = samples_for_dio(fs, x_length, option->frame_period);
  for (int i = 0; i < n; ++i) temporal_positions[i] = i * option->frame_period / 1000.0;
  // W/src/dio.cpp:264-266: FixF0Contour returns without writing f0 for very short inputs
  const int voice_range_minimum =
      static_cast<int>(0.5 + 1000.0 / option->frame_period / option->f0_floor) * 2 + 1;
  if (n <= voice_range_minimum) return;
  Batch b;
  DioParams p = {option->f0_floor, option->f0_ceil, option->channels_in_octave,
                 option->frame_period, option->speed, option->allowed_range};
  bool ok = ctx() && single_utt_batch(&b, x, x_length, fs, option->frame_period, n) &&
            batch_default_frames(&b);
  if (ok) { StageTimer t(&g_times.dio); ok = dio_run(&b, p, b.f0_raw.p); }
  ok = ok && WB_CUDA(cudaMemcpyAsync(f0, b.f0_raw.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx()->stream)) &&
       WB_CUDA(cudaStreamSynchronize(ctx()->stream));
  if (!ok) fill_nan(f0, n);
}

void StoneMask(const double* x, int x_length, int fs, const double* temporal_positions,
               const double* f0, int f0_length, double* refined_f0) {
  ApiGuard api_guard;
  if (f0_length <= 0) return;
  Batch b;
  bool ok = ctx() && single_utt_batch(&b, x, x_length, fs, 5.0, f0_length) &&
            upload_frames(&b, temporal_positions, f0, f0_length);
  if (ok) {
    StageTimer t(&g_times.stonemask);
    ok = stonemask_run(b.view(), fs, f0_length, b.frame_utt.p, b.frame_t.p, b.f0.p, b.f0_raw.p);
  }
  ok = ok && WB_CUDA(cudaMemcpyAsync(refined_f0, b.f0_raw.p, f0_length * sizeof(double), cudaMemcpyDeviceToHost, ctx()->stream)) &&
       WB_CUDA(cudaStreamSynchronize(ctx()->stream));
  if (!ok) fill_nan(refined_f0, f0_length);
}

int GetFFTSizeForCheapTrick(int fs, const CheapTrickOption* option) {   // W/src/cheaptrick.cpp:191-194
  return static_cast<int>(pow(2.0, 1.0 + static_cast<int>(log(3.0 * fs / option->f0_floor + 1) / kLog2)));
}
double GetF0FloorForCheapTrick(int fs, int fft_size) { return 3.0 * fs / (fft_size - 3.0); }
void InitializeCheapTrickOption(int fs, CheapTrickOption* option) {     // :230-239
  option->q1 = -0.15;
  option->f0_floor = kFloorF0;
  option->fft_size = GetFFTSizeForCheapTrick(fs, option);
}

static void scatter_rows(const std::vector<double>& flat, int rows, int cols, double** dst) {
  for (int i = 0; i < rows; ++i) memcpy(dst[i], flat.data() + (size_t)i * cols, cols * sizeof(double));
}

void CheapTrick(const double* x, int x_length, int fs, const double* temporal_positions,
                const double* f0, int f0_length, const CheapTrickOption* option,
                double** spectrogram) {
  ApiGuard api_guard;
  if (f0_length <= 0) return;
  const int cols = option->fft_size / 2 + 1;
  Batch b;
  std::vector<double> flat((size_t)f0_length * cols, kNaN);
  bool ok = ctx() && single_utt_batch(&b, x, x_length, fs, 5.0, f0_length) &&
            upload_frames(&b, temporal_positions, f0, f0_length) && b.sp.alloc(flat.size());
  if (ok) {
    StageTimer t(&g_times.cheaptrick);
    ok = cheaptrick_run(b.view(), fs, f0_length, b.frame_utt.p, b.frame_t.p, b.f0.p,
                        option->fft_size, option->q1, b.sp.p);
  }
  ok = ok && WB_CUDA(cudaMemcpyAsync(flat.data(), b.sp.p, flat.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx()->stream)) &&
       WB_CUDA(cudaStreamSynchronize(ctx()->stream));
  if (!ok) std::fill(flat.begin(), flat.end(), kNaN);
  scatter_rows(flat, f0_length, cols, spectrogram);
}

void InitializeD4COption(D4COption* option) { option->threshold = kThreshold; }

void D4C(const double* x, int x_length, int fs, const double* temporal_positions,
         const double* f0, int f0_length, int fft_size, const D4COption* option,
         double** aperiodicity) {
  ApiGuard api_guard;
  if (f0_length <= 0) return;
  const int cols = fft_size / 2 + 1;
  Batch b;
  std::vector<double> flat((size_t)f0_length * cols, kNaN);
  bool ok = ctx() && single_utt_batch(&b, x, x_length, fs, 5.0, f0_length) &&
            upload_frames(&b, temporal_positions, f0, f0_length) && b.ap.alloc(flat.size());
  if (ok) {
    StageTimer t(&g_times.d4c);
    ok = d4c_run(b.view(), fs, f0_length, b.frame_utt.p, b.frame_t.p, b.f0.p, fft_size,
                 option->threshold, b.ap.p);
  }
  ok = ok && WB_CUDA(cudaMemcpyAsync(flat.data(), b.ap.p, flat.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx()->stream)) &&
       WB_CUDA(cudaStreamSynchronize(ctx()->stream));
  if (!ok) std::fill(flat.begin(), flat.end(), kNaN);
  scatter_rows(flat, f0_length, cols, aperiodicity);
}

void Synthesis(const double* f0, int f0_length, const double* const* spectrogram,
               const double* const* aperiodicity, int fft_size, double frame_period, int fs,
               int y_length, double* y) {
  ApiGuard api_guard;
  if (y_length <= 0) return;
  const int cols = fft_size / 2 + 1;
  Batch b;
  int zero_len = 0;
  bool ok = ctx() && f0_length >= 2 && batch_layout(&b, fs, frame_period, 1, &zero_len, &f0_length);
  std::vector<double> flat((size_t)(f0_length > 0 ? f0_length : 1) * cols);
  Context* c = ctx();
  ok = ok && b.sp.alloc(flat.size()) && b.ap.alloc(flat.size());
  if (ok) {
    b.fft_size = fft_size;
    for (int i = 0; i < f0_length; ++i) memcpy(flat.data() + (size_t)i * cols, spectrogram[i], cols * sizeof(double));
    ok = WB_CUDA(cudaMemcpyAsync(b.sp.p, flat.data(), flat.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream)) &&
         WB_CUDA(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < f0_length; ++i) memcpy(flat.data() + (size_t)i * cols, aperiodicity[i], cols * sizeof(double));
    ok = ok && WB_CUDA(cudaMemcpyAsync(b.ap.p, flat.data(), flat.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream)) &&
         WB_CUDA(cudaMemcpyAsync(b.f0.p, f0, f0_length * sizeof(double), cudaMemcpyHostToDevice, c->stream)) &&
         WB_CUDA(cudaStreamSynchronize(c->stream));
  }
  if (ok) { StageTimer t(&g_times.synthesis); ok = synthesis_run(&b, &y_length); }
  ok = ok && WB_CUDA(cudaMemcpyAsync(y, b.y.p, (size_t)y_length * sizeof(double), cudaMemcpyDeviceToHost, c->stream)) &&
       WB_CUDA(cudaStreamSynchronize(c->stream));
  if (!ok) fill_nan(y, y_length);
}

void InitializeHarvestOption(HarvestOption* option) {   // W/src/harvest.cpp:1257-1262
  option->f0_ceil = kCeilF0;
  option->f0_floor = kFloorF0;
  option->frame_period = 5;
}
int GetSamplesForHarvest(int fs, int x_length, double frame_period) {
  return static_cast<int>(1000.0 * x_length / fs / frame_period) + 1;
}
void Harvest(const double* x, int x_length, int fs, const HarvestOption* option,
             double* temporal_positions, double* f0) {
  ApiGuard api_guard;
  const int n = GetSamplesForHarvest(fs, x_length, option->frame_period);
  for (int i = 0; i < n; ++i) temporal_positions[i] = i * option->frame_period / 1000.0;
  Batch b;
  HarvestParams p = {option->f0_floor, option->f0_ceil, option->frame_period};
  bool ok = ctx() && single_utt_batch(&b, x, x_length, fs, option->frame_period, n) &&
            batch_default_frames(&b);
  if (ok) { StageTimer t(&g_times.harvest); ok = harvest_run(&b, p, b.f0.p); }
  ok = ok && WB_CUDA(cudaMemcpyAsync(f0, b.f0.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx()->stream)) &&
       WB_CUDA(cudaStreamSynchronize(ctx()->stream));
  if (!ok) fill_nan(f0, n);
}

int GetNumberOfAperiodicities(int fs) {              // W/src/codec.cpp:211-214
  return static_cast<int>(fmin(kUpperLimit, fs / 2.0 - kFrequencyInterval) / kFrequencyInterval);
}

// rows <-> coded through the device codec; host code only gathers / scatters the double** rows
static void codec_host(const double* const* in, int f0_length, int fs, int fft_size, int ndim, bool encode,
                       double** out) {
  if (f0_length <= 0) return;
  const int cols_in = encode ? fft_size / 2 + 1 : ndim, cols_out = encode ? ndim : fft_size / 2 + 1;
  std::vector<double> flat_in((size_t)f0_length * cols_in), flat_out((size_t)f0_length * cols_out, kNaN);
  for (int i = 0; i < f0_length; ++i) memcpy(flat_in.data() + (size_t)i * cols_in, in[i], cols_in * sizeof(double));
  Context* c = ctx();
  DevBuf<double> d_in, d_out;
  bool ok = c && d_in.alloc(flat_in.size()) && d_out.alloc(flat_out.size()) &&
            WB_CUDA(cudaMemcpyAsync(d_in.p, flat_in.data(), flat_in.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (ok) ok = encode ? codec_encode_run(d_in.p, f0_length, fs, fft_size, ndim, 1.0, 0.0, 0.0, d_out.p)
                      : codec_decode_run(d_in.p, f0_length, fs, fft_size, ndim, d_out.p);
  ok = ok && WB_CUDA(cudaMemcpyAsync(flat_out.data(), d_out.p, flat_out.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream)) &&
       WB_CUDA(cudaStreamSynchronize(c->stream));
  if (!ok) std::fill(flat_out.begin(), flat_out.end(), kNaN);
  scatter_rows(flat_out, f0_length, cols_out, out);
}
void CodeSpectralEnvelope(const double* const* spectrogram, int f0_length, int fs, int fft_size,
                          int number_of_dimensions, double** coded_spectral_envelope) {
  ApiGuard api_guard;
  codec_host(spectrogram, f0_length, fs, fft_size, number_of_dimensions, true, coded_spectral_envelope);
}
void DecodeSpectralEnvelope(const double* const* coded_spectral_envelope, int f0_length, int fs, int fft_size,
                            int number_of_dimensions, double** spectrogram) {
  ApiGuard api_guard;
  codec_host(coded_spectral_envelope, f0_length, fs, fft_size, number_of_dimensions, false, spectrogram);
}

// =============================================================================================
// extension API
// =============================================================================================
const char* wb200_last_error(void) { return last_error(); }
int wb200_init(int device) {
  ApiGuard api_guard;
  if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", device); return 1; }
  return ctx() ? 0 : 1;
}
unsigned long long wb200_launch_count(void) { return g_launch_count; }
void wb200_stage_times(float* o) {
  o[0] = g_times.dio; o[1] = g_times.stonemask; o[2] = g_times.cheaptrick;
  o[3] = g_times.d4c; o[4] = g_times.synthesis; o[5] = g_times.harvest;
}
int wb200_set_stream(void* stream) {
  ApiGuard api_guard;
  if (!ctx()) return 1;
  set_stream(reinterpret_cast<cudaStream_t>(stream));
  return 0;
}
void wb200_kernel_timing(int on) { kernel_timing_enable(on != 0); }
void wb200_kernel_times_reset(void) { kernel_times_reset(); }
int wb200_kernel_time(const char* name, double* ms_total, long long* launches) {
  return kernel_time_query(name, ms_total, launches) ? 0 : 1;
}
double wb200_measure_fma_peak(int fp64) {
  ApiGuard api_guard; return measure_fma_peak(fp64 != 0); }
int wb200_option(const char* name) { return name ? option(name) : 0; }
int wb200_trim(void) { return trim_pool() ? 0 : 1; }
int wb200_sync(void) {
  ApiGuard api_guard;
  Context* c = ctx();
  if (!c) return 1;
  bool ok = flush_deferred(nullptr);
  ok = WB_CUDA(cudaStreamSynchronize(c->stream)) && ok;
  if (c->copy_stream) ok = WB_CUDA(cudaStreamSynchronize(c->copy_stream)) && ok;
  if (c->upload_stream) ok = WB_CUDA(cudaStreamSynchronize(c->upload_stream)) && ok;
  return ok ? 0 : 1;
}
int wb200_randn_stream(double* out, long long n) {
  ApiGuard api_guard;
  Context* c = ctx();
  if (!c || !ensure_randn((size_t)n)) return 1;
  std::vector<uint32_t> h((size_t)n);
  if (!WB_CUDA(cudaMemcpy(h.data(), c->d_randn, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost))) return 1;
  for (long long i = 0; i < n; ++i) out[i] = h[i] / 268435456.0 - 6.0;
  return 0;
}

wb200_batch* wb200_batch_create(int fs, double frame_period, int n_utt, const int* x_lengths) {
  ApiGuard api_guard;
  if (!ctx()) return nullptr;
  wb200_batch* h = new wb200_batch();
  std::vector<int> f_len(n_utt > 0 ? n_utt : 1);
  for (int u = 0; u < n_utt; ++u) f_len[u] = samples_for_dio(fs, x_lengths[u], frame_period);
  if (!batch_layout(&h->b, fs, frame_period, n_utt, x_lengths, f_len.data()) ||
      !batch_default_frames(&h->b)) {
    delete h;
    return nullptr;
  }
  return h;
}
void wb200_batch_destroy(wb200_batch* h) {
  ApiGuard api_guard;
  if (!h) return;
  if (ctx()) wb200_sync();             // library, upload and download streams
  delete h;
}
int wb200_batch_total_frames(const wb200_batch* h) { return h->b.total_frames; }
long long wb200_batch_total_samples(const wb200_batch* h) {
  long long s = 0;
  for (int v : h->b.h_x_len) s += v;
  return s;
}
int wb200_batch_frame_layout(const wb200_batch* h, int* f_off, int* f_len) {
  for (int u = 0; u < h->b.n_utt; ++u) { f_off[u] = h->b.h_f_off[u]; f_len[u] = h->b.h_f_len[u]; }
  return 0;
}

static int convert_pcm(wb200_batch* h, const int16_t* dev_pcm) {
  Batch& b = h->b;
  Context* c = ctx();
  if (!wait_upload(h)) return 1;           // a deferred / in-flight asynchronous upload of this batch comes first
  std::vector<long long> src(b.n_utt);
  long long o = 0;
  for (int u = 0; u < b.n_utt; ++u) { src[u] = o; o += b.h_x_len[u]; }
  DevBuf<long long> d_src;
  if (!d_src.alloc(b.n_utt)) return 1;
  if (!WB_CUDA(cudaMemcpyAsync(d_src.p, src.data(), b.n_utt * sizeof(long long), cudaMemcpyHostToDevice, c->stream))) return 1;
  dim3 grid(64, b.n_utt);
  pcm16_to_double_kernel<<<grid, 256, 0, c->stream>>>(dev_pcm, d_src.p, b.x_off.p, b.x_len.p, b.x.p);
  WB_LAUNCH_CHECK();
  return WB_CUDA(cudaStreamSynchronize(c->stream)) ? 0 : 1;
}
int wb200_batch_upload_pcm16(wb200_batch* h, const int16_t* host_pcm) {
  ApiGuard api_guard;
  Context* c = ctx();
  if (!c) return 1;
  const long long n = wb200_batch_total_samples(h);
  if (!wait_upload(h)) return 1;           // a deferred / in-flight asynchronous upload of this batch comes first
  if (!h->pcm_stage.alloc((size_t)n)) return 1;
  if (!WB_CUDA(cudaMemcpyAsync(h->pcm_stage.p, host_pcm, (size_t)n * sizeof(int16_t), cudaMemcpyHostToDevice, c->stream))) return 1;
  return convert_pcm(h, h->pcm_stage.p);
}
// Asynchronous variant: the copy and the int16 -> double conversion run on the upload stream, so
// they overlap whatever the library stream is computing (the previous batch); the stages of THIS
// batch wait for it on the device.  host_pcm must be pinned and stay valid until the first stage
// of this batch has been launched.
int wb200_batch_upload_pcm16_async(wb200_batch* h, const int16_t* host_pcm) {
  ApiGuard api_guard;
  Context* c = ctx();
  if (!c || !ensure_copy_streams(c)) return 1;
  if (defer_now()) {
    g_deferred.push_back({h, [h, host_pcm]() { return wb200_batch_upload_pcm16_async(h, host_pcm) == 0; }});
    return 0;
  }
  Batch& b = h->b;
  const long long n = wb200_batch_total_samples(h);
  if (!h->upload_done && !WB_CUDA(cudaEventCreateWithFlags(&h->upload_done, cudaEventDisableTiming))) return 1;
  if (!h->pcm_stage.p || !h->src_off.p) {            // first use: allocate on the library stream, once
    std::vector<long long> src(b.n_utt > 0 ? b.n_utt : 1);
    long long o = 0;
    for (int u = 0; u < b.n_utt; ++u) { src[u] = o; o += b.h_x_len[u]; }
    if (!h->pcm_stage.alloc((size_t)n + 1) || !h->src_off.alloc(b.n_utt)) return 1;
    if (!WB_CUDA(cudaMemcpyAsync(h->src_off.p, src.data(), b.n_utt * sizeof(long long), cudaMemcpyHostToDevice, c->stream)) ||
        !WB_CUDA(cudaStreamSynchronize(c->stream)))
      return 1;
  }
  // x may still be read by work queued on the library stream: order the upload behind it
  if (!WB_CUDA(cudaEventRecord(c->copy_event, c->stream)) || !WB_CUDA(cudaStreamWaitEvent(c->upload_stream, c->copy_event, 0))) return 1;
  if (n > 0) {
    if (!bulk_copy_async(h->pcm_stage.p, host_pcm, (size_t)n * sizeof(int16_t), cudaMemcpyHostToDevice, c->upload_stream)) return 1;
    pcm16_to_double_kernel<<<dim3(64, b.n_utt), 256, 0, c->upload_stream>>>(h->pcm_stage.p, h->src_off.p, b.x_off.p, b.x_len.p, b.x.p);
    WB_LAUNCH_CHECK();
  }
  if (!WB_CUDA(cudaEventRecord(h->upload_done, c->upload_stream))) return 1;
  h->upload_pending = true;
  return 0;
}
int wb200_batch_set_pcm16_device(wb200_batch* h, const int16_t* dev_pcm) {
  ApiGuard api_guard;
  if (!ctx()) return 1;
  return convert_pcm(h, dev_pcm);
}
int wb200_batch_upload_f64(wb200_batch* h, const double* host_x) {
  ApiGuard api_guard;
  Context* c = ctx();
  if (!c) return 1;
  Batch& b = h->b;
  if (!wait_upload(h)) return 1;           // a deferred / in-flight asynchronous upload of this batch comes first
  long long o = 0;
  for (int u = 0; u < b.n_utt; ++u) {
    if (b.h_x_len[u] > 0 &&
        !WB_CUDA(cudaMemcpyAsync(b.x.p + b.h_x_off[u], host_x + o, (size_t)b.h_x_len[u] * sizeof(double), cudaMemcpyHostToDevice, c->stream)))
      return 1;
    o += b.h_x_len[u];
  }
  return WB_CUDA(cudaStreamSynchronize(c->stream)) ? 0 : 1;
}

int wb200_batch_dio(wb200_batch* h, const DioOption* o) {
  ApiGuard api_guard;
  if (!ctx() || !wait_upload(h)) return 1;
  DioParams p = {o->f0_floor, o->f0_ceil, o->channels_in_octave, o->frame_period, o->speed, o->allowed_range};
  StageTimer t(&g_times.dio);
  return dio_run(&h->b, p, h->b.f0_raw.p) ? 0 : 1;
}
int wb200_batch_stonemask(wb200_batch* h) {
  ApiGuard api_guard;
  if (!ctx() || !wait_upload(h)) return 1;
  Batch& b = h->b;
  StageTimer t(&g_times.stonemask);
  return stonemask_run(b.view(), b.fs, b.total_frames, b.frame_utt.p, b.frame_t.p, b.f0_raw.p, b.f0.p) ? 0 : 1;
}
int wb200_batch_harvest(wb200_batch* h, const HarvestOption* o) {
  ApiGuard api_guard;
  if (!ctx() || !wait_upload(h)) return 1;
  HarvestParams p = {o->f0_floor, o->f0_ceil, o->frame_period};
  StageTimer t(&g_times.harvest);
  return harvest_run(&h->b, p, h->b.f0.p) ? 0 : 1;
}
int wb200_batch_cheaptrick(wb200_batch* h, const CheapTrickOption* o) {
  ApiGuard api_guard;
  if (!ctx() || !wait_upload(h)) return 1;
  Batch& b = h->b;
  b.fft_size = o->fft_size;
  if (!b.sp.alloc((size_t)b.total_frames * (o->fft_size / 2 + 1))) return 1;
  StageTimer t(&g_times.cheaptrick);
  return cheaptrick_run(b.view(), b.fs, b.total_frames, b.frame_utt.p, b.frame_t.p, b.f0.p, o->fft_size, o->q1, b.sp.p) ? 0 : 1;
}
int wb200_batch_d4c(wb200_batch* h, int fft_size, const D4COption* o) {
  ApiGuard api_guard;
  if (!ctx() || !wait_upload(h)) return 1;
  Batch& b = h->b;
  b.fft_size = fft_size;
  if (!b.ap.alloc((size_t)b.total_frames * (fft_size / 2 + 1))) return 1;
  StageTimer t(&g_times.d4c);
  return d4c_run(b.view(), b.fs, b.total_frames, b.frame_utt.p, b.frame_t.p, b.f0.p, fft_size, o->threshold, b.ap.p) ? 0 : 1;
}
int wb200_batch_synthesis(wb200_batch* h, const int* y_lengths) {
  ApiGuard api_guard;
  if (!ctx()) return 1;
  Batch& b = h->b;
  std::vector<int> yl(b.n_utt > 0 ? b.n_utt : 1);
  for (int u = 0; u < b.n_utt; ++u)
    yl[u] = y_lengths ? y_lengths[u]
                      : static_cast<int>((b.h_f_len[u] - 1) * b.frame_period / 1000.0 * b.fs) + 1;
  StageTimer t(&g_times.synthesis);
  return synthesis_run(&b, yl.data()) ? 0 : 1;
}

static int d2h(void* dst, const void* src, size_t bytes) {
  Context* c = ctx();
  if (!c || !src) { set_error("result not available (stage not run?)"); return 1; }
  return (WB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream)) &&
          WB_CUDA(cudaStreamSynchronize(c->stream))) ? 0 : 1;
}
static int h2d(void* dst, const void* src, size_t bytes) {
  Context* c = ctx();
  if (!c) return 1;
  return (WB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream)) &&
          WB_CUDA(cudaStreamSynchronize(c->stream))) ? 0 : 1;
}
int wb200_batch_get_f0(wb200_batch* h, double* out, int refined) {
  ApiGuard api_guard;
  return d2h(out, refined ? h->b.f0.p : h->b.f0_raw.p, (size_t)h->b.total_frames * sizeof(double));
}
int wb200_batch_set_f0(wb200_batch* h, const double* in, int refined) {
  ApiGuard api_guard;
  return h2d(refined ? h->b.f0.p : h->b.f0_raw.p, in, (size_t)h->b.total_frames * sizeof(double));
}
int wb200_batch_get_sp(wb200_batch* h, double* out) {
  ApiGuard api_guard;
  return d2h(out, h->b.sp.p, (size_t)h->b.total_frames * (h->b.fft_size / 2 + 1) * sizeof(double));
}
int wb200_batch_get_ap(wb200_batch* h, double* out) {
  ApiGuard api_guard;
  return d2h(out, h->b.ap.p, (size_t)h->b.total_frames * (h->b.fft_size / 2 + 1) * sizeof(double));
}
int wb200_batch_set_sp_ap(wb200_batch* h, int fft_size, const double* sp, const double* ap) {
  ApiGuard api_guard;
  Batch& b = h->b;
  b.fft_size = fft_size;
  const size_t n = (size_t)b.total_frames * (fft_size / 2 + 1);
  if (!b.sp.alloc(n) || !b.ap.alloc(n)) return 1;
  return h2d(b.sp.p, sp, n * sizeof(double)) || h2d(b.ap.p, ap, n * sizeof(double));
}
__global__ void widen_f32_kernel(const float* __restrict__ in, long long n, double* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (double)in[i];                                              // ToDouble, W/test/synth.cpp:66-70
}
int wb200_batch_set_params_f32(wb200_batch* h, int fft_size, const float* f0, const float* sp, const float* ap) {
  ApiGuard api_guard;
  Context* c = ctx();
  if (!c) return 1;
  Batch& b = h->b;
  if (fft_size < 32 || (fft_size & (fft_size - 1))) { set_error("set_params_f32: fft_size %d", fft_size); return 1; }
  b.fft_size = fft_size;
  const size_t F = (size_t)b.total_frames, n = F * (fft_size / 2 + 1);
  DevBuf<float> stage;
  if (!b.sp.alloc(n) || !b.ap.alloc(n) || !stage.alloc(n + 1)) return 1;
  if (F == 0) return 0;
  const float* src[3] = {f0, sp, ap};
  double* dst[3] = {b.f0.p, b.sp.p, b.ap.p};
  const size_t cnt[3] = {F, n, n};
  for (int k = 0; k < 3; ++k) {
    if (!src[k]) continue;
    if (!WB_CUDA(cudaMemcpyAsync(stage.p, src[k], cnt[k] * sizeof(float), cudaMemcpyHostToDevice, c->stream))) return 1;
    widen_f32_kernel<<<148 * 8, 256, 0, c->stream>>>(stage.p, (long long)cnt[k], dst[k]);
    WB_LAUNCH_CHECK();
  }
  return WB_CUDA(cudaStreamSynchronize(c->stream)) ? 0 : 1;
}
long long wb200_batch_total_y(const wb200_batch* h) {
  long long s = 0;
  for (int v : h->b.h_y_len) s += v;
  return s;
}
int wb200_batch_y_layout(const wb200_batch* h, long long* y_off, int* y_len) {
  long long o = 0;
  for (size_t u = 0; u < h->b.h_y_len.size(); ++u) { y_off[u] = o; y_len[u] = h->b.h_y_len[u]; o += h->b.h_y_len[u]; }
  return 0;
}
int wb200_batch_get_y(wb200_batch* h, double* out) {
  ApiGuard api_guard;
  Context* c = ctx();
  Batch& b = h->b;
  if (!c || !b.y.p) { set_error("synthesis has not been run"); return 1; }
  long long o = 0;
  for (int u = 0; u < b.n_utt; ++u) {
    if (b.h_y_len[u] > 0 &&
        !WB_CUDA(cudaMemcpyAsync(out + o, b.y.p + b.h_y_off[u], (size_t)b.h_y_len[u] * sizeof(double), cudaMemcpyDeviceToHost, c->stream)))
      return 1;
    o += b.h_y_len[u];
  }
  return WB_CUDA(cudaStreamSynchronize(c->stream)) ? 0 : 1;
}
// 16-bit staging of the waveform: pcm_out and the per-utterance output offsets are (re)built whenever
// the y layout of the last Synthesis call differs from the one they were built for (y_lengths is a
// per-call argument of wb200_batch_synthesis).  An asynchronous copy of the previous pass may still
// read pcm_out, so the library stream first waits for it -- before the buffer can be re-allocated.
static bool prepare_pcm_out(wb200_batch* h, long long n) {
  Context* c = ctx();
  Batch& b = h->b;
  if (!wait_downloads(h, kDlWave)) return false;
  if (h->pcm_out.p && h->out_off.p && h->out_layout == b.h_y_len && h->pcm_out.n >= (size_t)n + 1) return true;
  std::vector<long long> cum(b.n_utt > 0 ? b.n_utt : 1);
  long long o = 0;
  for (int u = 0; u < b.n_utt; ++u) { cum[u] = o; o += b.h_y_len[u]; }
  if (!h->pcm_out.alloc((size_t)n + 1) || !h->out_off.alloc(b.n_utt)) return false;
  if (b.n_utt > 0 &&
      (!WB_CUDA(cudaMemcpyAsync(h->out_off.p, cum.data(), b.n_utt * sizeof(long long), cudaMemcpyHostToDevice, c->stream)) ||
       !WB_CUDA(cudaStreamSynchronize(c->stream))))            // cum is a host temporary
    return false;
  h->out_layout = b.h_y_len;
  return true;
}
int wb200_batch_get_y_pcm16(wb200_batch* h, int16_t* out) {
  ApiGuard api_guard;
  Context* c = ctx();
  Batch& b = h->b;
  if (!c || !b.y.p) { set_error("synthesis has not been run"); return 1; }
  const long long n = wb200_batch_total_y(h);
  if (b.n_utt == 0 || n == 0) return 0;
  if (!prepare_pcm_out(h, n)) return 1;
  y_to_pcm16_kernel<<<dim3(64, b.n_utt), 256, 0, c->stream>>>(b.y.p, b.y_off.p, h->out_off.p, b.y_len.p, h->pcm_out.p);
  WB_LAUNCH_CHECK();
  return (WB_CUDA(cudaMemcpyAsync(out, h->pcm_out.p, (size_t)n * sizeof(int16_t), cudaMemcpyDeviceToHost, c->stream)) &&
          WB_CUDA(cudaStreamSynchronize(c->stream))) ? 0 : 1;
}
// The conversion runs on the library stream (after Synthesis), the copy on the download stream:
// it overlaps the next batch's computation.  `out` must be pinned; valid after wb200_sync().
int wb200_batch_get_y_pcm16_async(wb200_batch* h, int16_t* out) {
  ApiGuard api_guard;
  Context* c = ctx();
  Batch& b = h->b;
  if (!c || !b.y.p) { set_error("synthesis has not been run"); return 1; }
  if (!ensure_copy_streams(c)) return 1;
  if (defer_now()) {
    g_deferred.push_back({h, [h, out]() { return wb200_batch_get_y_pcm16_async(h, out) == 0; }});
    return 0;
  }
  const long long n = wb200_batch_total_y(h);
  if (b.n_utt == 0 || n == 0) return 0;
  if (!prepare_pcm_out(h, n)) return 1;                // also: the previous pass's copy may still read pcm_out
  y_to_pcm16_kernel<<<dim3(64, b.n_utt), 256, 0, c->stream>>>(b.y.p, b.y_off.p, h->out_off.p, b.y_len.p, h->pcm_out.p);
  WB_LAUNCH_CHECK();
  return (WB_CUDA(cudaEventRecord(c->copy_event, c->stream)) && WB_CUDA(cudaStreamWaitEvent(c->copy_stream, c->copy_event, 0)) &&
          bulk_copy_async(out, h->pcm_out.p, (size_t)n * sizeof(int16_t), cudaMemcpyDeviceToHost, c->copy_stream) &&
          mark_downloads(h, kDlWave)) ? 0 : 1;
}
// one utterance's slice of every result (any pointer may be NULL)
int wb200_batch_get_utterance(wb200_batch* h, int utt, double* f0_raw, double* f0, double* sp, double* ap, double* y) {
  ApiGuard api_guard;
  Context* c = ctx();
  Batch& b = h->b;
  if (!c) return 1;
  if (utt < 0 || utt >= b.n_utt) { set_error("get_utterance: utterance %d of %d", utt, b.n_utt); return 1; }
  const size_t fo = (size_t)b.h_f_off[utt], fl = (size_t)b.h_f_len[utt], H = (size_t)(b.fft_size / 2 + 1);
  struct { double* dst; const double* src; size_t n; } parts[5] = {
      {f0_raw, b.f0_raw.p ? b.f0_raw.p + fo : nullptr, fl}, {f0, b.f0.p ? b.f0.p + fo : nullptr, fl},
      {sp, b.sp.p ? b.sp.p + fo * H : nullptr, fl * H}, {ap, b.ap.p ? b.ap.p + fo * H : nullptr, fl * H},
      {y, b.y.p && (size_t)utt < b.h_y_len.size() ? b.y.p + b.h_y_off[utt] : nullptr,
       (size_t)utt < b.h_y_len.size() ? (size_t)b.h_y_len[utt] : 0}};
  for (auto& q : parts) {
    if (!q.dst || q.n == 0) continue;
    if (!q.src) { set_error("get_utterance: result not available (stage not run?)"); return 1; }
    if (!WB_CUDA(cudaMemcpyAsync(q.dst, q.src, q.n * sizeof(double), cudaMemcpyDeviceToHost, c->stream))) return 1;
  }
  return WB_CUDA(cudaStreamSynchronize(c->stream)) ? 0 : 1;
}
// Host-side wait for the asynchronous result copies of THIS batch only (coded features, 16-bit waveform):
// a pipelined caller hands batch i's host buffers to its consumer while batch i + 1 already computes.
int wb200_batch_wait_downloads(wb200_batch* h) {
  ApiGuard api_guard;
  if (!ctx()) return 1;
  bool ok = flush_deferred(h);
  for (int k = 0; k < 2; ++k)
    if (h->download_done[k]) ok = WB_CUDA(cudaEventSynchronize(h->download_done[k])) && ok;
  return ok ? 0 : 1;
}
void* wb200_batch_device_ptr(wb200_batch* h, const char* which) {
  Batch& b = h->b;
  const std::string w(which);
  if (w == "x") return b.x.p;
  if (w == "f0_raw") return b.f0_raw.p;
  if (w == "f0") return b.f0.p;
  if (w == "sp") return b.sp.p;
  if (w == "ap") return b.ap.p;
  if (w == "y") return b.y.p;
  return nullptr;
}
int wb200_batch_code(wb200_batch* h, int mgc_dim, int bap_dim) {
  ApiGuard api_guard;
  if (!ctx() || !wait_downloads(h, kDlCoded)) return 1;   // the previous pass's copies may still read lf0 / mgc / bap
  return batch_code_features(&h->b, mgc_dim, bap_dim) ? 0 : 1;
}
int wb200_batch_get_coded(wb200_batch* h, float* lf0, float* mgc, float* bap) {
  ApiGuard api_guard;
  Batch& b = h->b;
  const size_t F = (size_t)b.total_frames;
  if (lf0 && d2h(lf0, b.lf0.p, F * sizeof(float))) return 1;
  if (mgc && d2h(mgc, b.mgc.p, F * b.mgc_dim * sizeof(float))) return 1;
  if (bap && d2h(bap, b.bap.p, F * b.bap_dim * sizeof(float))) return 1;
  return 0;
}
// Same copies on a second stream, ordered after everything queued so far on the library stream;
// they run while later stages (Synthesis) compute.  Host buffers must be pinned and stay valid
// until wb200_sync() returns.
int wb200_batch_get_coded_async(wb200_batch* h, float* lf0, float* mgc, float* bap) {
  ApiGuard api_guard;
  Context* c = ctx();
  if (!c) return 1;
  Batch& b = h->b;
  if (!b.mgc.p) { set_error("features have not been coded"); return 1; }
  if (!ensure_copy_streams(c)) return 1;
  if (defer_now()) {
    g_deferred.push_back({h, [h, lf0, mgc, bap]() { return wb200_batch_get_coded_async(h, lf0, mgc, bap) == 0; }});
    return 0;
  }
  const size_t F = (size_t)b.total_frames;
  bool ok = WB_CUDA(cudaEventRecord(c->copy_event, c->stream)) &&
            WB_CUDA(cudaStreamWaitEvent(c->copy_stream, c->copy_event, 0));
  if (ok && lf0) ok = bulk_copy_async(lf0, b.lf0.p, F * sizeof(float), cudaMemcpyDeviceToHost, c->copy_stream);
  if (ok && mgc) ok = bulk_copy_async(mgc, b.mgc.p, F * b.mgc_dim * sizeof(float), cudaMemcpyDeviceToHost, c->copy_stream);
  if (ok && bap) ok = bulk_copy_async(bap, b.bap.p, F * b.bap_dim * sizeof(float), cudaMemcpyDeviceToHost, c->copy_stream);
  return ok && mark_downloads(h, kDlCoded) ? 0 : 1;
}
__global__ void mgc_unscale_kernel(const float* __restrict__ mgc, long long n, int ndim, double* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = (double)mgc[i] - ((i % ndim) == 0 ? 12.0 : 0.0);           // undo c0 + 12 (analysis.cpp:307)
}
__global__ void sp_unscale_kernel(double* __restrict__ sp, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) sp[i] *= 1e-4;                                            // undo sp * 1e4 (analysis.cpp:297)
}
int wb200_batch_decode_mgc(wb200_batch* h, int fft_size, int mgc_dim, const float* host_mgc) {
  ApiGuard api_guard;
  Context* c = ctx();
  if (!c) return 1;
  Batch& b = h->b;
  const long long F = b.total_frames, n = F * mgc_dim, ns = F * (fft_size / 2 + 1);
  DevBuf<float> d_f;
  DevBuf<double> d_c;
  if (!d_f.alloc((size_t)n) || !d_c.alloc((size_t)n) || !b.sp.alloc((size_t)ns)) return 1;
  b.fft_size = fft_size;
  if (F == 0) return 0;
  if (!WB_CUDA(cudaMemcpyAsync(d_f.p, host_mgc, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, c->stream))) return 1;
  mgc_unscale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_f.p, n, mgc_dim, d_c.p);
  WB_LAUNCH_CHECK();
  if (!codec_decode_run(d_c.p, (int)F, b.fs, fft_size, mgc_dim, b.sp.p)) return 1;
  sp_unscale_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, c->stream>>>(b.sp.p, ns);
  WB_LAUNCH_CHECK();
  return WB_CUDA(cudaStreamSynchronize(c->stream)) ? 0 : 1;
}
// The coded branch of the synth tool (W/test/synth.cpp:171-247, spec_dimension != 0): float32 lf0 / mgc / bap
// files -> f0 = exp(lf0) (0 stays 0), sp = DecodeSpectralEnvelope(mgc with c0 - 12) / 1e4, ap = exp(mgc2sp(bap
// with c0 + 9.210340)) / 1e4.  Any pointer may be NULL (that parameter of the batch is left as it is).
int wb200_batch_set_coded_f32(wb200_batch* h, int fft_size, int mgc_dim, int bap_dim, const float* lf0, const float* mgc,
                              const float* bap) {
  ApiGuard api_guard;
  Context* c = ctx();
  if (!c) return 1;
  Batch& b = h->b;
  if (fft_size < 32 || (fft_size & (fft_size - 1))) { set_error("set_coded_f32: fft_size %d", fft_size); return 1; }
  const size_t F = (size_t)b.total_frames, H = (size_t)(fft_size / 2 + 1);
  if (mgc && wb200_batch_decode_mgc(h, fft_size, mgc_dim, mgc)) return 1;
  b.fft_size = fft_size;
  DevBuf<float> stage;
  if (lf0) {
    if (!stage.alloc(F + 1)) return 1;
    if (F > 0 && (!WB_CUDA(cudaMemcpyAsync(stage.p, lf0, F * sizeof(float), cudaMemcpyHostToDevice, c->stream)) ||
                  !lf0_to_f0_run(stage.p, (int)F, b.f0.p)))
      return 1;
  }
  DevBuf<float> stage2;
  if (bap) {
    if (!stage2.alloc(F * bap_dim + 1) || !b.ap.alloc(F * H)) return 1;
    if (F > 0 && (!WB_CUDA(cudaMemcpyAsync(stage2.p, bap, F * bap_dim * sizeof(float), cudaMemcpyHostToDevice, c->stream)) ||
                  !bap_decode_run(stage2.p, (int)F, fft_size, bap_dim, b.ap.p)))
      return 1;
  }
  return WB_CUDA(cudaStreamSynchronize(c->stream)) ? 0 : 1;
}
int wb200_batch_compose_cmp(wb200_batch* h, const wb200_cmp_stream* streams, int n_streams) {
  ApiGuard api_guard;
  if (!ctx() || !wait_downloads(h, kDlCoded)) return 1;
  return batch_compose_cmp(&h->b, streams, n_streams) ? 0 : 1;
}
int wb200_batch_cmp_dim(const wb200_batch* h) { return h->b.cmp_dim; }
int wb200_batch_get_cmp(wb200_batch* h, float* out) {
  ApiGuard api_guard;
  return d2h(out, h->b.cmp.p, (size_t)h->b.total_frames * h->b.cmp_dim * sizeof(float));
}
// the same copy on the download stream (pinned `out`, valid after wb200_batch_wait_downloads / wb200_sync)
int wb200_batch_get_cmp_async(wb200_batch* h, float* out) {
  ApiGuard api_guard;
  Context* c = ctx();
  if (!c) return 1;
  Batch& b = h->b;
  if (!b.cmp.p || b.cmp_dim < 1) { set_error("cmp: wb200_batch_compose_cmp has not been run"); return 1; }
  if (!ensure_copy_streams(c)) return 1;
  const size_t n = (size_t)b.total_frames * b.cmp_dim;
  bool ok = WB_CUDA(cudaEventRecord(c->copy_event, c->stream)) && WB_CUDA(cudaStreamWaitEvent(c->copy_stream, c->copy_event, 0));
  if (ok && n > 0) ok = bulk_copy_async(out, b.cmp.p, n * sizeof(float), cudaMemcpyDeviceToHost, c->copy_stream);
  return ok && mark_downloads(h, kDlCoded) ? 0 : 1;
}
int wb200_batch_cmp_stats(wb200_batch* h, double* out) {
  ApiGuard api_guard;
  if (!ctx()) return 1;
  return batch_cmp_stats(&h->b, out) ? 0 : 1;
}
// data/scripts/addhtkheader.pl: pack("l", nframe) pack("l", 10000000 * frameshift / samprate)
// pack("s", byte) pack("s", type); host arithmetic only (12 bytes per utterance).
int wb200_htk_header(int n_frames, int samp_freq, int frame_shift, int byte_per_frame, int kind,
                     unsigned char* out12) {
  if (!out12 || samp_freq <= 0) return 1;
  const int32_t nf = n_frames;
  const int32_t shift = (int32_t)(10000000.0 * frame_shift / samp_freq);
  const int16_t bytes = (int16_t)byte_per_frame, type = (int16_t)kind;
  memcpy(out12, &nf, 4); memcpy(out12 + 4, &shift, 4); memcpy(out12 + 8, &bytes, 2); memcpy(out12 + 10, &type, 2);
  return 0;
}
int wb200_batch_gv_stats(wb200_batch* h, double* per_utt, double* partials) {
  ApiGuard api_guard;
  if (!ctx()) return 1;
  return batch_gv_stats(&h->b, per_utt, partials) ? 0 : 1;
}
int wb200_batch_lf0_stats(wb200_batch* h, double* out3);
int wb200_batch_feature_stats(wb200_batch* h, double* out) {
  ApiGuard api_guard;
  if (!ctx()) return 1;
  if (!batch_feature_stats(&h->b, out)) return 1;
  return wb200_batch_lf0_stats(h, out);                                 // row 0: voiced lf0
}
int wb200_batch_lf0_stats(wb200_batch* h, double* out3) {
  ApiGuard api_guard;
  Context* c = ctx();
  if (!c) return 1;
  DevBuf<double> d;
  if (!d.alloc(3)) return 1;
  if (!WB_CUDA(cudaMemsetAsync(d.p, 0, 3 * sizeof(double), c->stream))) return 1;
  if (h->b.total_frames > 0) {
    lf0_stats_kernel<<<148, 256, 0, c->stream>>>(h->b.f0.p, h->b.total_frames, d.p);
    WB_LAUNCH_CHECK();
  }
  return d2h(out3, d.p, 3 * sizeof(double));
}

}  // extern "C"
