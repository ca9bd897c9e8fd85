/* world-b200 — batched extension API of libworld_b200.so (plain C ABI).
 *
 * The drop-in WORLD entry points (include/world/{dio,stonemask,cheaptrick,d4c,synthesis,harvest}.h)
 * keep the reference's one-utterance-per-call signatures.  The reference pipeline forks one
 * `analysis` process per utterance (data/Makefile.in:125,214); that loop is what this API
 * replaces: many utterances are laid out once in HBM and every stage runs as kernels batched
 * over all frames (or pulses) of all utterances.  f0 / spectrogram / aperiodicity stay on
 * the device between stages.
 *
 * All functions return 0 on success and non-zero on failure; wb200_last_error() then holds a
 * message.  There is no CPU fallback: without a CUDA device every call fails.
 *
 * Threads: every entry point (these and the drop-in WORLD functions) may be called from any thread; the
 * library has one device context and serialises the calls internally, binding the calling thread to the
 * context's device.  Asynchronous copies queued by one call stay valid across calls of other threads.
 */
#ifndef WORLD_B200_H_
#define WORLD_B200_H_
#include <stdint.h>
#include "world/cheaptrick.h"
#include "world/codec.h"
#include "world/d4c.h"
#include "world/dio.h"
#include "world/harvest.h"
#include "world/stonemask.h"
#include "world/synthesis.h"

WORLD_BEGIN_C_DECLS

typedef struct wb200_batch wb200_batch;

/* error channel for the void-returning WORLD API */
WORLD_API const char *wb200_last_error(void);
/* bind the calling thread's CUDA device and build the context (twiddles, stream). */
WORLD_API int wb200_init(int device);
/* number of kernels this library has launched so far */
WORLD_API unsigned long long wb200_launch_count(void);
/* device time (ms, CUDA events on the library stream) of the last call of each stage:
 * out[0..5] = dio, stonemask, cheaptrick, d4c, synthesis, harvest */
WORLD_API void wb200_stage_times(float *out6);
/* run all library work on a caller-owned CUDA stream (cudaStream_t); the caller keeps it alive */
WORLD_API int wb200_set_stream(void *stream);
/* per-kernel device timing (CUDA events on the library stream around each launch of the named
 * kernel).  Off by default.  wb200_kernel_time returns 0 and the accumulated milliseconds and
 * launch count since the last reset, 1 if the kernel never ran. */
WORLD_API void wb200_kernel_timing(int on);
WORLD_API void wb200_kernel_times_reset(void);
WORLD_API int wb200_kernel_time(const char *name, double *ms_total, long long *launches);
/* measured CUDA-core FMA peak in TFLOP/s (fp64 != 0: double, else float) — the roofline
 * denominator of SURVEY.md 8(d) */
WORLD_API double wb200_measure_fma_peak(int fp64);
/* state (0 / 1) of a run-time switch of the library: "lovetrain_fp32", "stonemask_dft", "dio_fused",
 * "harvest_fused", "harvest_refine_thread" (environment WB_D4C_LT32, WB_STONEMASK_DFT, WB_DIO_FUSED,
 * WB_HARVEST_FUSED, WB_HARVEST_REFINE_THREAD = 0 | 1 override the defaults); unknown names give 0 */
WORLD_API int wb200_option(const char *name);
/* The library keeps its scratch memory (per-stage buffers, gigabytes for a corpus-sized batch) in a CUDA
 * memory pool of its own and never returns it on its own; wb200_trim() waits for the library stream and
 * hands everything that is not in use back to the driver (e.g. before a training step that needs the HBM). */
WORLD_API int wb200_trim(void);
/* first `n` values of the randn table as doubles (test hook: must equal the reference's
 * randn() stream after randn_reseed(), W/src/matlabfunctions.cpp:247-277) */
WORLD_API int wb200_randn_stream(double *out, long long n);

/* ---- batch life cycle ----------------------------------------------------------------- */
/* x_lengths[n_utt] samples per utterance.  Frames per utterance follow GetSamplesForDIO
 * (W/src/dio.cpp:638-640) and temporal positions are i * frame_period / 1000. */
WORLD_API wb200_batch *wb200_batch_create(int fs, double frame_period, int n_utt,
                                          const int *x_lengths);
WORLD_API void wb200_batch_destroy(wb200_batch *b);
WORLD_API int wb200_batch_total_frames(const wb200_batch *b);
WORLD_API long long wb200_batch_total_samples(const wb200_batch *b);
/* f_off[n_utt], f_len[n_utt]: where each utterance's frames sit in the flat frame table */
WORLD_API int wb200_batch_frame_layout(const wb200_batch *b, int *f_off, int *f_len);

/* inputs: utterances back to back (no padding) in host memory (pinned memory makes the copy
 * asynchronous) or already in device memory. pcm16 is converted as pcm / 32768.0, the
 * convention of the reference's wavread (W/test/audioio.cpp:229-251). */
WORLD_API int wb200_batch_upload_pcm16(wb200_batch *b, const int16_t *host_pcm);
WORLD_API int wb200_batch_upload_f64(wb200_batch *b, const double *host_x);
WORLD_API int wb200_batch_set_pcm16_device(wb200_batch *b, const int16_t *dev_pcm);
/* asynchronous upload on a second stream: overlaps the computation of another batch; the stages
 * of this batch wait for it on the device.  host_pcm must be pinned. */
WORLD_API int wb200_batch_upload_pcm16_async(wb200_batch *b, const int16_t *host_pcm);

/* ---- stages (each = the reference function of the same name over the whole batch) -------- */
WORLD_API int wb200_batch_dio(wb200_batch *b, const DioOption *option);        /* -> raw f0 */
WORLD_API int wb200_batch_stonemask(wb200_batch *b);                           /* raw f0 -> f0 */
WORLD_API int wb200_batch_harvest(wb200_batch *b, const HarvestOption *option);/* -> f0 */
WORLD_API int wb200_batch_cheaptrick(wb200_batch *b, const CheapTrickOption *option);
WORLD_API int wb200_batch_d4c(wb200_batch *b, int fft_size, const D4COption *option);
/* y_lengths may be NULL: int((f0_length-1) * frame_period / 1000 * fs) + 1 (W/test/synth.cpp:259) */
WORLD_API int wb200_batch_synthesis(wb200_batch *b, const int *y_lengths);

/* ---- results / injection of upstream values (flat over the frame table) ------------------- */
WORLD_API int wb200_batch_get_f0(wb200_batch *b, double *host_f0, int refined);
WORLD_API int wb200_batch_set_f0(wb200_batch *b, const double *host_f0, int refined);
WORLD_API int wb200_batch_get_sp(wb200_batch *b, double *host_sp);   /* [frames][fft/2+1] */
WORLD_API int wb200_batch_get_ap(wb200_batch *b, double *host_ap);
WORLD_API int wb200_batch_set_sp_ap(wb200_batch *b, int fft_size, const double *host_sp,
                                    const double *host_ap);
/* the synth tool's raw parameter files (W/test/synth.cpp:160-190, spec_dimension == 0): float32
 * f0 [frames] and sp / ap [frames][fft_size/2+1], widened to double on the device (ToDouble).
 * The host buffers may be pinned; the entry of Synthesis-only runs (BASELINE config 4). */
WORLD_API int wb200_batch_set_params_f32(wb200_batch *b, int fft_size, const float *host_f0,
                                         const float *host_sp, const float *host_ap);
WORLD_API long long wb200_batch_total_y(const wb200_batch *b);
WORLD_API int wb200_batch_y_layout(const wb200_batch *b, long long *y_off, int *y_len);
WORLD_API int wb200_batch_get_y(wb200_batch *b, double *host_y);     /* back to back */
/* 16-bit output as the reference's wavwrite: trunc(y * 32767) clamped (W/test/audioio.cpp:115-170) */
WORLD_API int wb200_batch_get_y_pcm16(wb200_batch *b, int16_t *host_pcm);
/* asynchronous variant on a download stream (pinned `host_pcm`, valid after wb200_sync()) */
WORLD_API int wb200_batch_get_y_pcm16_async(wb200_batch *b, int16_t *host_pcm);
/* block the calling thread until the asynchronous result copies queued for THIS batch (get_coded_async,
 * get_y_pcm16_async) have landed in their host buffers; work queued for other batches keeps running */
WORLD_API int wb200_batch_wait_downloads(wb200_batch *b);
/* one utterance's slice of the results (any pointer may be NULL): f0_raw / f0 [f_len], sp / ap
 * [f_len][fft_size/2+1], y [y_len] -- spot checks of a corpus-sized batch without copying all of it */
WORLD_API int wb200_batch_get_utterance(wb200_batch *b, int utt, double *host_f0_raw, double *host_f0,
                                        double *host_sp, double *host_ap, double *host_y);
/* device pointers for zero-copy consumers: which = "x","f0_raw","f0","sp","ap","y" */
WORLD_API void *wb200_batch_device_ptr(wb200_batch *b, const char *which);
/* corpus statistics of voiced log-f0 over the batch: out = {count, sum, sum of squares};
 * the per-GPU partials that the NCCL all-reduce combines */
WORLD_API int wb200_batch_lf0_stats(wb200_batch *b, double *out3);
/* ---- the analysis tool's coded outputs (W/test/analysis.cpp:293-390) for the whole batch ------
 * lf0 = log f0 (0 stays 0); mgc = CodeSpectralEnvelope(sp * 1e4, mgc_dim) with c0 + 12;
 * bap = CodeSpectralEnvelope(ap * 1e4, bap_dim) with c0 - 9.21034; all float32 like the tool's files.
 * Needs CheapTrick and D4C results in the batch. */
WORLD_API int wb200_batch_code(wb200_batch *b, int mgc_dim, int bap_dim);
WORLD_API int wb200_batch_get_coded(wb200_batch *b, float *host_lf0, float *host_mgc, float *host_bap);
/* the same copies queued on a second stream so that they overlap the stages launched afterwards
 * (Synthesis); the (pinned) host buffers are valid after wb200_sync() */
WORLD_API int wb200_batch_get_coded_async(wb200_batch *b, float *host_lf0, float *host_mgc, float *host_bap);
/* decode mgc back into the batch's spectrogram (DecodeSpectralEnvelope, then the inverse of the
 * tool's scalings): the entry of Synthesis-only runs that start from float32 mgc files */
WORLD_API int wb200_batch_decode_mgc(wb200_batch *b, int fft_size, int mgc_dim, const float *host_mgc);
/* the coded branch of the synth tool (W/test/synth.cpp:171-247, spec_dimension != 0) for the whole batch: float32
 * lf0 [frames], mgc [frames][mgc_dim], bap [frames][bap_dim] as the analysis tool writes them -> f0 = exp(lf0) (0
 * stays 0), spectrogram = DecodeSpectralEnvelope(mgc with c0 - 12) / 1e4, aperiodicity = exp(mgc2sp(bap with
 * c0 + 9.210340, order bap_dim or bap_dim - 1 when odd, alpha 0.55, gamma 0)) / 1e4 (W/test/sptkfunctions.cpp
 * :186-275).  The reference fills only the first `order` bins of an aperiodicity row and, for an even bap_dim,
 * reads one coefficient past its buffer; here every bin is defined by the same formula and the missing
 * coefficient is 0.  Any pointer may be NULL.  The entry of Synthesis-only runs from coded files (config 4). */
WORLD_API int wb200_batch_set_coded_f32(wb200_batch *b, int fft_size, int mgc_dim, int bap_dim,
                                        const float *host_lf0, const float *host_mgc, const float *host_bap);
/* per-GPU partials of the corpus statistics: out[(1 + mgc_dim)][3] = {count, sum, sum of squares}
 * of voiced lf0 (row 0) and of every mgc dimension over all frames (rows 1..mgc_dim); the NCCL
 * all-reduce of these rows gives the corpus mean / variance */
WORLD_API int wb200_batch_feature_stats(wb200_batch *b, double *out);
/* global-variance statistics (scripts/Training.pl make_data_gv :1402-1456; data/Makefile.in:447-458): for every
 * utterance the per-dimension variance over its frames of the static streams [mgc | lf0 (voiced frames) | bap],
 * computed as SPTK `vstat -d -o 2` does (double sums in frame order, E[x^2] - mean^2); NaN where a stream has
 * no frame.  per_utt[n_utt][mgc_dim + 1 + bap_dim] (may be NULL); partials[mgc_dim + 1 + bap_dim][3] = {count,
 * sum, sum of squares} over this batch's utterances of those variances rounded to float32 (the reference's
 * tmp.var1 is a float file) -- summed over batches / GPUs they give the mean and variance of the variances
 * (stats/gv.var).  Needs wb200_batch_code. */
WORLD_API int wb200_batch_gv_stats(wb200_batch *b, double *per_utt, double *partials);
/* ---- training observation vectors (data/Makefile.in:276-321, the `cmp` target) ----------------
 * Every stream (mgc, lf0, bap, and any stream computed on the host such as the two-dimensional
 * lf0 and vib of data/scripts/Extract.py) is extended by its delta windows exactly as
 * data/scripts/window.pl does (double accumulation, frames clamped to the utterance, -1.0e10 as
 * the ignore value) and written side by side, in the order given, into one float32 matrix
 * [total_frames][cmp_dim] — the `merge` chain of the Makefile.  Window files data/win/*.win[123]
 * hold "size c1 c2 ... csize"; pass those numbers in win_size / win_coef. */
#define WB200_CMP_MAX_STREAMS 8
#define WB200_CMP_MAX_WINDOWS 4
#define WB200_CMP_MAX_WIN_SIZE 15
enum { WB200_CMP_SRC_HOST = 0, WB200_CMP_SRC_MGC = 1, WB200_CMP_SRC_LF0 = 2, WB200_CMP_SRC_BAP = 3 };
typedef struct {
  int source;                 /* WB200_CMP_SRC_*: host_data, or the batch's own coded features */
  int dim;                    /* static dimensionality (ignored for the batch's own features) */
  const float *host_data;     /* [total_frames][dim] statics, utterances back to back (SRC_HOST) */
  int n_win;                  /* NMGCWIN / NLF0WIN / NBAPWIN / NVIBWIN */
  int win_size[WB200_CMP_MAX_WINDOWS];
  double win_coef[WB200_CMP_MAX_WINDOWS][WB200_CMP_MAX_WIN_SIZE];
} wb200_cmp_stream;
WORLD_API int wb200_batch_compose_cmp(wb200_batch *b, const wb200_cmp_stream *streams, int n_streams);
WORLD_API int wb200_batch_cmp_dim(const wb200_batch *b);           /* floats per frame */
WORLD_API int wb200_batch_get_cmp(wb200_batch *b, float *host_cmp); /* [total_frames][cmp_dim] */
/* the same copy on the download stream (pinned host_cmp; valid after wb200_batch_wait_downloads / wb200_sync) */
WORLD_API int wb200_batch_get_cmp_async(wb200_batch *b, float *host_cmp);
/* per-GPU partials {count, sum, sum of squares} of every cmp column: out[cmp_dim][3] */
WORLD_API int wb200_batch_cmp_stats(wb200_batch *b, double *out);
/* the 12-byte HTK header of data/scripts/addhtkheader.pl: int32 nframe, int32 frame shift in
 * 100 ns units (10000000 * frame_shift / samp_freq, truncated), int16 bytes per frame, int16
 * parameter kind (9 = USER), native byte order */
WORLD_API int wb200_htk_header(int n_frames, int samp_freq, int frame_shift, int byte_per_frame,
                               int kind, unsigned char *out12);
/* block until everything queued on the library stream has finished */
/* Deferred bulk copies (0 / 1, default 0).  With deferral on, wb200_batch_upload_pcm16_async, wb200_batch_get_coded_async
 * and wb200_batch_get_y_pcm16_async only record the request; the copies are issued at the next safe point -- the start
 * of StoneMask / CheapTrick / D4C's main kernel, the stretch of a pass without host interaction -- or at the latest when something
 * waits for them (wb200_batch_wait_downloads, wb200_sync, the next pass over the same batch, the stage that needs the
 * uploaded samples).  A bulk copy in flight delays the small read-backs of the stage that runs beside it by its whole
 * PCIe time (measured); a pipelined caller switches this on.  Host buffers must stay valid until the copies are done. */
WORLD_API int wb200_set_copy_deferral(int on);
WORLD_API int wb200_sync(void);

WORLD_END_C_DECLS
#endif
