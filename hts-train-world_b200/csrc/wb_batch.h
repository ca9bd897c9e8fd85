// world-b200: host-side description of a batch of utterances resident in HBM and the
// per-stage launchers implemented in the wb_*.cu files.
#pragma once
#include <vector>
#include "wb_common.cuh"

namespace wb {

// Device view of the utterance table (all pointers are device pointers).
struct UttView {
  const double* x;          // concatenated samples of all utterances
  const long long* x_off;   // [n_utt] first sample of utterance u in x
  const int* x_len;         // [n_utt]
  const int* f_off;         // [n_utt] first frame of utterance u in the frame table
  const int* f_len;         // [n_utt]
  int n_utt;
  int max_f_len = 0;        // host-side knowledge: the largest f_len (0 = unknown); sizes the randn table without a read-back
};

struct Batch {
  int fs = 0;
  int n_utt = 0;
  double frame_period = 5.0;           // ms
  std::vector<long long> h_x_off;
  std::vector<int> h_x_len, h_f_off, h_f_len;
  long long total_samples = 0;
  int total_frames = 0;
  int max_x_len = 0, max_f_len = 0;

  DevBuf<double> x;
  DevBuf<long long> x_off;
  DevBuf<int> x_len, f_off, f_len;
  DevBuf<int> frame_utt;               // [total_frames]
  DevBuf<double> frame_t;              // [total_frames] seconds
  DevBuf<double> f0_raw, f0;           // [total_frames]
  int fft_size = 0;                    // CheapTrick / Synthesis size
  DevBuf<double> sp, ap;               // [total_frames][fft_size/2+1]
  // coded features of the analysis tool (W/test/analysis.cpp:293-390), float32 like its files
  int mgc_dim = 0, bap_dim = 0;
  DevBuf<float> lf0, mgc, bap;         // [total_frames], [total_frames][mgc_dim], [total_frames][bap_dim]
  // training observation vectors (data/Makefile.in:276-321): statics + delta windows of every stream
  int cmp_dim = 0;
  DevBuf<float> cmp;                   // [total_frames][cmp_dim]
  // synthesis
  std::vector<long long> h_y_off;
  std::vector<int> h_y_len;
  long long total_y = 0;
  DevBuf<long long> y_off;
  DevBuf<int> y_len;
  DevBuf<double> y;

  UttView view() const {
    UttView v;
    v.x = x.p; v.x_off = x_off.p; v.x_len = x_len.p; v.f_off = f_off.p; v.f_len = f_len.p;
    v.n_utt = n_utt;
    v.max_f_len = 0;
    for (int f : h_f_len) v.max_f_len = f > v.max_f_len ? f : v.max_f_len;
    return v;
  }
};

// Lay out n_utt utterances; f_len[u] frames each (may be 0 = "decide later with set_frames").
bool batch_layout(Batch* b, int fs, double frame_period, int n_utt, const int* x_len,
                  const int* f_len);
// Fill frame_utt and frame_t = i * frame_period / 1000 (W/src/dio.cpp:605-606).
bool batch_default_frames(Batch* b);

// ---- stage launchers (all asynchronous on ctx()->stream unless noted) -----------------------
struct DioParams { double f0_floor, f0_ceil, channels_in_octave, frame_period; int speed; double allowed_range; };
bool dio_run(Batch* b, const DioParams& p, double* d_f0_out);
bool stonemask_run(const UttView& u, int fs, int total_frames, const int* frame_utt,
                   const double* frame_t, const double* f0_in, double* f0_out);
bool cheaptrick_run(const UttView& u, int fs, int total_frames, const int* frame_utt,
                    const double* frame_t, const double* f0, int fft_size, double q1,
                    double* sp);
bool d4c_run(const UttView& u, int fs, int total_frames, const int* frame_utt,
             const double* frame_t, const double* f0, int fft_size, double threshold,
             double* ap);
// f0/sp/ap in the batch -> b->y (allocated here); y_len per utterance given by caller.
bool synthesis_run(Batch* b, const int* y_len);
struct HarvestParams { double f0_floor, f0_ceil, frame_period; };
bool harvest_run(Batch* b, const HarvestParams& p, double* d_f0_out);

// mel-DCT codec (wb_codec.cu): rows [n_frames][fft_size/2+1] <-> coded [n_frames][ndim], device pointers
bool codec_encode_run(const double* d_rows, int n_frames, int fs, int fft_size, int ndim, double scale,
                      double zero_floor, double c0_add, double* d_out, bool f32log = false);
bool codec_decode_run(const double* d_coded, int n_frames, int fs, int fft_size, int ndim, double* d_rows);
bool batch_code_features(Batch* b, int mgc_dim, int bap_dim);
bool bap_decode_run(const float* d_bap, int n_frames, int fft_size, int bap_dim, double* d_rows);   // W/test/synth.cpp:221-247
bool lf0_to_f0_run(const float* d_lf0, int n, double* d_f0);                                          // ToF0, W/test/synth.cpp:81-89
bool batch_feature_stats(Batch* b, double* h_out);
bool batch_gv_stats(Batch* b, double* h_per_utt, double* h_partials);

// zero-phase IIR decimation of every utterance (wb_harvest.cu), used by Dio when option.speed > 1
bool decimate_run(const Batch* b, int r, const std::vector<int>& want_len, DevBuf<double>* y,
                  DevBuf<long long>* y_off, DevBuf<int>* y_len, std::vector<int>* out_len);

// a[3], b[2] of the 3rd-order decimation filter for ratio r = 2..12 (W/src/matlabfunctions.cpp:29-112)
bool decimate_filter_coefficients(int r, double* a, double* b);

// generic helper: exclusive prefix sum of counts within each utterance's frame range.
// out[f] = sum of counts[g] for g in [f_off[u], f); totals[u] = sum over the utterance.
bool segmented_exclusive_scan(const long long* counts, const int* f_off, const int* f_len,
                              int n_utt, long long* out, long long* totals);

// per-stage device time of the last call (ms), measured with CUDA events
struct StageTimes { float dio = 0, stonemask = 0, cheaptrick = 0, d4c = 0, synthesis = 0, harvest = 0; };
extern StageTimes g_times;

}  // namespace wb
