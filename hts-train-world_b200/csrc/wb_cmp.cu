// world-b200: composing the training observation vectors ("cmp") on the device.
//
// Reference: data/Makefile.in:276-321 (the WORLD branch of the `cmp` target) — per utterance
//   perl scripts/window.pl DIM stream.f32 win1 win2 win3 > tmp.stream     (data/scripts/window.pl)
//   merge +f -s 0 -l ... -L ...  tmp.mgc < tmp.lf0 ... > tmp.cmp            (streams side by side)
//   perl scripts/addhtkheader.pl SAMPFREQ FRAMESHIFT BYTEPERFRAME 9 tmp.cmp (data/scripts/addhtkheader.pl)
// Here: one kernel writes every (frame, window, dimension) element of every stream straight
// into its column of the [total_frames][cmp_dim] float32 matrix, so the windowed streams never
// exist on their own.  The arithmetic follows window.pl literally: the float32 statics are
// widened to double, the taps are accumulated in double in the order k = -nlr .. nlr starting
// from 0.0, frame indices are clamped to the utterance, and an element becomes -1.0e10 (the
// "ignore value" of window.pl:55) when any tap that lies between the first and the last non-zero
// coefficient of the window reads -1.0e10.  The result is rounded to float32 once (pack "f").
// The kernel is a pure HBM stream: 4 bytes written per element, the statics come from L2.
#include "../../include/world_b200.h"
#include "wb_batch.h"

namespace wb {

namespace {

constexpr double kIgnoreValue = -1.0e+10;     // window.pl:55

// one (stream, window) pair = `dim` adjacent columns of the cmp frame
struct CmpSegment {
  const float* src;        // [total_frames][dim] statics (device)
  int dim, col0, nlr;
  int chk_lo, chk_hi;      // taps (0-based, 0 .. 2 nlr) between the first and last non-zero coefficient
  double coef[WB200_CMP_MAX_WIN_SIZE];
};

// Warps walk frames (grid-stride), lanes are the columns of the merged frame: a warp stores 32
// consecutive floats of one row and reads runs of consecutive statics, so both sides of this
// pure HBM stream are coalesced.  The (stream, window) segment of every column comes
// from a table in shared memory.
__global__ void __launch_bounds__(256)
cmp_compose_kernel(const CmpSegment* __restrict__ segs, int n_segs, const int* __restrict__ frame_utt,
                   const int* __restrict__ f_off, const int* __restrict__ f_len, int total_frames,
                   int cmp_dim, float* __restrict__ out) {
  extern __shared__ unsigned char smem_raw[];
  CmpSegment* sg = reinterpret_cast<CmpSegment*>(smem_raw);
  unsigned char* col_seg = smem_raw + (size_t)n_segs * sizeof(CmpSegment);
  for (int i = threadIdx.x; i < n_segs * (int)(sizeof(CmpSegment) / 4); i += blockDim.x)
    reinterpret_cast<int*>(sg)[i] = reinterpret_cast<const int*>(segs)[i];
  __syncthreads();
  for (int c = threadIdx.x; c < cmp_dim; c += blockDim.x) {
    int k = 0;
    while (k + 1 < n_segs && c >= sg[k + 1].col0) ++k;
    col_seg[c] = (unsigned char)k;
  }
  __syncthreads();
  // one warp per frame row: eight rows are in flight per CTA, which hides the dependent look-ups
  // (frame -> utterance -> frame range) behind the streams of the other warps
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int f = blockIdx.x * wpb + (threadIdx.x >> 5); f < total_frames; f += gridDim.x * wpb) {
    const int u = frame_utt[f];
    const int first = f_off[u], T = f_len[u], t = f - first;
#pragma unroll 4
    for (int c = lane; c < cmp_dim; c += 32) {
      const CmpSegment& s = sg[col_seg[c]];
      const int j = c - s.col0;
      double acc = 0.0;
      bool boundary = false;
      for (int k = -s.nlr; k <= s.nlr; ++k) {
        const int l = min(T - 1, max(0, t + k));                         // window.pl:95-103 / :111-119
        const double v = (double)s.src[(size_t)(first + l) * s.dim + j];
        const int tap = k + s.nlr;
        if (tap >= s.chk_lo && tap <= s.chk_hi && v == kIgnoreValue) boundary = true;
        acc = __dadd_rn(acc, __dmul_rn(s.coef[tap], v));                  // no contraction: perl adds a rounded product
      }
      out[(size_t)f * cmp_dim + c] = (float)(boundary ? kIgnoreValue : acc);
    }
  }
}

// per-column {count, sum, sum of squares} of a [frames][ndim] float matrix, read as one contiguous
// stream: the CTA has R * ndim threads and the grid stride is a multiple of ndim, so a thread always
// meets the same column; the R threads of a column are joined in shared memory.
__global__ void __launch_bounds__(1024)
cmp_stats_kernel(const float* __restrict__ m, long long n_elems, int ndim, double* __restrict__ out3) {
  extern __shared__ double sh[];                    // [3][blockDim.x]
  const int T = blockDim.x, tid = threadIdx.x;      // T % ndim == 0
  double c = 0.0, s = 0.0, q = 0.0;
  for (long long i = blockIdx.x * (long long)T + tid; i < n_elems; i += (long long)gridDim.x * T) {
    const double x = m[i];
    c += 1.0; s += x; q += x * x;
  }
  sh[tid] = c; sh[T + tid] = s; sh[2 * T + tid] = q;
  __syncthreads();
  if (tid < ndim) {
    double v[3] = {0.0, 0.0, 0.0};
    for (int t = tid; t < T; t += ndim) { v[0] += sh[t]; v[1] += sh[T + t]; v[2] += sh[2 * T + t]; }
    atomicAdd(&out3[tid * 3], v[0]); atomicAdd(&out3[tid * 3 + 1], v[1]); atomicAdd(&out3[tid * 3 + 2], v[2]);
  }
}

}  // namespace

bool batch_compose_cmp(Batch* b, const wb200_cmp_stream* streams, int n_streams) {
  Context* c = ctx();
  if (!c) return false;
  if (n_streams < 1 || n_streams > WB200_CMP_MAX_STREAMS) { set_error("cmp: %d streams (1..%d supported)", n_streams, WB200_CMP_MAX_STREAMS); return false; }
  const int F = b->total_frames;
  std::vector<CmpSegment> segs;
  std::vector<DevBuf<float>> staged(n_streams);
  int col = 0;
  for (int s = 0; s < n_streams; ++s) {
    const wb200_cmp_stream& st = streams[s];
    const float* src = nullptr;
    int dim = st.dim;
    switch (st.source) {
      case WB200_CMP_SRC_MGC: src = b->mgc.p; dim = b->mgc_dim; break;
      case WB200_CMP_SRC_LF0: src = b->lf0.p; dim = 1; break;
      case WB200_CMP_SRC_BAP: src = b->bap.p; dim = b->bap_dim; break;
      case WB200_CMP_SRC_HOST:
        if (!st.host_data || dim < 1) { set_error("cmp: stream %d has no data", s); return false; }
        if (!staged[s].alloc((size_t)F * dim + 1)) return false;
        if (F > 0) WB_CUDA_OR_RETURN(cudaMemcpyAsync(staged[s].p, st.host_data, (size_t)F * dim * sizeof(float), cudaMemcpyHostToDevice, c->stream), false);
        src = staged[s].p;
        break;
      default: set_error("cmp: stream %d: unknown source %d", s, st.source); return false;
    }
    if (!src || dim < 1) { set_error("cmp: stream %d: features have not been coded (wb200_batch_code)", s); return false; }
    if (st.n_win < 1 || st.n_win > WB200_CMP_MAX_WINDOWS) { set_error("cmp: stream %d: %d windows (1..%d supported)", s, st.n_win, WB200_CMP_MAX_WINDOWS); return false; }
    for (int w = 0; w < st.n_win; ++w) {
      const int size = st.win_size[w];
      if (size < 1 || size > WB200_CMP_MAX_WIN_SIZE || size % 2 != 1) {      // window.pl:83-85 dies on even sizes
        set_error("cmp: stream %d window %d: size %d (must be odd, <= %d)", s, w, size, WB200_CMP_MAX_WIN_SIZE);
        return false;
      }
      CmpSegment g;
      g.src = src; g.dim = dim; g.col0 = col; g.nlr = (size - 1) / 2;
      for (int i = 0; i < WB200_CMP_MAX_WIN_SIZE; ++i) g.coef[i] = i < size ? st.win_coef[w][i] : 0.0;
      g.chk_lo = 0; g.chk_hi = size - 1;                                     // window.pl:70-81
      while (g.chk_lo < size && g.coef[g.chk_lo] == 0.0) ++g.chk_lo;
      while (g.chk_hi >= 0 && g.coef[g.chk_hi] == 0.0) --g.chk_hi;
      segs.push_back(g);
      col += dim;
    }
  }
  b->cmp_dim = col;
  if (!b->cmp.alloc((size_t)F * col + 1)) return false;
  if (F == 0) return true;
  DevBuf<CmpSegment> d_segs;
  if (!d_segs.alloc(segs.size())) return false;
  WB_CUDA_OR_RETURN(cudaMemcpyAsync(d_segs.p, segs.data(), segs.size() * sizeof(CmpSegment), cudaMemcpyHostToDevice, c->stream), false);
  {
    KernelTimer kt("cmp_compose_kernel");
    const size_t smem = segs.size() * sizeof(CmpSegment) + (size_t)col + 16;
    cmp_compose_kernel<<<c->sm_count * 8, 256, smem, c->stream>>>(d_segs.p, (int)segs.size(), b->frame_utt.p, b->f_off.p, b->f_len.p, F, col, b->cmp.p);
    WB_LAUNCH_CHECK(); kt.stop();
  }
  // the host vectors (segs, staged uploads) must outlive the copies queued above
  WB_CUDA_OR_RETURN(cudaStreamSynchronize(c->stream), false);
  return true;
}

bool batch_cmp_stats(Batch* b, double* h_out) {   // [cmp_dim][3]
  Context* c = ctx();
  if (!c) return false;
  if (!b->cmp.p || b->cmp_dim < 1) { set_error("cmp stats: wb200_batch_compose_cmp has not been run"); return false; }
  const int F = b->total_frames, nd = b->cmp_dim;
  DevBuf<double> d;
  if (!d.alloc((size_t)nd * 3)) return false;
  cudaStream_t st = c->stream;
  if (!dev_fill(d.p, 0, (size_t)nd * 3 * sizeof(double))) return false;
  if (F > 0) {
    if (nd > 1024) { set_error("cmp stats: %d columns (<= 1024 supported)", nd); return false; }
    const int threads = nd * (256 / nd > 0 ? 256 / nd : 1);
    cmp_stats_kernel<<<c->sm_count * 4, threads, 3 * threads * sizeof(double), st>>>(b->cmp.p, (long long)F * nd, nd, d.p);
    WB_LAUNCH_CHECK();
  }
  if (!read_back(h_out, d.p, (size_t)nd * 3 * sizeof(double))) return false;
  return true;
}

}  // namespace wb
