"""world-b200: B200-native WORLD analysis/synthesis (Dio, StoneMask, Harvest, CheapTrick, D4C,
Synthesis) behind the reference's C API.  This module is the Python host-side mirror of that
API: numpy in / numpy out wrappers over the C ABI of libworld_b200.so, plus the batched
`Corpus` handle that replaces the reference's one-process-per-utterance loop
(data/Makefile.in:125-242).

There is deliberately no CPU path: if the shared library is missing or no CUDA device is
visible, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libworld_b200.so")
if os.environ.get("WB200_LIB"):          # an experimental build of the same sources (build.py --variant): A/B measurements
    LIB_PATH = os.path.join(_HERE, os.environ["WB200_LIB"]) if not os.path.isabs(os.environ["WB200_LIB"]) else os.environ["WB200_LIB"]


class DioOption(C.Structure):          # W/src/world/dio.h:16-23
    _fields_ = [("f0_floor", C.c_double), ("f0_ceil", C.c_double),
                ("channels_in_octave", C.c_double), ("frame_period", C.c_double),
                ("speed", C.c_int), ("allowed_range", C.c_double)]


class CheapTrickOption(C.Structure):   # W/src/world/cheaptrick.h:16-20
    _fields_ = [("q1", C.c_double), ("f0_floor", C.c_double), ("fft_size", C.c_int)]


class D4COption(C.Structure):          # W/src/world/d4c.h:16-18
    _fields_ = [("threshold", C.c_double)]


class HarvestOption(C.Structure):      # W/src/world/harvest.h:16-20
    _fields_ = [("f0_floor", C.c_double), ("f0_ceil", C.c_double), ("frame_period", C.c_double)]


_dp = C.POINTER(C.c_double)
_dpp = C.POINTER(_dp)
_ip = C.POINTER(C.c_int)
_lib = None


CMP_MAX_STREAMS, CMP_MAX_WINDOWS, CMP_MAX_WIN_SIZE = 8, 4, 15
CMP_SRC = {"host": 0, "mgc": 1, "lf0": 2, "bap": 3}


class CmpStream(C.Structure):          # include/world_b200.h: wb200_cmp_stream
    _fields_ = [("source", C.c_int), ("dim", C.c_int), ("host_data", C.c_void_p), ("n_win", C.c_int),
                ("win_size", C.c_int * CMP_MAX_WINDOWS),
                ("win_coef", (C.c_double * CMP_MAX_WIN_SIZE) * CMP_MAX_WINDOWS)]


# data/win/{mgc,lf0,bap,vib}.win[123] of the reference: static, delta, delta-delta
DEFAULT_WINDOWS = ((1.0,), (-0.5, 0.0, 0.5), (1.0, -2.0, 1.0))


def parse_window_file(text):
    """One data/win/*.win file: "size c1 ... csize" on its first line (data/scripts/window.pl:63-66)."""
    tok = text.split("\n")[0].split()
    size = int(float(tok[0]))
    return tuple(float(v) for v in tok[1:1 + size])


def htk_header(n_frames, samp_freq, frame_shift, byte_per_frame, kind=9):
    """The 12 bytes data/scripts/addhtkheader.pl puts in front of a cmp file."""
    out = (C.c_ubyte * 12)()
    _check(lib().wb200_htk_header(int(n_frames), int(samp_freq), int(frame_shift), int(byte_per_frame), int(kind), out),
           "htk_header")
    return bytes(out)


class WorldB200Error(RuntimeError):
    pass


def _ptr(a):
    return a.ctypes.data_as(_dp)


def _rows(a2d):
    n = a2d.shape[0]
    arr = (_dp * n)()
    base, stride = a2d.ctypes.data, a2d.strides[0]
    for i in range(n):
        arr[i] = C.cast(base + i * stride, _dp)
    return arr


def lib():
    """The loaded C-ABI library.  Raises loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WorldB200Error(
            "%s not found: run `python hts-train-world_b200/build.py` (nvcc, sm_100a). "
            "There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL | os.RTLD_NOW)
    vp, i32, i64, f64 = C.c_void_p, C.c_int, C.c_longlong, C.c_double
    sigs = {
        # WORLD API
        "Dio": (None, [_dp, i32, i32, C.POINTER(DioOption), _dp, _dp]),
        "InitializeDioOption": (None, [C.POINTER(DioOption)]),
        "GetSamplesForDIO": (i32, [i32, i32, f64]),
        "StoneMask": (None, [_dp, i32, i32, _dp, _dp, i32, _dp]),
        "CheapTrick": (None, [_dp, i32, i32, _dp, _dp, i32, C.POINTER(CheapTrickOption), _dpp]),
        "InitializeCheapTrickOption": (None, [i32, C.POINTER(CheapTrickOption)]),
        "GetFFTSizeForCheapTrick": (i32, [i32, C.POINTER(CheapTrickOption)]),
        "GetF0FloorForCheapTrick": (f64, [i32, i32]),
        "D4C": (None, [_dp, i32, i32, _dp, _dp, i32, i32, C.POINTER(D4COption), _dpp]),
        "InitializeD4COption": (None, [C.POINTER(D4COption)]),
        "Synthesis": (None, [_dp, i32, _dpp, _dpp, i32, f64, i32, i32, _dp]),
        "Harvest": (None, [_dp, i32, i32, C.POINTER(HarvestOption), _dp, _dp]),
        "InitializeHarvestOption": (None, [C.POINTER(HarvestOption)]),
        "GetSamplesForHarvest": (i32, [i32, i32, f64]),
        # extension API (include/world_b200.h)
        "wb200_last_error": (C.c_char_p, []),
        "wb200_init": (i32, [i32]),
        "wb200_launch_count": (C.c_ulonglong, []),
        "wb200_stage_times": (None, [C.POINTER(C.c_float)]),
        "wb200_randn_stream": (i32, [_dp, i64]),
        "wb200_set_stream": (i32, [vp]),
        "wb200_kernel_timing": (None, [i32]),
        "wb200_kernel_times_reset": (None, []),
        "wb200_kernel_time": (i32, [C.c_char_p, _dp, C.POINTER(i64)]),
        "wb200_measure_fma_peak": (f64, [i32]),
        "wb200_option": (i32, [C.c_char_p]),
        "wb200_trim": (i32, []),
        "wb200_batch_create": (vp, [i32, f64, i32, _ip]),
        "wb200_batch_destroy": (None, [vp]),
        "wb200_batch_total_frames": (i32, [vp]),
        "wb200_batch_total_samples": (i64, [vp]),
        "wb200_batch_frame_layout": (i32, [vp, _ip, _ip]),
        "wb200_batch_upload_pcm16": (i32, [vp, vp]),
        "wb200_batch_upload_pcm16_async": (i32, [vp, vp]),
        "wb200_batch_get_y_pcm16_async": (i32, [vp, vp]),
        "wb200_batch_upload_f64": (i32, [vp, vp]),
        "wb200_batch_set_pcm16_device": (i32, [vp, vp]),
        "wb200_batch_dio": (i32, [vp, C.POINTER(DioOption)]),
        "wb200_batch_stonemask": (i32, [vp]),
        "wb200_batch_harvest": (i32, [vp, C.POINTER(HarvestOption)]),
        "wb200_batch_cheaptrick": (i32, [vp, C.POINTER(CheapTrickOption)]),
        "wb200_batch_d4c": (i32, [vp, i32, C.POINTER(D4COption)]),
        "wb200_batch_synthesis": (i32, [vp, _ip]),
        "wb200_batch_get_f0": (i32, [vp, _dp, i32]),
        "wb200_batch_set_f0": (i32, [vp, _dp, i32]),
        "wb200_batch_get_sp": (i32, [vp, vp]),
        "wb200_batch_get_ap": (i32, [vp, vp]),
        "wb200_batch_set_sp_ap": (i32, [vp, i32, _dp, _dp]),
        "wb200_batch_set_params_f32": (i32, [vp, i32, vp, vp, vp]),
        "wb200_batch_total_y": (i64, [vp]),
        "wb200_batch_y_layout": (i32, [vp, C.POINTER(i64), _ip]),
        "wb200_batch_get_y": (i32, [vp, vp]),
        "wb200_batch_get_y_pcm16": (i32, [vp, vp]),
        "wb200_batch_get_utterance": (i32, [vp, i32, vp, vp, vp, vp, vp]),
        "wb200_batch_wait_downloads": (i32, [vp]),
        "wb200_batch_device_ptr": (vp, [vp, C.c_char_p]),
        "wb200_batch_lf0_stats": (i32, [vp, _dp]),
        "wb200_batch_code": (i32, [vp, i32, i32]),
        "wb200_batch_get_coded": (i32, [vp, vp, vp, vp]),
        "wb200_batch_get_coded_async": (i32, [vp, vp, vp, vp]),
        "wb200_batch_decode_mgc": (i32, [vp, i32, i32, vp]),
        "wb200_batch_set_coded_f32": (i32, [vp, i32, i32, i32, vp, vp, vp]),
        "wb200_batch_feature_stats": (i32, [vp, _dp]),
        "wb200_batch_gv_stats": (i32, [vp, vp, vp]),
        "wb200_batch_compose_cmp": (i32, [vp, C.POINTER(CmpStream), i32]),
        "wb200_batch_cmp_dim": (i32, [vp]),
        "wb200_batch_get_cmp": (i32, [vp, vp]),
        "wb200_batch_get_cmp_async": (i32, [vp, vp]),
        "wb200_batch_cmp_stats": (i32, [vp, _dp]),
        "wb200_htk_header": (i32, [i32, i32, i32, i32, i32, C.POINTER(C.c_ubyte)]),
        "GetNumberOfAperiodicities": (i32, [i32]),
        "CodeSpectralEnvelope": (None, [_dpp, i32, i32, i32, i32, _dpp]),
        "DecodeSpectralEnvelope": (None, [_dpp, i32, i32, i32, i32, _dpp]),
        "wb200_sync": (i32, []),
        "wb200_set_copy_deferral": (i32, [i32]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


EXPORTED_SYMBOLS = None  # filled lazily by tests from include/*.h


def last_error():
    return lib().wb200_last_error().decode()


def _check(rc, what):
    if rc != 0:
        raise WorldB200Error("%s failed: %s" % (what, last_error()))


def init(device=0):
    _check(lib().wb200_init(int(device)), "wb200_init")


def launch_count():
    return int(lib().wb200_launch_count())


def stage_times():
    out = (C.c_float * 6)()
    lib().wb200_stage_times(out)
    return dict(zip(["dio", "stonemask", "cheaptrick", "d4c", "synthesis", "harvest"], list(out)))


def set_stream(cuda_stream):
    """cuda_stream: a cudaStream_t as an integer (e.g. torch.cuda.Stream().cuda_stream)."""
    _check(lib().wb200_set_stream(C.c_void_p(int(cuda_stream))), "wb200_set_stream")


def kernel_timing(on=True):
    lib().wb200_kernel_timing(int(bool(on)))


def kernel_times_reset():
    lib().wb200_kernel_times_reset()


def kernel_time(name):
    """-> (total ms, launches) of the named kernel since the last reset."""
    ms, n = C.c_double(0.0), C.c_longlong(0)
    lib().wb200_kernel_time(name.encode(), C.byref(ms), C.byref(n))
    return ms.value, n.value


def build_info():
    """Run-time switches of the library that change what a kernel executes (bench.py's roofline needs
    to know the precision of each transform)."""
    return {"lovetrain_fp32": bool(lib().wb200_option(b"lovetrain_fp32"))}


def fma_peak_tflops(fp64=True):
    return float(lib().wb200_measure_fma_peak(int(bool(fp64))))


def set_copy_deferral(on):
    """Deferred bulk copies (include/world_b200.h): asynchronous uploads / downloads are issued right before D4C's main
    kernel instead of beside a stage whose read-backs they would delay.  For pipelined callers."""
    _check(lib().wb200_set_copy_deferral(int(bool(on))), "set_copy_deferral")


def trim():
    """Hand the library's cached scratch memory back to the driver (include/world_b200.h)."""
    _check(lib().wb200_trim(), "trim")


def sync():
    _check(lib().wb200_sync(), "wb200_sync")


def randn_stream(n):
    out = np.zeros(int(n))
    _check(lib().wb200_randn_stream(_ptr(out), int(n)), "wb200_randn_stream")
    return out


def _nan_check(a, what):
    if a.size and np.isnan(a).all():
        raise WorldB200Error("%s failed: %s" % (what, last_error()))
    return a


# ---------------------------------------------------------------------------------------------
# one-utterance calls: same names, argument meaning and defaults as the reference C API
# ---------------------------------------------------------------------------------------------
def dio_option(frame_period=5.0, f0_floor=71.0, f0_ceil=800.0, speed=1, allowed_range=0.1,
               channels_in_octave=2.0):
    o = DioOption()
    lib().InitializeDioOption(C.byref(o))
    o.frame_period, o.f0_floor, o.f0_ceil = frame_period, f0_floor, f0_ceil
    o.speed, o.allowed_range, o.channels_in_octave = speed, allowed_range, channels_in_octave
    return o


def dio(x, fs, **kw):
    x = np.ascontiguousarray(x, np.float64)
    o = dio_option(**kw)
    n = lib().GetSamplesForDIO(fs, len(x), o.frame_period)
    t, f0 = np.zeros(n), np.zeros(n)
    lib().Dio(_ptr(x), len(x), fs, C.byref(o), _ptr(t), _ptr(f0))
    return t, _nan_check(f0, "Dio")


def stonemask(x, fs, t, f0):
    x, t, f0 = (np.ascontiguousarray(a, np.float64) for a in (x, t, f0))
    out = np.zeros_like(f0)
    lib().StoneMask(_ptr(x), len(x), fs, _ptr(t), _ptr(f0), len(f0), _ptr(out))
    return _nan_check(out, "StoneMask")


def harvest(x, fs, frame_period=5.0, f0_floor=71.0, f0_ceil=800.0):
    x = np.ascontiguousarray(x, np.float64)
    o = HarvestOption()
    lib().InitializeHarvestOption(C.byref(o))
    o.frame_period, o.f0_floor, o.f0_ceil = frame_period, f0_floor, f0_ceil
    n = lib().GetSamplesForHarvest(fs, len(x), frame_period)
    t, f0 = np.zeros(n), np.zeros(n)
    lib().Harvest(_ptr(x), len(x), fs, C.byref(o), _ptr(t), _ptr(f0))
    return t, _nan_check(f0, "Harvest")


def cheaptrick_option(fs, q1=-0.15, f0_floor=71.0, fft_size=None):
    o = CheapTrickOption()
    lib().InitializeCheapTrickOption(fs, C.byref(o))
    o.q1, o.f0_floor = q1, f0_floor
    o.fft_size = fft_size or lib().GetFFTSizeForCheapTrick(fs, C.byref(o))
    return o


def cheaptrick(x, fs, t, f0, **kw):
    x, t, f0 = (np.ascontiguousarray(a, np.float64) for a in (x, t, f0))
    o = cheaptrick_option(fs, **kw)
    sp = np.zeros((len(f0), o.fft_size // 2 + 1))
    lib().CheapTrick(_ptr(x), len(x), fs, _ptr(t), _ptr(f0), len(f0), C.byref(o), _rows(sp))
    return _nan_check(sp, "CheapTrick")


def d4c(x, fs, t, f0, fft_size, threshold=0.0):
    x, t, f0 = (np.ascontiguousarray(a, np.float64) for a in (x, t, f0))
    o = D4COption()
    lib().InitializeD4COption(C.byref(o))
    o.threshold = threshold
    ap = np.zeros((len(f0), fft_size // 2 + 1))
    lib().D4C(_ptr(x), len(x), fs, _ptr(t), _ptr(f0), len(f0), fft_size, C.byref(o), _rows(ap))
    return _nan_check(ap, "D4C")


def synthesis(f0, sp, ap, fft_size, frame_period, fs, y_length=None):
    f0, sp, ap = (np.ascontiguousarray(a, np.float64) for a in (f0, sp, ap))
    if y_length is None:   # W/test/synth.cpp:259
        y_length = int((len(f0) - 1) * frame_period / 1000.0 * fs) + 1
    y = np.zeros(y_length)
    lib().Synthesis(_ptr(f0), len(f0), _rows(sp), _rows(ap), fft_size, float(frame_period), fs,
                    y_length, _ptr(y))
    return _nan_check(y, "Synthesis")


def code_spectral_envelope(sp, fs, fft_size, number_of_dimensions):
    sp = np.ascontiguousarray(sp, np.float64)
    out = np.zeros((sp.shape[0], number_of_dimensions))
    lib().CodeSpectralEnvelope(_rows(sp), sp.shape[0], fs, fft_size, number_of_dimensions, _rows(out))
    return _nan_check(out, "CodeSpectralEnvelope")


def decode_spectral_envelope(coded, fs, fft_size):
    coded = np.ascontiguousarray(coded, np.float64)
    out = np.zeros((coded.shape[0], fft_size // 2 + 1))
    lib().DecodeSpectralEnvelope(_rows(coded), coded.shape[0], fs, fft_size, coded.shape[1], _rows(out))
    return _nan_check(out, "DecodeSpectralEnvelope")


# ---------------------------------------------------------------------------------------------
# the batched path
# ---------------------------------------------------------------------------------------------
class Corpus:
    """A batch of utterances resident in HBM (include/world_b200.h)."""

    def __init__(self, fs, x_lengths, frame_period=5.0):
        self.fs, self.frame_period = int(fs), float(frame_period)
        self.x_lengths = np.ascontiguousarray(x_lengths, np.int32)
        self.n_utt = len(self.x_lengths)
        self._h = lib().wb200_batch_create(self.fs, self.frame_period, self.n_utt,
                                           self.x_lengths.ctypes.data_as(_ip))
        if not self._h:
            raise WorldB200Error("wb200_batch_create failed: " + last_error())
        self.total_frames = lib().wb200_batch_total_frames(self._h)
        self.f_off = np.zeros(self.n_utt, np.int32)
        self.f_len = np.zeros(self.n_utt, np.int32)
        lib().wb200_batch_frame_layout(self._h, self.f_off.ctypes.data_as(_ip),
                                       self.f_len.ctypes.data_as(_ip))
        self.fft_size = None

    def close(self):
        if self._h:
            lib().wb200_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # inputs -------------------------------------------------------------------------------
    def upload_pcm16(self, pcm):
        """pcm: int16 numpy array or pinned torch tensor, utterances back to back."""
        ptr = pcm.data_ptr() if hasattr(pcm, "data_ptr") else pcm.ctypes.data
        _check(lib().wb200_batch_upload_pcm16(self._h, ptr), "upload_pcm16")
        self._keep = pcm

    def upload_pcm16_async(self, pcm):
        """Pinned int16 tensor / array; the copy overlaps whatever is running (another batch)."""
        ptr = pcm.data_ptr() if hasattr(pcm, "data_ptr") else pcm.ctypes.data
        _check(lib().wb200_batch_upload_pcm16_async(self._h, ptr), "upload_pcm16_async")
        self._keep = pcm

    def y_pcm16_async(self, out):
        ptr = out.data_ptr() if hasattr(out, "data_ptr") else out.ctypes.data
        _check(lib().wb200_batch_get_y_pcm16_async(self._h, ptr), "get_y_pcm16_async")

    def coded_async(self, lf0, mgc, bap):
        _check(lib().wb200_batch_get_coded_async(self._h, *(a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data
                                                            for a in (lf0, mgc, bap))), "get_coded_async")

    def wait_downloads(self):
        """Block until this batch's asynchronous result copies (coded_async, y_pcm16_async) have landed."""
        _check(lib().wb200_batch_wait_downloads(self._h), "wait_downloads")

    def set_pcm16_device(self, dev_tensor):
        _check(lib().wb200_batch_set_pcm16_device(self._h, dev_tensor.data_ptr()), "set_pcm16_device")

    def upload_f64(self, x):
        x = np.ascontiguousarray(x, np.float64)
        _check(lib().wb200_batch_upload_f64(self._h, x.ctypes.data), "upload_f64")

    # stages -------------------------------------------------------------------------------
    def dio(self, **kw):
        kw.setdefault("frame_period", self.frame_period)
        o = dio_option(**kw)
        _check(lib().wb200_batch_dio(self._h, C.byref(o)), "batch dio")

    def stonemask(self):
        _check(lib().wb200_batch_stonemask(self._h), "batch stonemask")

    def harvest(self, f0_floor=71.0, f0_ceil=800.0):
        o = HarvestOption(f0_floor, f0_ceil, self.frame_period)
        _check(lib().wb200_batch_harvest(self._h, C.byref(o)), "batch harvest")

    def cheaptrick(self, **kw):
        o = cheaptrick_option(self.fs, **kw)
        self.fft_size = o.fft_size
        _check(lib().wb200_batch_cheaptrick(self._h, C.byref(o)), "batch cheaptrick")

    def d4c(self, threshold=0.0, fft_size=None):
        self.fft_size = fft_size or self.fft_size or cheaptrick_option(self.fs).fft_size
        o = D4COption(threshold)
        _check(lib().wb200_batch_d4c(self._h, self.fft_size, C.byref(o)), "batch d4c")

    def synthesis(self, y_lengths=None):
        p = None
        if y_lengths is not None:
            y_lengths = np.ascontiguousarray(y_lengths, np.int32)
            p = y_lengths.ctypes.data_as(_ip)
        _check(lib().wb200_batch_synthesis(self._h, p), "batch synthesis")

    def analyze(self, threshold=0.0, f0="dio"):
        """Dio -> StoneMask -> CheapTrick -> D4C with the analysis tool's options
        (W/test/analysis.cpp:93-203); f0="harvest" swaps the F0 estimator (BASELINE config 3)."""
        if f0 == "harvest":
            self.harvest()
        else:
            self.dio()
            self.stonemask()
        self.cheaptrick()
        self.d4c(threshold=threshold)

    # results ------------------------------------------------------------------------------
    def f0(self, refined=True):
        out = np.zeros(self.total_frames)
        _check(lib().wb200_batch_get_f0(self._h, _ptr(out), int(refined)), "get_f0")
        return out

    def set_f0(self, f0, refined=True):
        f0 = np.ascontiguousarray(f0, np.float64)
        assert len(f0) == self.total_frames
        _check(lib().wb200_batch_set_f0(self._h, _ptr(f0), int(refined)), "set_f0")

    def sp(self):
        out = np.zeros((self.total_frames, self.fft_size // 2 + 1))
        _check(lib().wb200_batch_get_sp(self._h, out.ctypes.data), "get_sp")
        return out

    def ap(self):
        out = np.zeros((self.total_frames, self.fft_size // 2 + 1))
        _check(lib().wb200_batch_get_ap(self._h, out.ctypes.data), "get_ap")
        return out

    def set_sp_ap(self, fft_size, sp, ap):
        sp, ap = (np.ascontiguousarray(a, np.float64) for a in (sp, ap))
        self.fft_size = int(fft_size)
        _check(lib().wb200_batch_set_sp_ap(self._h, self.fft_size, _ptr(sp), _ptr(ap)), "set_sp_ap")

    def set_params_f32(self, fft_size, f0, sp, ap):
        """float32 f0 / sp / ap as the synth tool reads them (W/test/synth.cpp:160-190); numpy arrays or
        pinned torch tensors."""
        self.fft_size = int(fft_size)
        ptr = lambda a: None if a is None else (a.data_ptr() if hasattr(a, "data_ptr") else np.ascontiguousarray(a, np.float32).ctypes.data)
        keep = [None if a is None or hasattr(a, "data_ptr") else np.ascontiguousarray(a, np.float32) for a in (f0, sp, ap)]
        ptrs = [ptr(k if k is not None else a) for k, a in zip(keep, (f0, sp, ap))]
        _check(lib().wb200_batch_set_params_f32(self._h, self.fft_size, *ptrs), "set_params_f32")

    def utterance(self, u, want=("f0_raw", "f0", "sp", "ap", "y")):
        """One utterance's slice of the results as a dict (spot checks of a corpus-sized batch)."""
        n, H = int(self.f_len[u]), (self.fft_size or 0) // 2 + 1
        out = {}
        if "f0_raw" in want:
            out["f0_raw"] = np.zeros(n)
        if "f0" in want:
            out["f0"] = np.zeros(n)
        if "sp" in want:
            out["sp"] = np.zeros((n, H))
        if "ap" in want:
            out["ap"] = np.zeros((n, H))
        if "y" in want:
            out["y"] = np.zeros(int(self.y_layout()[1][u]))
        p = [out[k].ctypes.data if k in out else None for k in ("f0_raw", "f0", "sp", "ap", "y")]
        _check(lib().wb200_batch_get_utterance(self._h, int(u), *p), "get_utterance")
        return out

    def y_layout(self):
        off = np.zeros(self.n_utt, np.int64)
        ln = np.zeros(self.n_utt, np.int32)
        lib().wb200_batch_y_layout(self._h, off.ctypes.data_as(C.POINTER(C.c_longlong)),
                                   ln.ctypes.data_as(_ip))
        return off, ln

    def y(self):
        out = np.zeros(int(lib().wb200_batch_total_y(self._h)))
        _check(lib().wb200_batch_get_y(self._h, out.ctypes.data), "get_y")
        return out

    def y_pcm16(self, out=None):
        n = int(lib().wb200_batch_total_y(self._h))
        if out is None:
            out = np.zeros(n, np.int16)
        ptr = out.data_ptr() if hasattr(out, "data_ptr") else out.ctypes.data
        _check(lib().wb200_batch_get_y_pcm16(self._h, ptr), "get_y_pcm16")
        return out

    def code(self, mgc_dim=50, bap_dim=24):
        """The analysis tool's float32 lf0 / mgc / bap (W/test/analysis.cpp:293-390); dimensions as
        data/Makefile.in:214 passes them (MGCORDER + 1 = 50, 24 aperiodicity coefficients)."""
        self.mgc_dim, self.bap_dim = int(mgc_dim), int(bap_dim)
        _check(lib().wb200_batch_code(self._h, self.mgc_dim, self.bap_dim), "batch code")

    def coded(self):
        lf0 = np.zeros(self.total_frames, np.float32)
        mgc = np.zeros((self.total_frames, self.mgc_dim), np.float32)
        bap = np.zeros((self.total_frames, self.bap_dim), np.float32)
        _check(lib().wb200_batch_get_coded(self._h, lf0.ctypes.data, mgc.ctypes.data, bap.ctypes.data), "get_coded")
        return lf0, mgc, bap

    def decode_mgc(self, fft_size, mgc):
        mgc = np.ascontiguousarray(mgc, np.float32)
        self.fft_size = int(fft_size)
        _check(lib().wb200_batch_decode_mgc(self._h, self.fft_size, mgc.shape[1], mgc.ctypes.data), "decode_mgc")

    def set_coded_f32(self, fft_size, lf0=None, mgc=None, bap=None):
        """The coded branch of the synth tool (include/world_b200.h): float32 lf0 / mgc / bap -> f0 / sp / ap."""
        self.fft_size = int(fft_size)
        arrs = [None if a is None else np.ascontiguousarray(a, np.float32) for a in (lf0, mgc, bap)]
        mgc_dim = arrs[1].shape[1] if arrs[1] is not None else 0
        bap_dim = arrs[2].shape[1] if arrs[2] is not None else 0
        ptrs = [None if a is None else a.ctypes.data for a in arrs]
        _check(lib().wb200_batch_set_coded_f32(self._h, self.fft_size, mgc_dim, bap_dim, *ptrs), "set_coded_f32")

    def feature_stats(self):
        """[(1 + mgc_dim), 3] = {count, sum, sum of squares}: row 0 voiced lf0, rows 1.. mgc dims."""
        out = np.zeros((1 + self.mgc_dim, 3))
        _check(lib().wb200_batch_feature_stats(self._h, _ptr(out)), "feature_stats")
        return out

    def gv_stats(self):
        """Global-variance statistics (include/world_b200.h): (per-utterance variances [n_utt][mgc+1+bap],
        partials [mgc+1+bap][3] over this batch's utterances).  Needs code()."""
        ncol = self.mgc_dim + 1 + self.bap_dim
        per = np.zeros((self.n_utt, ncol))
        part = np.zeros((ncol, 3))
        _check(lib().wb200_batch_gv_stats(self._h, per.ctypes.data, part.ctypes.data), "gv_stats")
        return per, part

    def compose_cmp(self, streams=("mgc", "lf0", "bap"), windows=None, fetch=True):
        """The `cmp` target of data/Makefile.in:276-321 for the whole batch: every stream extended by
        its delta windows (data/scripts/window.pl) and merged side by side.  `streams`: names of the
        batch's own coded features ("mgc", "lf0", "bap") or (name, array[total_frames, dim]) pairs
        for streams computed on the host (Extract.py's lf0 / vib).  `windows`: one tuple of
        coefficient tuples per stream (default data/win/*.win[123]).  Returns [total_frames, cmp_dim]
        float32."""
        n = len(streams)
        arr = (CmpStream * n)()
        keep = []
        for i, st in enumerate(streams):
            wins = DEFAULT_WINDOWS if windows is None else windows[i]
            if isinstance(st, str):
                arr[i].source, arr[i].dim, arr[i].host_data = CMP_SRC[st], 0, None
            else:
                data = np.ascontiguousarray(st[1], np.float32)
                data = data.reshape(self.total_frames, data.shape[1] if data.ndim == 2 else 1)
                keep.append(data)
                arr[i].source, arr[i].dim, arr[i].host_data = CMP_SRC["host"], data.shape[1], data.ctypes.data
            if len(wins) > CMP_MAX_WINDOWS:
                raise WorldB200Error("at most %d windows per stream" % CMP_MAX_WINDOWS)
            arr[i].n_win = len(wins)
            for w, coef in enumerate(wins):
                if len(coef) > CMP_MAX_WIN_SIZE:
                    raise WorldB200Error("window longer than %d taps" % CMP_MAX_WIN_SIZE)
                arr[i].win_size[w] = len(coef)
                for k, v in enumerate(coef):
                    arr[i].win_coef[w][k] = float(v)
        _check(lib().wb200_batch_compose_cmp(self._h, arr, n), "compose_cmp")
        self.cmp_dim = int(lib().wb200_batch_cmp_dim(self._h))
        if not fetch:                       # the caller takes the matrix with cmp_async()
            return None
        out = np.zeros((self.total_frames, self.cmp_dim), np.float32)
        _check(lib().wb200_batch_get_cmp(self._h, out.ctypes.data), "get_cmp")
        return out

    def cmp_async(self, out):
        """Queue the copy of the cmp matrix into the pinned [total_frames, cmp_dim] float32 buffer `out`."""
        ptr = out.data_ptr() if hasattr(out, "data_ptr") else out.ctypes.data
        _check(lib().wb200_batch_get_cmp_async(self._h, ptr), "get_cmp_async")

    def cmp_stats(self):
        """[cmp_dim, 3] = {count, sum, sum of squares} of every cmp column (per-GPU partials)."""
        out = np.zeros((self.cmp_dim, 3))
        _check(lib().wb200_batch_cmp_stats(self._h, _ptr(out)), "cmp_stats")
        return out

    def lf0_stats(self):
        out = np.zeros(3)
        _check(lib().wb200_batch_lf0_stats(self._h, _ptr(out)), "lf0_stats")
        return out

    def frames_of(self, u):
        return slice(int(self.f_off[u]), int(self.f_off[u] + self.f_len[u]))
