"""TEST INFRASTRUCTURE ONLY — ctypes binding of the *unmodified reference* WORLD_v2
library compiled by oracle/Makefile into oracle/_ref/libworld_ref.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (hts-train-world_b200) never does.

The functions mirror the reference C API one to one (W = externs/WORLD_v2):
  Dio            W/src/dio.cpp:642      StoneMask   W/src/stonemask.cpp:211
  CheapTrick     W/src/cheaptrick.cpp:200   D4C     W/src/d4c.cpp:337
  Synthesis      W/src/synthesis.cpp:338    Harvest W/src/harvest.cpp:1223
Options are the ones the reference's analysis tool sets (W/test/analysis.cpp:101-116,
152-163,190) unless overridden.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class DioOption(C.Structure):
    _fields_ = [("f0_floor", C.c_double), ("f0_ceil", C.c_double),
                ("channels_in_octave", C.c_double), ("frame_period", C.c_double),
                ("speed", C.c_int), ("allowed_range", C.c_double)]


class CheapTrickOption(C.Structure):
    _fields_ = [("q1", C.c_double), ("f0_floor", C.c_double), ("fft_size", C.c_int)]


class D4COption(C.Structure):
    _fields_ = [("threshold", C.c_double)]


class HarvestOption(C.Structure):
    _fields_ = [("f0_floor", C.c_double), ("f0_ceil", C.c_double),
                ("frame_period", C.c_double)]


_dp = C.POINTER(C.c_double)
_dpp = C.POINTER(_dp)


def _ptr(a):
    return a.ctypes.data_as(_dp)


def _rows(a2d):
    """double** view over a C-contiguous 2-D float64 array."""
    n = a2d.shape[0]
    arr = (_dp * n)()
    base = a2d.ctypes.data
    stride = a2d.strides[0]
    for i in range(n):
        arr[i] = C.cast(base + i * stride, _dp)
    return arr


def bind_world_api(lib):
    """Attach argtypes for the WORLD C API to a loaded library (reference or ours)."""
    lib.Dio.argtypes = [_dp, C.c_int, C.c_int, C.POINTER(DioOption), _dp, _dp]
    lib.Dio.restype = None
    lib.InitializeDioOption.argtypes = [C.POINTER(DioOption)]
    lib.GetSamplesForDIO.argtypes = [C.c_int, C.c_int, C.c_double]
    lib.GetSamplesForDIO.restype = C.c_int
    lib.StoneMask.argtypes = [_dp, C.c_int, C.c_int, _dp, _dp, C.c_int, _dp]
    lib.StoneMask.restype = None
    lib.CheapTrick.argtypes = [_dp, C.c_int, C.c_int, _dp, _dp, C.c_int,
                               C.POINTER(CheapTrickOption), _dpp]
    lib.CheapTrick.restype = None
    lib.InitializeCheapTrickOption.argtypes = [C.c_int, C.POINTER(CheapTrickOption)]
    lib.GetFFTSizeForCheapTrick.argtypes = [C.c_int, C.POINTER(CheapTrickOption)]
    lib.GetFFTSizeForCheapTrick.restype = C.c_int
    lib.GetF0FloorForCheapTrick.argtypes = [C.c_int, C.c_int]
    lib.GetF0FloorForCheapTrick.restype = C.c_double
    lib.D4C.argtypes = [_dp, C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int,
                        C.POINTER(D4COption), _dpp]
    lib.D4C.restype = None
    lib.InitializeD4COption.argtypes = [C.POINTER(D4COption)]
    lib.Synthesis.argtypes = [_dp, C.c_int, _dpp, _dpp, C.c_int, C.c_double, C.c_int,
                              C.c_int, _dp]
    lib.Synthesis.restype = None
    lib.Harvest.argtypes = [_dp, C.c_int, C.c_int, C.POINTER(HarvestOption), _dp, _dp]
    lib.Harvest.restype = None
    lib.InitializeHarvestOption.argtypes = [C.POINTER(HarvestOption)]
    lib.GetSamplesForHarvest.argtypes = [C.c_int, C.c_int, C.c_double]
    lib.GetSamplesForHarvest.restype = C.c_int
    if hasattr(lib, "CodeSpectralEnvelope"):
        lib.CodeSpectralEnvelope.argtypes = [_dpp, C.c_int, C.c_int, C.c_int, C.c_int, _dpp]
        lib.CodeSpectralEnvelope.restype = None
        lib.DecodeSpectralEnvelope.argtypes = [_dpp, C.c_int, C.c_int, C.c_int, C.c_int, _dpp]
        lib.DecodeSpectralEnvelope.restype = None
    return lib


class WorldLib:
    """Numpy-level calls into any library that exports the WORLD C API."""

    def __init__(self, path):
        self.path = path
        self.lib = bind_world_api(C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW))

    # ---- F0 -----------------------------------------------------------------
    def dio_option(self, frame_period=5.0, f0_floor=71.0, f0_ceil=800.0, speed=1,
                   allowed_range=0.1, channels_in_octave=2.0):
        o = DioOption()
        self.lib.InitializeDioOption(C.byref(o))
        o.frame_period, o.f0_floor, o.f0_ceil = frame_period, f0_floor, f0_ceil
        o.speed, o.allowed_range, o.channels_in_octave = speed, allowed_range, channels_in_octave
        return o

    def dio(self, x, fs, **kw):
        x = np.ascontiguousarray(x, np.float64)
        o = self.dio_option(**kw)
        n = self.lib.GetSamplesForDIO(fs, len(x), o.frame_period)
        t = np.zeros(n)
        f0 = np.zeros(n)
        self.lib.Dio(_ptr(x), len(x), fs, C.byref(o), _ptr(t), _ptr(f0))
        return t, f0

    def stonemask(self, x, fs, t, f0):
        x = np.ascontiguousarray(x, np.float64)
        t = np.ascontiguousarray(t, np.float64)
        f0 = np.ascontiguousarray(f0, np.float64)
        out = np.zeros_like(f0)
        self.lib.StoneMask(_ptr(x), len(x), fs, _ptr(t), _ptr(f0), len(f0), _ptr(out))
        return out

    def harvest(self, x, fs, frame_period=5.0, f0_floor=71.0, f0_ceil=800.0):
        x = np.ascontiguousarray(x, np.float64)
        o = HarvestOption()
        self.lib.InitializeHarvestOption(C.byref(o))
        o.frame_period, o.f0_floor, o.f0_ceil = frame_period, f0_floor, f0_ceil
        n = self.lib.GetSamplesForHarvest(fs, len(x), frame_period)
        t = np.zeros(n)
        f0 = np.zeros(n)
        self.lib.Harvest(_ptr(x), len(x), fs, C.byref(o), _ptr(t), _ptr(f0))
        return t, f0

    # ---- envelope / aperiodicity ------------------------------------------------
    def cheaptrick_option(self, fs, q1=-0.15, f0_floor=71.0, fft_size=None):
        o = CheapTrickOption()
        self.lib.InitializeCheapTrickOption(fs, C.byref(o))
        o.q1, o.f0_floor = q1, f0_floor
        o.fft_size = fft_size or self.lib.GetFFTSizeForCheapTrick(fs, C.byref(o))
        return o

    def cheaptrick(self, x, fs, t, f0, **kw):
        x = np.ascontiguousarray(x, np.float64)
        t = np.ascontiguousarray(t, np.float64)
        f0 = np.ascontiguousarray(f0, np.float64)
        o = self.cheaptrick_option(fs, **kw)
        sp = np.zeros((len(f0), o.fft_size // 2 + 1))
        self.lib.CheapTrick(_ptr(x), len(x), fs, _ptr(t), _ptr(f0), len(f0), C.byref(o),
                            _rows(sp))
        return sp

    def d4c(self, x, fs, t, f0, fft_size, threshold=0.0):
        x = np.ascontiguousarray(x, np.float64)
        t = np.ascontiguousarray(t, np.float64)
        f0 = np.ascontiguousarray(f0, np.float64)
        o = D4COption()
        self.lib.InitializeD4COption(C.byref(o))
        o.threshold = threshold
        ap = np.zeros((len(f0), fft_size // 2 + 1))
        self.lib.D4C(_ptr(x), len(x), fs, _ptr(t), _ptr(f0), len(f0), fft_size, C.byref(o),
                     _rows(ap))
        return ap

    # ---- synthesis ----------------------------------------------------------------
    def synthesis(self, f0, sp, ap, fft_size, frame_period, fs, y_length=None):
        f0 = np.ascontiguousarray(f0, np.float64)
        sp = np.ascontiguousarray(sp, np.float64)
        ap = np.ascontiguousarray(ap, np.float64)
        if y_length is None:   # W/test/synth.cpp:259
            y_length = int((len(f0) - 1) * frame_period / 1000.0 * fs) + 1
        y = np.zeros(y_length)
        self.lib.Synthesis(_ptr(f0), len(f0), _rows(sp), _rows(ap), fft_size,
                           float(frame_period), fs, y_length, _ptr(y))
        return y

    # ---- codec (W/src/codec.cpp:266-324) and the tool's coded outputs (W/test/analysis.cpp:293-390) --
    def code_spectral_envelope(self, sp, fs, fft_size, ndim):
        sp = np.ascontiguousarray(sp, np.float64)
        out = np.zeros((sp.shape[0], ndim))
        self.lib.CodeSpectralEnvelope(_rows(sp), sp.shape[0], fs, fft_size, ndim, _rows(out))
        return out

    def decode_spectral_envelope(self, coded, fs, fft_size):
        coded = np.ascontiguousarray(coded, np.float64)
        out = np.zeros((coded.shape[0], fft_size // 2 + 1))
        self.lib.DecodeSpectralEnvelope(_rows(coded), coded.shape[0], fs, fft_size, coded.shape[1], _rows(out))
        return out

    def tool_features(self, f0, sp, ap, fs, fft_size, mgc_dim=50, bap_dim=24):
        """float32 lf0 / mgc / bap exactly as the analysis tool writes them."""
        s = sp * 1e4
        s[s == 0.0] = 0.0001
        mgc = self.code_spectral_envelope(s, fs, fft_size, mgc_dim)
        mgc[:, 0] += 12.0
        bap = self.code_spectral_envelope(ap * 1e4, fs, fft_size, bap_dim)
        bap[:, 0] -= 9.210340
        tiny = (bap[:, 0] > 0) & (bap[:, 0] < 1e-4)
        bap[tiny, 0] = 0
        lf0 = np.where(f0 != 0, np.log(np.where(f0 != 0, f0, 1.0)), 0.0)
        return lf0.astype(np.float32), mgc.astype(np.float32), bap.astype(np.float32)

    # ---- the tool's whole analysis (W/test/analysis.cpp:243-398 without the codec tail) --
    def analyze(self, x, fs, frame_period=5.0, threshold=0.0):
        t, f0_raw = self.dio(x, fs, frame_period=frame_period)
        f0 = self.stonemask(x, fs, t, f0_raw)
        o = self.cheaptrick_option(fs)
        sp = self.cheaptrick(x, fs, t, f0)
        ap = self.d4c(x, fs, t, f0, o.fft_size, threshold=threshold)
        return dict(t=t, f0_raw=f0_raw, f0=f0, sp=sp, ap=ap, fft_size=o.fft_size)


def ref_path(opt=False):
    return os.path.join(_HERE, "_ref", "libworld_ref_O3.so" if opt else "libworld_ref.so")


_cache = {}


def load(opt=False):
    """The reference library.  Raises with a clear message when it was not built."""
    p = ref_path(opt)
    if p not in _cache:
        if not os.path.exists(p):
            raise FileNotFoundError(
                p + " missing: run `make -C oracle` in a container that has /root/reference")
        _cache[p] = WorldLib(p)
    return _cache[p]


def randn_stream(n):
    """First n values of the reference's randn() after randn_reseed()
    (W/src/matlabfunctions.cpp:247-277), straight from the compiled reference."""
    lib = load().lib
    lib.randn.restype = C.c_double
    lib.randn_reseed()
    return np.array([lib.randn() for _ in range(n)])
