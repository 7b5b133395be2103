"""ctypes access to the CPU oracle (oracle/liboracle.so) and to the compiled reference (oracle/_ref/libgr4ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs -- never by the gnuradio4_b200 package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

WINDOWS = ["None", "Rectangular", "Hamming", "Hann", "HannExp", "Blackman", "Nuttall", "BlackmanHarris", "BlackmanNuttall", "FlatTop", "Exponential", "Kaiser"]
OPS = {"add": 0, "subtract": 1, "multiply": 2, "divide": 3}

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR, "all"], check=True, capture_output=True)


def _as_f32(x):
    """complex64 / float32 array -> float32 view (interleaved)."""
    x = np.ascontiguousarray(x)
    if x.dtype == np.complex64:
        return x.view(np.float32)
    return np.ascontiguousarray(x, dtype=np.float32)


class _Lib:
    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.path = path

    def _fn(self, name, argtypes, restype=C.c_int):
        fn = getattr(self.lib, self.prefix + name)
        fn.argtypes = argtypes
        fn.restype = restype
        return fn

    # ---- windows / design -------------------------------------------------------------------------------------
    def window(self, kind, n, beta=1.6, dtype=np.float32):
        kind = WINDOWS.index(kind) if isinstance(kind, str) else int(kind)
        out = np.zeros(n, dtype=dtype)
        if dtype == np.float32:
            rc = self._fn("window_f32", [C.c_int, C.c_size_t, C.c_float, _f32p])(kind, n, beta, out)
        else:
            rc = self._fn("window_f64", [C.c_int, C.c_size_t, C.c_double, _f64p])(kind, n, beta, out)
        if rc != 0:
            raise ValueError("window rejected arguments")
        return out

    def fir_generate(self, ntaps, window, fc, beta=1.6, normalise_dc=True):
        window = WINDOWS.index(window) if isinstance(window, str) else int(window)
        out = np.zeros(ntaps, dtype=np.float32)
        rc = self._fn("fir_generate_f32", [C.c_size_t, C.c_int, C.c_float, C.c_float, C.c_int, _f32p])(ntaps, window, fc, beta, int(normalise_dc), out)
        if rc != 0:
            raise ValueError(f"fir_generate failed rc={rc}")
        return out

    def fir_design(self, ftype, order, f_low, f_high, fs, gain=1.0, attenuation_db=40.0, beta=1.6, window="Kaiser", dtype=np.float32):
        window = WINDOWS.index(window) if isinstance(window, str) else int(window)
        ftype = ["LOWPASS", "HIGHPASS", "BANDPASS", "BANDSTOP"].index(ftype) if isinstance(ftype, str) else int(ftype)
        cap = 1 << 16
        out = np.zeros(cap, dtype=dtype)
        name, ptr = ("fir_design_f32", _f32p) if dtype == np.float32 else ("fir_design_f64", _f64p)
        n = self._fn(name, [C.c_int, C.c_size_t, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, ptr, C.c_size_t], C.c_long)(ftype, order, f_low, f_high, fs, gain, attenuation_db, beta, window, out, cap)
        if n < 0:
            raise ValueError(f"fir_design failed rc={n}")
        return out[:n].copy()

    def fir_magnitude_response(self, b, f_norm):
        b = np.ascontiguousarray(b, dtype=np.float64)
        return self._fn("fir_magnitude_response_f64", [_f64p, C.c_size_t, C.c_double], C.c_double)(b, b.size, f_norm)

    # ---- FFT ----------------------------------------------------------------------------------------------------
    def magnitude(self, X, db=False, shift=False):
        X = np.ascontiguousarray(X, dtype=np.complex64)
        out = np.zeros(X.size, dtype=np.float32)
        self._fn("magnitude_f32", [_f32p, C.c_size_t, C.c_int, C.c_int, _f32p])(X.view(np.float32), X.size, int(db), int(shift), out)
        return out

    def phase(self, X, deg=False, unwrap=False, shift=False):
        X = np.ascontiguousarray(X, dtype=np.complex64)
        out = np.zeros(X.size, dtype=np.float32)
        self._fn("phase_f32", [_f32p, C.c_size_t, C.c_int, C.c_int, C.c_int, _f32p])(X.view(np.float32), X.size, int(deg), int(unwrap), int(shift), out)
        return out

    def unwrap_phase(self, phase):
        p = np.array(phase, dtype=np.float64)
        self._fn("unwrap_phase_f64", [_f64p, C.c_size_t])(p, p.size)
        return p

    def fft_block(self, x, nfft, window, db=False, deg=False, unwrap=False, want_ranges=True):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        batch = x.size // nfft
        window = np.ascontiguousarray(window, dtype=np.float32)
        signals = np.zeros((batch, 4, nfft), dtype=np.float32)
        ranges = np.zeros((batch, 4, 2), dtype=np.float32)
        rc = self._fn("fft_block_cf32", [_f32p, C.c_size_t, C.c_size_t, _f32p, C.c_int, C.c_int, C.c_int, _f32p, _f32p])(x.view(np.float32), nfft, batch, window, int(db), int(deg), int(unwrap), signals, ranges)
        if rc != 0:
            raise ValueError("fft_block failed")
        return (signals, ranges) if want_ranges else signals

    def fft_block_real(self, x, nfft, window, db=False, deg=False, unwrap=False, want_ranges=True):
        """FFT block on real input: signals[c][4][nfft/2] = {|X[0..N/2)|*2/N, arg X[0..N/2), Re X[N/2..N), Im X[N/2..N)}."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        batch = x.size // nfft
        window = np.ascontiguousarray(window, dtype=np.float32)
        signals = np.zeros((batch, 4, nfft // 2), dtype=np.float32)
        ranges = np.zeros((batch, 4, 2), dtype=np.float32)
        rc = self._fn("fft_block_f32", [_f32p, C.c_size_t, C.c_size_t, _f32p, C.c_int, C.c_int, C.c_int, _f32p, _f32p])(x, nfft, batch, window, int(db), int(deg), int(unwrap), signals, ranges)
        if rc != 0:
            raise ValueError(f"fft_block_real failed ({rc})")
        return (signals, ranges) if want_ranges else signals

    # ---- math / mixer -------------------------------------------------------------------------------------------
    def mathop_const(self, op, x, value):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        out = np.zeros_like(x)
        value = complex(value)
        rc = self._fn("mathop_const_cf32", [C.c_int, _f32p, _f32p, C.c_size_t, C.c_float, C.c_float])(OPS[op], x.view(np.float32), out.view(np.float32), x.size, value.real, value.imag)
        assert rc == 0
        return out

    def mathop_multi(self, op, inputs):
        inputs = [np.ascontiguousarray(i, dtype=np.complex64) for i in inputs]
        n = inputs[0].size
        out = np.zeros(n, dtype=np.complex64)
        arr = (C.c_void_p * len(inputs))(*[i.ctypes.data for i in inputs])
        rc = self._fn("mathop_multi_cf32", [C.c_int, C.c_void_p, C.c_size_t, _f32p, C.c_size_t])(OPS[op], arr, len(inputs), out.view(np.float32), n)
        assert rc == 0
        return out

    # ---- sample-format converters ----------------------------------------------------------------------------------
    _ITEMS = {np.dtype(np.float32): 0, np.dtype(np.int16): 1, np.dtype(np.int8): 2}

    def interleaved_to_complex(self, interleaved):
        interleaved = np.ascontiguousarray(interleaved)
        n = interleaved.size // 2
        out = np.zeros(n, dtype=np.complex64)
        rc = self._fn("interleaved_to_complex_cf32", [C.c_int, C.c_void_p, _f32p, C.c_size_t])(self._ITEMS[interleaved.dtype], interleaved.ctypes.data, out.view(np.float32), n)
        assert rc == 0
        return out

    def complex_to_interleaved(self, x, dtype):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        out = np.zeros(2 * x.size, dtype=dtype)
        rc = self._fn("complex_to_interleaved_cf32", [C.c_int, _f32p, C.c_void_p, C.c_size_t])(self._ITEMS[np.dtype(dtype)], x.view(np.float32), out.ctypes.data, x.size)
        assert rc == 0
        return out

    def rotator(self, x, phase_increment, phase=0.0):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        out = np.zeros_like(x)
        ph = C.c_float(phase)
        rc = self._fn("rotator_cf32", [_f32p, _f32p, C.c_size_t, C.c_float, C.POINTER(C.c_float)])(x.view(np.float32), out.view(np.float32), x.size, phase_increment, C.byref(ph))
        assert rc == 0
        return out, ph.value


class Oracle(_Lib):
    """oracle/liboracle.so -- our CPU restatement."""

    def __init__(self, path):
        super().__init__(path, "oracle_")

    def fft(self, x, nfft=None):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        nfft = nfft or x.size
        out = np.zeros_like(x)
        self._fn("fft_c2c_f32", [_f32p, _f32p, C.c_size_t, C.c_size_t])(x.view(np.float32), out.view(np.float32), nfft, x.size // nfft)
        return out

    def fft_f64(self, x, nfft=None):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        nfft = nfft or x.size
        out = np.zeros(x.size, dtype=np.complex128)
        self._fn("fft_c2c_f32_via_f64", [_f32p, _f64p, C.c_size_t, C.c_size_t])(x.view(np.float32), out.view(np.float64), nfft, x.size // nfft)
        return out

    def _state(self, state, n):
        if state is None:
            return None, None
        assert state.dtype == np.float32 and state.size == n and state.flags.c_contiguous
        return state, state.ctypes.data_as(C.c_void_p)

    def fir(self, taps, x, state=None, decimate=1):
        """x float32 (real stream) or complex64; state: float32 array of (ntaps-1)*channels values, updated in place."""
        taps = np.ascontiguousarray(taps, dtype=np.float32)
        x = np.ascontiguousarray(x)
        cplx = x.dtype == np.complex64
        if not cplx:
            x = np.ascontiguousarray(x, dtype=np.float32)
        n = x.size
        ch = 2 if cplx else 1
        sp = None
        if state is not None:
            assert state.dtype == np.float32 and state.size == (taps.size - 1) * ch
            sp = state.ctypes.data_as(C.c_void_p)
        out = np.zeros(n // decimate, dtype=x.dtype)
        if decimate == 1:
            name = "fir_cf32" if cplx else "fir_f32"
            rc = self._fn(name, [_f32p, C.c_size_t, _f32p, _f32p, C.c_size_t, C.c_void_p])(taps, taps.size, _as_f32(x), _as_f32(out), n, sp)
        else:
            name = "fir_decim_cf32" if cplx else "fir_decim_f32"
            rc = self._fn(name, [_f32p, C.c_size_t, C.c_size_t, _f32p, _f32p, C.c_size_t, C.c_void_p])(taps, taps.size, decimate, _as_f32(x), _as_f32(out), n, sp)
        if rc != 0:
            raise ValueError("fir rejected arguments")
        return out

    def fir_f64(self, taps, x):
        taps = np.ascontiguousarray(taps, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros_like(x)
        self._fn("fir_f64", [_f64p, C.c_size_t, _f64p, _f64p, C.c_size_t, C.c_void_p])(taps, taps.size, x, out, x.size, None)
        return out

    def decimate(self, x, decim):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        out = np.zeros((x.size + decim - 1) // decim, dtype=np.complex64)
        self._fn("decimate_cf32", [_f32p, _f32p, C.c_size_t, C.c_size_t])(x.view(np.float32), out.view(np.float32), x.size, decim)
        return out

    def rotator_phases(self, n, phase_increment, phase=0.0, want=True):
        ph = C.c_float(phase)
        out = np.zeros(n, dtype=np.float32) if want else None
        self._fn("rotator_phases_f32", [C.c_size_t, C.c_float, C.POINTER(C.c_float), C.c_void_p])(n, phase_increment, C.byref(ph), out.ctypes.data_as(C.c_void_p) if want else None)
        return out, ph.value

    def rotator_phase_increment(self, frequency_shift, sample_rate=1.0):
        return self._fn("rotator_phase_increment_f32", [C.c_float, C.c_float], C.c_float)(frequency_shift, sample_rate)

    def resampler(self, taps, interpolation, decimation, x, state=None):
        taps = np.ascontiguousarray(taps, dtype=np.float32)
        x = np.ascontiguousarray(x, dtype=np.complex64)
        out = np.zeros(x.size // decimation * interpolation, dtype=np.complex64)
        sp = state.ctypes.data_as(C.c_void_p) if state is not None else None
        rc = self._fn("resampler_cf32", [_f32p, C.c_size_t, C.c_size_t, C.c_size_t, _f32p, _f32p, C.c_size_t, C.c_void_p])(taps, taps.size, interpolation, decimation, x.view(np.float32), out.view(np.float32), x.size, sp)
        assert rc == 0
        return out

    def pfb_filter(self, proto, n_channels, x, state=None):
        proto = np.ascontiguousarray(proto, dtype=np.float32)
        x = np.ascontiguousarray(x, dtype=np.complex64)
        taps_per_branch = proto.size // n_channels
        frames = x.size // n_channels
        out = np.zeros(frames * n_channels, dtype=np.complex64)
        sp = state.ctypes.data_as(C.c_void_p) if state is not None else None
        rc = self._fn("pfb_filter_cf32", [_f32p, C.c_size_t, C.c_size_t, _f32p, _f32p, C.c_size_t, C.c_void_p])(proto, n_channels, taps_per_branch, x.view(np.float32), out.view(np.float32), frames, sp)
        assert rc == 0
        return out

    def pfb_channelizer(self, proto, n_channels, x, state=None):
        proto = np.ascontiguousarray(proto, dtype=np.float32)
        x = np.ascontiguousarray(x, dtype=np.complex64)
        taps_per_branch = proto.size // n_channels
        frames = x.size // n_channels
        out = np.zeros(frames * n_channels, dtype=np.complex64)
        sp = None
        if state is not None:
            assert state.dtype == np.complex64 and state.size == (taps_per_branch - 1) * n_channels
            sp = state.ctypes.data_as(C.c_void_p)
        rc = self._fn("pfb_channelizer_cf32", [_f32p, C.c_size_t, C.c_size_t, _f32p, _f32p, C.c_size_t, C.c_void_p])(proto, n_channels, taps_per_branch, x.view(np.float32), out.view(np.float32), frames, sp)
        assert rc == 0
        return out.reshape(frames, n_channels)


class Ref(_Lib):
    """oracle/_ref/libgr4ref.so -- the reference's own sources compiled in place."""

    def __init__(self, path):
        super().__init__(path, "gr4ref_")

    def fft(self, x, nfft=None):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        nfft = nfft or x.size
        out = np.zeros_like(x)
        rc = self._fn("fft_c2c_f32_batch", [_f32p, _f32p, C.c_size_t, C.c_size_t])(x.view(np.float32), out.view(np.float32), nfft, x.size // nfft)
        if rc != 0:
            raise ValueError("reference FFT threw")
        return out

    def fir(self, taps, x, decimate=1):
        taps = np.ascontiguousarray(taps, dtype=np.float32)
        x = np.ascontiguousarray(x)
        cplx = x.dtype == np.complex64
        if not cplx:
            x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.zeros(x.size // decimate, dtype=x.dtype)
        if decimate == 1:
            name = "fir_cf32" if cplx else "fir_f32"
            self._fn(name, [_f32p, C.c_size_t, _f32p, _f32p, C.c_size_t])(taps, taps.size, _as_f32(x), _as_f32(out), x.size)
        else:
            name = "fir_decim_cf32" if cplx else "fir_decim_f32"
            self._fn(name, [_f32p, C.c_size_t, C.c_size_t, _f32p, _f32p, C.c_size_t])(taps, taps.size, decimate, _as_f32(x), _as_f32(out), x.size)
        return out

    def fir_f64(self, taps, x):
        taps = np.ascontiguousarray(taps, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros_like(x)
        self._fn("fir_f64", [_f64p, C.c_size_t, _f64p, _f64p, C.c_size_t])(taps, taps.size, x, out, x.size)
        return out


_oracle = None
_ref = None


def load_oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        src = os.path.join(ORACLE_DIR, "oracle.cpp")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            subprocess.run(["make", "-C", ORACLE_DIR, "liboracle.so"], check=True, capture_output=True)
        _oracle = Oracle(path)
    return _oracle


def load_ref():
    """Returns None when oracle/_ref/libgr4ref.so does not exist and cannot be built (no /root/reference)."""
    global _ref
    if _ref is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libgr4ref.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", ORACLE_DIR, "ref"], check=False, capture_output=True)
        if not os.path.exists(path):
            return None
        _ref = Ref(path)
    return _ref
