"""Host-side mirror of the reference's block interface for the accelerated path.

Block names, setting names and error behaviour follow fair-acc/gnuradio4 (paths relative to /root/reference):
  fir_filter                blocks/filter/include/gnuradio-4.0/filter/time_domain_filter.hpp:20-48   (setting `b`)
  BasicDecimatingFilter     ... time_domain_filter.hpp:129-211 (filter_type, filter_response, filter_order, f_low, f_high,
                            sample_rate, decimate, fir_design_method)
  Decimator                 ... time_domain_filter.hpp:213-245 (decim)
  FFT                       blocks/fourier/include/gnuradio-4.0/fourier/fft.hpp:29-171 (fftSize, window, outputInDb,
                            outputInDeg, unwrapPhase, sample_rate)
  Add/Subtract/Multiply/DivideConst, Add/Subtract/Multiply/Divide   blocks/math/include/gnuradio-4.0/math/Math.hpp
  Rotator                   blocks/math/include/gnuradio-4.0/math/Rotator.hpp (sample_rate, frequency_shift,
                            phase_increment, initial_phase; frequency_shift XOR phase_increment)
Every block's `process_bulk` takes and returns CUDA tensors (torch supplies device memory and the current stream) and
issues exactly one C-ABI call (include/gr4b200.h) -- the body a `processBulk_cuda` member would have in C++.
Every block accepts `compute_domain` ("gpu:cuda:N", grammar of core/include/gnuradio-4.0/ComputeDomain.hpp:47-100).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import FILTER_TYPES, OPS, WINDOWS, Gr4b200Error, check, check_ptr


def parse_compute_domain(text):
    """ComputeDomain::parse: kind[:backend[:index]] -> (kind, backend, index); host aliases map to ("host","none",-1)."""
    if text in ("", "host", "default_cpu", "default_io"):
        return ("host", "none", -1)
    parts = text.split(":")
    kind = parts[0] if parts[0] in ("gpu", "fpga", "tpu") else "host"
    if kind == "host":
        return ("host", "none", -1)
    backend = parts[1] if len(parts) > 1 and parts[1] else ("sycl" if kind == "gpu" else "none")
    index = -1
    if len(parts) > 2:
        try:
            index = int(parts[2])
        except ValueError:
            index = -1
    return (kind, backend, index)


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)  # of the current device: block methods run on their own device


def _device_of_arguments(args, kwargs):
    """Device a block is being constructed for: `compute_domain="gpu:cuda:N"` if given with an index, else the device of a
    block passed in (DDC(mixer, fir), FirFft(fir, fft)), else the current device ("gpu:cuda" = wherever the caller is)."""
    domain = kwargs.get("compute_domain")
    if domain is not None:
        kind, backend, index = parse_compute_domain(domain)
        if kind == "gpu" and backend == "cuda" and index >= 0:
            return index
    for a in args:
        if isinstance(a, _Block):
            return a.device
    return torch.cuda.current_device() if torch.cuda.is_available() else 0


def _on_own_device(method):
    """Plans are allocated on, and launches issued from, the block's own device whatever the caller's current device is
    (the C ABI refuses a plan used from another device)."""
    import functools

    @functools.wraps(method)
    def bound(self, *args, **kwargs):
        try:
            guard = torch.cuda.device(self.device) if torch.cuda.is_available() else None
        except Exception:  # module teardown at interpreter exit (__del__): torch is half gone, the plan dies with the context
            guard = None
        if guard is None:
            return method(self, *args, **kwargs)
        with guard:
            return method(self, *args, **kwargs)

    return bound


def _construct_on_own_device(init):
    import functools

    @functools.wraps(init)
    def bound(self, *args, **kwargs):
        if not torch.cuda.is_available() or getattr(self, "_constructing", False):
            return init(self, *args, **kwargs)
        self._constructing = True  # subclass constructors call up the chain: bind once, at the outermost call
        try:
            self._construct_device = _device_of_arguments(args, kwargs)
            with torch.cuda.device(self._construct_device):
                return init(self, *args, **kwargs)
        finally:
            self._constructing = False

    return bound


def _require_cf32(x, what):
    if not isinstance(x, torch.Tensor) or not x.is_cuda or x.dtype != torch.complex64 or not x.is_contiguous():
        raise Gr4b200Error(f"{what}: expected a contiguous CUDA complex64 tensor (device edge buffer)")
    return x


def window(kind, n, beta=1.6):
    """gr::algorithm::window::create<float> -> numpy float32[n]."""
    kind = WINDOWS.index(kind) if isinstance(kind, str) else int(kind)
    out = np.zeros(n, dtype=np.float32)
    check(_lib.load().gr4b200_window_f32_host(kind, n, beta, out.ctypes.data_as(C.c_void_p)), "window")
    return out


def fir_generate(n_taps, window_type="Hamming", fc=0.1, beta=1.6, normalise_dc=True):
    """fir::generateCoefficients<float> (+ DC normalisation) -> numpy float32[n_taps]."""
    w = WINDOWS.index(window_type) if isinstance(window_type, str) else int(window_type)
    out = np.zeros(n_taps, dtype=np.float32)
    check(_lib.load().gr4b200_fir_generate_f32_host(n_taps, w, fc, beta, int(normalise_dc), out.ctypes.data_as(C.c_void_p)), "fir_generate")
    return out


def fir_design(filter_response="LOWPASS", filter_order=4, f_low=0.1, f_high=0.2, sample_rate=1.0, gain=1.0, attenuation_db=40.0, beta=1.6, window_type="Kaiser"):
    """fir::designFilter<float>(type, FilterParameters, window) -> numpy float32 taps."""
    w = WINDOWS.index(window_type) if isinstance(window_type, str) else int(window_type)
    t = FILTER_TYPES.index(filter_response) if isinstance(filter_response, str) else int(filter_response)
    cap = 1 << 16
    out = np.zeros(cap, dtype=np.float32)
    n = _lib.load().gr4b200_fir_design_f32_host(t, filter_order, f_low, f_high, sample_rate, gain, attenuation_db, beta, w, out.ctypes.data_as(C.c_void_p), cap)
    if n <= 0:
        raise Gr4b200Error(f"fir_design failed ({n})")
    return out[:n].copy()


class _Block:
    """Common part: compute_domain handling and numerator/denominator of the resampling ratio."""

    input_chunk_size = 1
    output_chunk_size = 1

    _GUARDED = ("settings_changed", "process_bulk", "process_bulk_real", "compute", "compute_real", "launch", "reset", "filter_stage", "fft_stage", "__del__")

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        for name in _Block._GUARDED:
            method = cls.__dict__.get(name)
            if callable(method):
                setattr(cls, name, _on_own_device(method))
        if "__init__" in cls.__dict__:
            cls.__init__ = _construct_on_own_device(cls.__dict__["__init__"])

    def __init__(self, compute_domain="gpu:cuda"):
        self.compute_domain = compute_domain
        kind, backend, index = parse_compute_domain(compute_domain)
        if kind != "gpu" or backend != "cuda":
            raise Gr4b200Error(f"compute_domain '{compute_domain}' is not a CUDA device domain; this package has no host path")
        # "gpu:cuda:N" binds device N; "gpu:cuda" binds the device that is current at construction
        self.device = index if index >= 0 else getattr(self, "_construct_device", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self._lib = _lib.load()

    in_item_bytes = 8   # complex<float>
    out_item_bytes = 8

    def n_outputs_for(self, n_in):
        return n_in // self.input_chunk_size * self.output_chunk_size

    def launch(self, stream, in_ptr, out_ptr, n_in):
        """One work chunk on raw device pointers: the body of `processBulk_cuda(stream, in, out, nIn, nOut)`."""
        raise NotImplementedError


class _MathOpConst(_Block):
    op = None

    def __init__(self, value=1.0, compute_domain="gpu:cuda"):
        super().__init__(compute_domain)
        self.value = complex(value)

    def launch(self, stream, in_ptr, out_ptr, n_in):
        check(self._lib.gr4b200_mathop_const_cf32(stream, OPS[self.op], in_ptr, out_ptr, n_in, self.value.real, self.value.imag), type(self).__name__)

    def process_bulk(self, x, out=None):
        x = _require_cf32(x, type(self).__name__)
        out = torch.empty_like(x) if out is None else out
        self.launch(_stream_ptr(), x.data_ptr(), out.data_ptr(), x.numel())
        return out


class AddConst(_MathOpConst):
    op = "add"


class SubtractConst(_MathOpConst):
    op = "subtract"


class MultiplyConst(_MathOpConst):
    op = "multiply"


class DivideConst(_MathOpConst):
    op = "divide"


class _MathOpMulti(_Block):
    op = None

    def __init__(self, n_inputs=2, compute_domain="gpu:cuda"):
        super().__init__(compute_domain)
        if not 1 <= n_inputs <= 32:  # Math.hpp:93 Limits<1U, 32U>
            raise Gr4b200Error("n_inputs must be in [1, 32]")
        self.n_inputs = n_inputs

    def process_bulk(self, inputs, out=None):
        if len(inputs) != self.n_inputs:
            raise Gr4b200Error(f"expected {self.n_inputs} inputs, got {len(inputs)}")
        inputs = [_require_cf32(i, type(self).__name__) for i in inputs]
        n = inputs[0].numel()
        if any(i.numel() != n for i in inputs):
            raise Gr4b200Error("all inputs must have the same length")
        out = torch.empty_like(inputs[0]) if out is None else out
        ptrs = (C.c_void_p * len(inputs))(*[i.data_ptr() for i in inputs])
        check(self._lib.gr4b200_mathop_multi_cf32(_stream_ptr(), OPS[self.op], ptrs, len(inputs), out.data_ptr(), n), type(self).__name__)
        return out


class Add(_MathOpMulti):
    op = "add"


class Subtract(_MathOpMulti):
    op = "subtract"


class Multiply(_MathOpMulti):
    op = "multiply"


class Divide(_MathOpMulti):
    op = "divide"


class Decimator(_Block):
    def __init__(self, decim=1, compute_domain="gpu:cuda"):
        super().__init__(compute_domain)
        self.decim = int(decim)
        self.input_chunk_size = self.decim

    def launch(self, stream, in_ptr, out_ptr, n_in):
        check(self._lib.gr4b200_decimate_cf32(stream, in_ptr, out_ptr, n_in, self.decim), "Decimator")

    def process_bulk(self, x, out=None):
        x = _require_cf32(x, "Decimator")
        n_out = (x.numel() + self.decim - 1) // self.decim
        out = torch.empty(n_out, dtype=torch.complex64, device=x.device) if out is None else out
        self.launch(_stream_ptr(), x.data_ptr(), out.data_ptr(), x.numel())
        return out


_ITEM_TYPES = {torch.float32: (0, 4), torch.int16: (1, 2), torch.int8: (2, 1)}  # GR4B200_ITEM_*, bytes per item


class InterleavedToComplex(_Block):
    """gr::blocks::type::converter::InterleavedToComplex<R, std::complex<float>> (basic/ConverterBlocks.hpp:258-277):
    2 n items (re, im, re, im, ...) of float32 / int16 / int8 in, n complex<float> out."""

    input_chunk_size = 2

    def __init__(self, dtype=torch.int16, compute_domain="gpu:cuda"):
        super().__init__(compute_domain)
        if dtype not in _ITEM_TYPES:
            raise Gr4b200Error(f"InterleavedToComplex: unsupported item type {dtype}")
        self.dtype = dtype
        self.item_type, self.in_item_bytes = _ITEM_TYPES[dtype]

    def launch(self, stream, in_ptr, out_ptr, n_in):
        check(self._lib.gr4b200_interleaved_to_complex_cf32(stream, self.item_type, in_ptr, out_ptr, n_in // 2), "InterleavedToComplex")

    def process_bulk(self, interleaved, out=None):
        if not (isinstance(interleaved, torch.Tensor) and interleaved.is_cuda and interleaved.dtype == self.dtype and interleaved.is_contiguous()):
            raise Gr4b200Error(f"InterleavedToComplex: expected a contiguous CUDA tensor of {self.dtype}")
        n = interleaved.numel() // 2
        out = torch.empty(n, dtype=torch.complex64, device=interleaved.device) if out is None else out
        self.launch(_stream_ptr(), interleaved.data_ptr(), out.data_ptr(), 2 * n)
        return out


class ComplexToInterleaved(_Block):
    """gr::blocks::type::converter::ComplexToInterleaved<std::complex<float>, R> (basic/ConverterBlocks.hpp:235-256):
    n complex<float> in, 2 n items of float32 / int16 / int8 out (static_cast: truncation toward zero)."""

    output_chunk_size = 2

    def __init__(self, dtype=torch.int16, compute_domain="gpu:cuda"):
        super().__init__(compute_domain)
        if dtype not in _ITEM_TYPES:
            raise Gr4b200Error(f"ComplexToInterleaved: unsupported item type {dtype}")
        self.dtype = dtype
        self.item_type, self.out_item_bytes = _ITEM_TYPES[dtype]

    def launch(self, stream, in_ptr, out_ptr, n_in):
        check(self._lib.gr4b200_complex_to_interleaved_cf32(stream, self.item_type, in_ptr, out_ptr, n_in), "ComplexToInterleaved")

    def process_bulk(self, x, out=None):
        x = _require_cf32(x, "ComplexToInterleaved")
        out = torch.empty(2 * x.numel(), dtype=self.dtype, device=x.device) if out is None else out
        self.launch(_stream_ptr(), x.data_ptr(), out.data_ptr(), x.numel())
        return out


class Rotator(_Block):
    def __init__(self, sample_rate=1.0, frequency_shift=None, phase_increment=None, initial_phase=0.0, compute_domain="gpu:cuda"):
        super().__init__(compute_domain)
        self._plan = None
        self.sample_rate = float(sample_rate)
        self.initial_phase = float(initial_phase)
        self.frequency_shift = 0.0
        self.phase_increment = 0.0
        self.settings_changed(frequency_shift=frequency_shift, phase_increment=phase_increment)

    def settings_changed(self, frequency_shift=None, phase_increment=None, initial_phase=None, sample_rate=None):
        """Rotator::settingsChanged (Rotator.hpp:40-49): exactly one of frequency_shift / phase_increment; phase restarts."""
        if sample_rate is not None:
            self.sample_rate = float(sample_rate)
        if initial_phase is not None:
            self.initial_phase = float(initial_phase)
        if frequency_shift is not None and phase_increment is not None:
            raise Gr4b200Error("cannot set both 'frequency_shift' and 'phase_increment' in new setting (XOR)")
        if frequency_shift is not None:
            self.frequency_shift = float(frequency_shift)
            self.phase_increment = float(self._lib.gr4b200_rotator_phase_increment(self.frequency_shift, self.sample_rate))
        elif phase_increment is not None:
            self.phase_increment = float(np.float32(phase_increment))
            self.frequency_shift = float(np.float32(np.float32(self.phase_increment) / (np.float32(2) * np.float32(np.pi)))) * self.sample_rate
        if self._plan:
            self._lib.gr4b200_rotator_plan_destroy(self._plan)
        self._plan = check_ptr(self._lib.gr4b200_rotator_plan_create(self.phase_increment, self.initial_phase), "rotator_plan_create")

    @property
    def accumulated_phase(self):
        with torch.cuda.device(self.device):
            return float(self._lib.gr4b200_rotator_get_phase(self._plan))

    def launch(self, stream, in_ptr, out_ptr, n_in):
        check(self._lib.gr4b200_rotator_cf32(self._plan, stream, in_ptr, out_ptr, n_in), "Rotator")

    def process_bulk(self, x, out=None):
        x = _require_cf32(x, "Rotator")
        out = torch.empty_like(x) if out is None else out
        self.launch(_stream_ptr(), x.data_ptr(), out.data_ptr(), x.numel())
        return out

    def __del__(self):
        if getattr(self, "_plan", None):
            self._lib.gr4b200_rotator_plan_destroy(self._plan)
            self._plan = None


class fir_filter(_Block):  # noqa: N801 -- reference spelling
    """y[n] = sum_k b[k] x[n-k] on a complex<float> (re/im independently) or float stream; history carried across calls.
    exact=True (default): the reference's summation order and rounding, bit-identical; exact=False: fused multiply-add;
    overlap_save=True: the tolerance mode through the 4096-point transform (complex<float>, full rate), bound by HBM."""

    def __init__(self, b=(1.0,), decimate=1, exact=True, overlap_save=False, compute_domain="gpu:cuda"):
        super().__init__(compute_domain)
        self._plan = None
        self.exact = bool(exact) and not overlap_save
        self.overlap_save = bool(overlap_save)
        self.decimate = int(decimate)
        self.input_chunk_size = self.decimate
        self.settings_changed(b=b)

    def settings_changed(self, b=None):
        if b is not None:
            self.b = np.ascontiguousarray(b, dtype=np.float32)
            if self.b.size == 0:
                raise Gr4b200Error("fir_filter: empty coefficient vector")
            if self._plan and not self.overlap_save:
                # a running filter keeps its past samples as the reference does (time_domain_filter.hpp:39-43): the history
                # buffer is replaced, and zeroed, only when the new `b` no longer fits it (32 samples, then bit_ceil(b.size()))
                check(self._lib.gr4b200_fir_plan_set_taps(self._plan, _stream_ptr(), self.b.ctypes.data_as(C.c_void_p), self.b.size), "fir_plan_set_taps")
                return
            if self._plan:
                self._lib.gr4b200_fir_plan_destroy(self._plan)
            self._plan = check_ptr(self._lib.gr4b200_fir_plan_create(self.b.ctypes.data_as(C.c_void_p), self.b.size, self.decimate, _lib.FIR_OVERLAP_SAVE if self.overlap_save else (_lib.FIR_EXACT if self.exact else _lib.FIR_FAST)), "fir_plan_create")

    def reset(self):
        check(self._lib.gr4b200_fir_plan_reset(self._plan, _stream_ptr()), "fir_plan_reset")

    def process_bulk(self, x, out=None):
        if not isinstance(x, torch.Tensor) or not x.is_cuda or not x.is_contiguous() or x.dtype not in (torch.complex64, torch.float32):
            raise Gr4b200Error("fir_filter: expected a contiguous CUDA complex64 or float32 tensor")
        if x.numel() % self.decimate != 0:
            raise Gr4b200Error("fir_filter: input length must be a multiple of the decimation factor")
        out = torch.empty(x.numel() // self.decimate, dtype=x.dtype, device=x.device) if out is None else out
        fn = self._lib.gr4b200_fir_cf32 if x.dtype == torch.complex64 else self._lib.gr4b200_fir_f32
        check(fn(self._plan, _stream_ptr(), x.data_ptr(), out.data_ptr(), x.numel()), "fir_filter")
        return out

    def launch(self, stream, in_ptr, out_ptr, n_in):
        check(self._lib.gr4b200_fir_cf32(self._plan, stream, in_ptr, out_ptr, n_in), "fir_filter")

    def __del__(self):
        if getattr(self, "_plan", None):
            self._lib.gr4b200_fir_plan_destroy(self._plan)
            self._plan = None


class BasicDecimatingFilter(fir_filter):
    """BasicFilterProto<T, Resampling<1,1,false>> restricted to filter_type == FIR (IIR is sequential: out of scope)."""

    def __init__(self, filter_type="FIR", filter_response="LOWPASS", filter_order=3, f_low=0.1, f_high=0.2, sample_rate=1.0, decimate=1, fir_design_method="Kaiser", exact=True, compute_domain="gpu:cuda"):
        if filter_type != "FIR":
            raise Gr4b200Error("only filter_type == 'FIR' runs on the device (IIR feedback is inherently serial)")
        self.filter_response, self.filter_order, self.f_low, self.f_high, self.sample_rate, self.fir_design_method = filter_response, filter_order, f_low, f_high, sample_rate, fir_design_method
        taps = fir_design(filter_response, filter_order, f_low, f_high, sample_rate, window_type=fir_design_method)
        super().__init__(b=taps, decimate=decimate, exact=exact, compute_domain=compute_domain)


class FFT(_Block):
    """FFT block: input chunks of fftSize samples; `process_bulk` returns the planar DataSet signal_values
    [chunk][4][N] = {Magnitude (fft-shifted), Phase (fft-shifted), Re, Im} (+ signal_ranges when asked), `compute`
    returns the plain spectrum like gr::algorithm::FFT::compute."""

    def __init__(self, fftSize=1024, window="Hann", outputInDb=False, outputInDeg=False, unwrapPhase=False, sample_rate=1.0, compute_domain="gpu:cuda"):  # noqa: N803 -- reference spelling
        super().__init__(compute_domain)
        self._plan = None
        self._plain = None
        self.outputInDb, self.outputInDeg, self.unwrapPhase, self.sample_rate = bool(outputInDb), bool(outputInDeg), bool(unwrapPhase), float(sample_rate)
        self.settings_changed(fftSize=fftSize, window=window)

    def settings_changed(self, fftSize=None, window=None):  # noqa: N803
        if fftSize is not None:
            self.fftSize = int(fftSize)
        if window is not None:
            names = [w.lower() for w in WINDOWS]
            self.window = WINDOWS[names.index(window.lower())] if isinstance(window, str) and window.lower() in names else getattr(self, "window", "Hann")
        self.input_chunk_size = self.fftSize
        self._window = globals()["window"](self.window, self.fftSize)
        for plan in (self._plan, self._plain):
            if plan:
                self._lib.gr4b200_fft_plan_destroy(plan)
        self._plan = check_ptr(self._lib.gr4b200_fft_plan_create(self.fftSize, self._window.ctypes.data_as(C.c_void_p)), "fft_plan_create")
        self._plain = check_ptr(self._lib.gr4b200_fft_plan_create(self.fftSize, None), "fft_plan_create")

    def frequency_axis(self):
        """createDataset axis (fft.hpp:188-196): i * fs/N - (N/2) * fs/N."""
        width = np.float32(self.sample_rate) / np.float32(self.fftSize)
        return (np.arange(self.fftSize, dtype=np.float32) * width - np.float32(self.fftSize // 2) * width).astype(np.float32)

    def flags(self):
        return (_lib.FFT_OUTPUT_IN_DB if self.outputInDb else 0) | (_lib.FFT_OUTPUT_IN_DEG if self.outputInDeg else 0) | (_lib.FFT_UNWRAP_PHASE if self.unwrapPhase else 0)

    def compute(self, x, out=None, windowed=False):
        x = _require_cf32(x, "FFT")
        if x.numel() % self.fftSize != 0:
            raise Gr4b200Error("FFT: input length must be a multiple of fftSize")
        out = torch.empty_like(x) if out is None else out
        check(self._lib.gr4b200_fft_c2c_cf32(self._plan if windowed else self._plain, _stream_ptr(), x.data_ptr(), out.data_ptr(), x.numel() // self.fftSize), "FFT")
        return out

    def compute_real(self, x, out=None, windowed=False):
        """gr::algorithm::FFT<float>::compute: real input, full N-bin spectrum per transform (Hermitian mirror included)."""
        if not isinstance(x, torch.Tensor) or not x.is_cuda or x.dtype != torch.float32 or not x.is_contiguous():
            raise Gr4b200Error("FFT.compute_real: expected a contiguous CUDA float32 tensor")
        if x.numel() % self.fftSize != 0:
            raise Gr4b200Error("FFT: input length must be a multiple of fftSize")
        out = torch.empty(x.numel(), dtype=torch.complex64, device=x.device) if out is None else out
        check(self._lib.gr4b200_fft_r2c_f32(self._plan if windowed else self._plain, _stream_ptr(), x.data_ptr(), out.data_ptr(), x.numel() // self.fftSize), "FFT")
        return out

    @property
    def out_item_bytes(self):  # one output item = one DataSet's signal_values: 4 planes of N floats
        return 4 * self.fftSize * 4

    output_chunk_size = 1

    def launch(self, stream, in_ptr, out_ptr, n_in):
        check(self._lib.gr4b200_fft_block_cf32(self._plan, stream, in_ptr, n_in // self.fftSize, self.flags(), out_ptr, None), "FFT")

    def process_bulk(self, x, signals=None, ranges=None, want_ranges=False):
        x = _require_cf32(x, "FFT")
        if x.numel() % self.fftSize != 0:
            raise Gr4b200Error("FFT: input length must be a multiple of fftSize")
        batch = x.numel() // self.fftSize
        signals = torch.empty((batch, 4, self.fftSize), dtype=torch.float32, device=x.device) if signals is None else signals
        if want_ranges and ranges is None:
            ranges = torch.empty((batch, 4, 2), dtype=torch.float32, device=x.device)
        check(self._lib.gr4b200_fft_block_cf32(self._plan, _stream_ptr(), x.data_ptr(), batch, self.flags(), signals.data_ptr(), ranges.data_ptr() if ranges is not None else None), "FFT")
        return (signals, ranges) if want_ranges else signals

    def process_bulk_real(self, x, signals=None, ranges=None, want_ranges=False):
        """The block on a real (float32) stream: planes of fftSize/2 values, [chunk][4][N/2] =
        {Magnitude of bins [0, N/2), Phase of the same bins, Re and Im of bins [N/2, N)} as the reference's FFT<float> emits them."""
        if not isinstance(x, torch.Tensor) or not x.is_cuda or x.dtype != torch.float32 or not x.is_contiguous():
            raise Gr4b200Error("FFT.process_bulk_real: expected a contiguous CUDA float32 tensor")
        if x.numel() % self.fftSize != 0:
            raise Gr4b200Error("FFT: input length must be a multiple of fftSize")
        batch = x.numel() // self.fftSize
        signals = torch.empty((batch, 4, self.fftSize // 2), dtype=torch.float32, device=x.device) if signals is None else signals
        if want_ranges and ranges is None:
            ranges = torch.empty((batch, 4, 2), dtype=torch.float32, device=x.device)
        check(self._lib.gr4b200_fft_block_f32(self._plan, _stream_ptr(), x.data_ptr(), batch, self.flags(), signals.data_ptr(), ranges.data_ptr() if ranges is not None else None), "FFT")
        return (signals, ranges) if want_ranges else signals

    def __del__(self):
        for name in ("_plan", "_plain"):
            if getattr(self, name, None):
                self._lib.gr4b200_fft_plan_destroy(getattr(self, name))
                setattr(self, name, None)


class FirFft(_Block):
    """fir_filter -> FFT block as one device call: the reference's compile-time Merge (BlockMerging.hpp:125-138) of the
    metric's two blocks. The filtered stream stays in shared memory; output = the FFT block's DataSet planes, bit for bit
    what the two blocks produce back to back."""

    output_chunk_size = 1

    def __init__(self, fir, fft):
        super().__init__(fir.compute_domain)
        self.fir, self.fft = fir, fft
        if not self._lib.gr4b200_fir_fft_fused_supported(fir._plan, fft._plan, fft.flags()):
            raise Gr4b200Error("FirFft: needs a full-rate complex FIR, fftSize 4096 and no phase unwrapping; connect the two blocks instead")
        self.input_chunk_size = fft.fftSize

    @property
    def out_item_bytes(self):
        return self.fft.out_item_bytes

    def launch(self, stream, in_ptr, out_ptr, n_in):
        check(self._lib.gr4b200_fir_fft_block_cf32(self.fir._plan, self.fft._plan, stream, in_ptr, n_in, self.fft.flags(), out_ptr), "FirFft")

    def process_bulk(self, x, signals=None):
        x = _require_cf32(x, "FirFft")
        if x.numel() % self.fft.fftSize != 0:
            raise Gr4b200Error("FirFft: input length must be a multiple of fftSize")
        batch = x.numel() // self.fft.fftSize
        signals = torch.empty((batch, 4, self.fft.fftSize), dtype=torch.float32, device=x.device) if signals is None else signals
        self.launch(_stream_ptr(), x.data_ptr(), signals.data_ptr(), x.numel())
        return signals


class DDC(_Block):
    """Rotator -> decimating FIR as one device call (the reference's compile-time Merge idea applied on the device)."""

    def __init__(self, mixer, fir):
        super().__init__(mixer.compute_domain)
        if mixer.device != fir.device:
            raise Gr4b200Error("DDC: the mixer and the filter live on different devices")
        self.mixer, self.fir = mixer, fir
        self.input_chunk_size = fir.decimate

    def launch(self, stream, in_ptr, out_ptr, n_in):
        check(self._lib.gr4b200_ddc_cf32(self.mixer._plan, self.fir._plan, stream, in_ptr, out_ptr, n_in), "DDC")

    def process_bulk(self, x, out=None):
        x = _require_cf32(x, "DDC")
        out = torch.empty(x.numel() // self.fir.decimate, dtype=torch.complex64, device=x.device) if out is None else out
        self.launch(_stream_ptr(), x.data_ptr(), out.data_ptr(), x.numel())
        return out


class PolyphaseChannelizer(_Block):
    """Critically sampled M-channel polyphase filter bank (no reference implementation exists; definition in DESIGN.md):
    stage 1 = polyphase FIR bank, stage 2 = M-point FFT per frame. Output [frame][channel]."""

    def __init__(self, prototype, n_channels, compute_domain="gpu:cuda"):
        super().__init__(compute_domain)
        self.prototype = np.ascontiguousarray(prototype, dtype=np.float32)
        self.n_channels = int(n_channels)
        if self.prototype.size % self.n_channels != 0:
            raise Gr4b200Error("prototype length must be a multiple of n_channels")
        self.taps_per_branch = self.prototype.size // self.n_channels
        self.input_chunk_size = self.n_channels
        self.output_chunk_size = self.n_channels
        self._plan = check_ptr(self._lib.gr4b200_pfb_plan_create(self.prototype.ctypes.data_as(C.c_void_p), self.n_channels, self.taps_per_branch), "pfb_plan_create")
        self._fft = None  # created with the first FFT stage: the filter bank alone works for any channel count

    def filter_stage(self, x, out=None):
        x = _require_cf32(x, "PolyphaseChannelizer")
        out = torch.empty_like(x) if out is None else out
        check(self._lib.gr4b200_pfb_filter_cf32(self._plan, _stream_ptr(), x.data_ptr(), out.data_ptr(), x.numel() // self.n_channels), "pfb_filter")
        return out

    def fft_stage(self, u, out=None):
        u = _require_cf32(u, "PolyphaseChannelizer")
        out = torch.empty_like(u) if out is None else out
        if self._fft is None:
            self._fft = check_ptr(self._lib.gr4b200_fft_plan_create(self.n_channels, None), "fft_plan_create")
        check(self._lib.gr4b200_fft_c2c_cf32(self._fft, _stream_ptr(), u.data_ptr(), out.data_ptr(), u.numel() // self.n_channels), "pfb_fft")
        return out

    @property
    def fused(self):
        return bool(self._lib.gr4b200_pfb_fused_supported(self._plan))

    def process_bulk(self, x, out=None, fused=None):
        """y[frame][channel]; one fused kernel when the shape allows it (256 channels, 4 / 8 / 12 taps per branch),
        otherwise filter bank and FFT back to back through an HBM edge."""
        use_fused = self.fused if fused is None else fused
        if not use_fused:
            return self.fft_stage(self.filter_stage(x), out).view(-1, self.n_channels)
        x = _require_cf32(x, "PolyphaseChannelizer")
        out = torch.empty_like(x) if out is None else out
        check(self._lib.gr4b200_pfb_channelizer_cf32(self._plan, _stream_ptr(), x.data_ptr(), out.data_ptr(), x.numel() // self.n_channels), "pfb_channelizer")
        return out.view(-1, self.n_channels)

    def __del__(self):
        if getattr(self, "_plan", None):
            self._lib.gr4b200_pfb_plan_destroy(self._plan)
            self._plan = None
        if getattr(self, "_fft", None):
            self._lib.gr4b200_fft_plan_destroy(self._fft)
            self._fft = None


class PolyphaseResampler(_Block):
    """Rational resampler, interpolation / decimation (no reference implementation exists; definition in DESIGN.md):
    y[m] = sum_k h[(m M) mod L + k L] x[floor(m M / L) - k]. Consumes multiples of `decimation`, produces L outputs per M inputs."""

    def __init__(self, taps, interpolation=1, decimation=1, compute_domain="gpu:cuda"):
        super().__init__(compute_domain)
        self.taps = np.ascontiguousarray(taps, dtype=np.float32)
        self.interpolation, self.decimation = int(interpolation), int(decimation)
        self.input_chunk_size, self.output_chunk_size = self.decimation, self.interpolation
        self._plan = check_ptr(self._lib.gr4b200_resampler_plan_create(self.taps.ctypes.data_as(C.c_void_p), self.taps.size, self.interpolation, self.decimation), "resampler_plan_create")

    def launch(self, stream, in_ptr, out_ptr, n_in):
        check(self._lib.gr4b200_resampler_cf32(self._plan, stream, in_ptr, out_ptr, n_in), "PolyphaseResampler")

    def process_bulk(self, x, out=None):
        x = _require_cf32(x, "PolyphaseResampler")
        if x.numel() % self.decimation != 0:
            raise Gr4b200Error("PolyphaseResampler: input length must be a multiple of the decimation factor")
        out = torch.empty(x.numel() // self.decimation * self.interpolation, dtype=torch.complex64, device=x.device) if out is None else out
        self.launch(_stream_ptr(), x.data_ptr(), out.data_ptr(), x.numel())
        return out

    def __del__(self):
        if getattr(self, "_plan", None):
            self._lib.gr4b200_resampler_plan_destroy(self._plan)
            self._plan = None
