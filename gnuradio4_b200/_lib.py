"""ctypes binding of the C ABI declared in include/gr4b200.h (libgr4b200.so, built in-tree by csrc/Makefile).

There is no CPU fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgr4b200.so")

OK = 0
ERROR = -100

OPS = {"add": 0, "subtract": 1, "multiply": 2, "divide": 3}
WINDOWS = ["None", "Rectangular", "Hamming", "Hann", "HannExp", "Blackman", "Nuttall", "BlackmanHarris", "BlackmanNuttall", "FlatTop", "Exponential", "Kaiser"]
FILTER_TYPES = ["LOWPASS", "HIGHPASS", "BANDPASS", "BANDSTOP"]
FIR_EXACT, FIR_FAST, FIR_OVERLAP_SAVE = 1, 0, 2
FFT_OUTPUT_IN_DB, FFT_OUTPUT_IN_DEG, FFT_UNWRAP_PHASE = 1, 2, 4

_vp, _sz, _f, _i, _u, _d, _l = C.c_void_p, C.c_size_t, C.c_float, C.c_int, C.c_uint, C.c_double, C.c_long

# name -> (restype, argtypes): every symbol include/gr4b200.h declares
SIGNATURES = {
    "gr4b200_abi_version": (_i, []),
    "gr4b200_last_error": (C.c_char_p, []),
    "gr4b200_device_count": (_i, []),
    "gr4b200_init": (_i, [_i]),
    "gr4b200_device_sm_count": (_i, [_i]),
    "gr4b200_malloc": (_vp, [_sz]),
    "gr4b200_free": (_i, [_vp]),
    "gr4b200_malloc_host": (_vp, [_sz]),
    "gr4b200_free_host": (_i, [_vp]),
    "gr4b200_memset": (_i, [_vp, _i, _sz, _vp]),
    "gr4b200_copy_h2d": (_i, [_vp, _vp, _sz, _vp]),
    "gr4b200_copy_d2h": (_i, [_vp, _vp, _sz, _vp]),
    "gr4b200_copy_d2d": (_i, [_vp, _vp, _sz, _vp]),
    "gr4b200_copy_d2h_2d": (_i, [_vp, _sz, _vp, _sz, _sz, _sz, _vp]),
    "gr4b200_launch_count": (C.c_ulonglong, []),
    "gr4b200_ipc_export": (_i, [_vp, _vp]),
    "gr4b200_ipc_open": (_vp, [_vp]),
    "gr4b200_ipc_close": (_i, [_vp]),
    "gr4b200_stream_write_value32": (_i, [_vp, _vp, _u]),
    "gr4b200_stream_wait_value32": (_i, [_vp, _vp, _u]),
    "gr4b200_stream_create": (_vp, []),
    "gr4b200_stream_destroy": (_i, [_vp]),
    "gr4b200_stream_synchronize": (_i, [_vp]),
    "gr4b200_event_create": (_vp, []),
    "gr4b200_event_destroy": (_i, [_vp]),
    "gr4b200_event_record": (_i, [_vp, _vp]),
    "gr4b200_stream_wait_event": (_i, [_vp, _vp]),
    "gr4b200_event_synchronize": (_i, [_vp]),
    "gr4b200_event_query": (_i, [_vp]),
    "gr4b200_event_elapsed_ms": (_i, [_vp, _vp, C.POINTER(_f)]),
    "gr4b200_ring_create": (_vp, [_i, _sz, _sz]),
    "gr4b200_ring_destroy": (_i, [_vp]),
    "gr4b200_ring_capacity": (_sz, [_vp]),
    "gr4b200_ring_available": (_sz, [_vp]),
    "gr4b200_ring_writable": (_sz, [_vp]),
    "gr4b200_ring_reserve": (_vp, [_vp, _sz, _vp]),
    "gr4b200_ring_publish": (_i, [_vp, _sz, _vp]),
    "gr4b200_ring_get": (_vp, [_vp, _sz, _vp]),
    "gr4b200_ring_consume": (_i, [_vp, _sz, _vp]),
    "gr4b200_ring_add_reader": (_i, [_vp]),
    "gr4b200_ring_available_for": (_sz, [_vp, _i]),
    "gr4b200_ring_get_for": (_vp, [_vp, _i, _sz, _vp]),
    "gr4b200_ring_consume_for": (_i, [_vp, _i, _sz, _vp]),
    "gr4b200_ring_pending_for": (_sz, [_vp, _i]),
    "gr4b200_ring_read_for": (_i, [_vp, _i, _sz, _vp, _vp]),
    "gr4b200_mathop_const_cf32": (_i, [_vp, _i, _vp, _vp, _sz, _f, _f]),
    "gr4b200_mathop_multi_cf32": (_i, [_vp, _i, _vp, _sz, _vp, _sz]),
    "gr4b200_interleaved_to_complex_cf32": (_i, [_vp, _i, _vp, _vp, _sz]),
    "gr4b200_complex_to_interleaved_cf32": (_i, [_vp, _i, _vp, _vp, _sz]),
    "gr4b200_decimate_cf32": (_i, [_vp, _vp, _vp, _sz, _sz]),
    "gr4b200_rotator_plan_create": (_vp, [_f, _f]),
    "gr4b200_rotator_plan_destroy": (_i, [_vp]),
    "gr4b200_rotator_set_phase": (_i, [_vp, _f]),
    "gr4b200_rotator_get_phase": (_f, [_vp]),
    "gr4b200_rotator_phase_increment": (_f, [_f, _f]),
    "gr4b200_rotator_cf32": (_i, [_vp, _vp, _vp, _vp, _sz]),
    "gr4b200_fir_plan_create": (_vp, [_vp, _sz, _sz, _i]),
    "gr4b200_fir_plan_destroy": (_i, [_vp]),
    "gr4b200_fir_plan_reset": (_i, [_vp, _vp]),
    "gr4b200_fir_plan_set_taps": (_i, [_vp, _vp, _vp, _sz]),
    "gr4b200_fir_cf32": (_i, [_vp, _vp, _vp, _vp, _sz]),
    "gr4b200_fir_f32": (_i, [_vp, _vp, _vp, _vp, _sz]),
    "gr4b200_fir_plan_history_items": (_sz, [_vp]),
    "gr4b200_fir_cf32_contiguous": (_i, [_vp, _vp, _vp, _vp, _sz]),
    "gr4b200_fir_f32_contiguous": (_i, [_vp, _vp, _vp, _vp, _sz]),
    "gr4b200_window_f32_host": (_i, [_i, _sz, _f, _vp]),
    "gr4b200_fir_generate_f32_host": (_i, [_sz, _i, _f, _f, _i, _vp]),
    "gr4b200_fir_design_f32_host": (_l, [_i, _sz, _d, _d, _d, _d, _d, _d, _i, _vp, _sz]),
    "gr4b200_fft_plan_create": (_vp, [_sz, _vp]),
    "gr4b200_fft_plan_destroy": (_i, [_vp]),
    "gr4b200_fft_plan_size": (_sz, [_vp]),
    "gr4b200_fft_c2c_cf32": (_i, [_vp, _vp, _vp, _vp, _sz]),
    "gr4b200_fft_r2c_f32": (_i, [_vp, _vp, _vp, _vp, _sz]),
    "gr4b200_fft_block_cf32": (_i, [_vp, _vp, _vp, _sz, _u, _vp, _vp]),
    "gr4b200_fft_block_f32": (_i, [_vp, _vp, _vp, _sz, _u, _vp, _vp]),
    "gr4b200_ddc_cf32": (_i, [_vp, _vp, _vp, _vp, _vp, _sz]),
    "gr4b200_pfb_plan_create": (_vp, [_vp, _sz, _sz]),
    "gr4b200_pfb_plan_destroy": (_i, [_vp]),
    "gr4b200_pfb_plan_reset": (_i, [_vp, _vp]),
    "gr4b200_pfb_filter_cf32": (_i, [_vp, _vp, _vp, _vp, _sz]),
    "gr4b200_fir_fft_fused_supported": (_i, [_vp, _vp, C.c_uint]),
    "gr4b200_fir_fft_block_cf32": (_i, [_vp, _vp, _vp, _vp, _sz, C.c_uint, _vp]),
    "gr4b200_pfb_fused_supported": (_i, [_vp]),
    "gr4b200_pfb_channelizer_cf32": (_i, [_vp, _vp, _vp, _vp, _sz]),
    "gr4b200_resampler_plan_create": (_vp, [_vp, _sz, _sz, _sz]),
    "gr4b200_resampler_plan_destroy": (_i, [_vp]),
    "gr4b200_resampler_plan_reset": (_i, [_vp, _vp]),
    "gr4b200_resampler_cf32": (_i, [_vp, _vp, _vp, _vp, _sz]),
    "gr4b200_peer_enable": (_i, [_i, _i]),
    "gr4b200_peer_copy": (_i, [_vp, _i, _vp, _i, _sz, _vp]),
}


class Gr4b200Error(RuntimeError):
    pass


_lib = None


def load():
    """Loads libgr4b200.so and types every entry point. Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Gr4b200Error(f"{LIB_PATH} is missing: build it with `make -C gnuradio4_b200/csrc` (or __graft_entry__.build()); there is no CPU fallback")
        lib = C.CDLL(os.environ.get("GR4B200_LIB", LIB_PATH))  # GR4B200_LIB: an alternative build of the same library (kernel A/B runs)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(status, what):
    if status != OK:
        message = load().gr4b200_last_error().decode(errors="replace")
        raise Gr4b200Error(f"{what} failed with status {status}: {message}")


def check_ptr(ptr, what):
    if not ptr:
        message = load().gr4b200_last_error().decode(errors="replace")
        raise Gr4b200Error(f"{what} failed: {message}")
    return ptr
