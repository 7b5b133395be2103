"""CPU-side check of the CUDA kernels' per-thread arithmetic: tests/host_emulation.cu runs the very functions the
__global__ kernels call (gnuradio4_b200/csrc/{fir,fft,rotator}_core.cuh) on the host, thread by thread, and this file
compares them with the oracle -- index mapping, summation order and the mixer's phase lifting are verified without a GPU."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def emul():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    out = os.path.join(ROOT, "build", "libhostemul.so")
    src = os.path.join(ROOT, "tests", "host_emulation.cu")
    deps = [src] + [os.path.join(ROOT, "gnuradio4_b200", "csrc", f) for f in ("fir_core.cuh", "fft_radix.cuh", "fft_large.cuh", "rotator_core.cuh", "sincos_core.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["nvcc", "-std=c++17", "-O1", "-Xcompiler", "-fPIC,-ffp-contract=off,-mfma,-pthread", "-shared", "-Wno-deprecated-gpu-targets", "-o", out, src], check=True, capture_output=True)
    lib = C.CDLL(out)
    lib.emul_fir.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, f32p, f32p, C.c_longlong, C.c_void_p]
    lib.emul_fir_bank_conflicts.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.emul_fft.argtypes = [C.c_int, f32p, f32p, C.c_longlong, C.c_void_p]
    lib.emul_fft_conflict_degree.argtypes = [C.c_int]
    lib.emul_fft_column_conflict_degree.argtypes = [C.c_int]
    lib.emul_fft_large.argtypes = [C.c_int, f32p, f32p, C.c_longlong, C.c_void_p, C.c_int]
    lib.emul_rotator_phases.argtypes = [C.c_float, C.c_float, C.c_ulonglong, np.ctypeslib.ndpointer(dtype=np.uint64), C.c_int, f32p]
    lib.emul_rotator_cycle.argtypes = [C.c_float, C.c_float, np.ctypeslib.ndpointer(dtype=np.uint64), C.c_int, f32p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    lib.emul_sincos_mismatches.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_uint]
    lib.emul_sincos_mismatches.restype = C.c_longlong
    return lib


def crand(rng, n):
    return (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)


@pytest.mark.parametrize("n_taps", [1, 3, 32, 33, 48, 64, 127, 128, 129, 130, 255, 300])
@pytest.mark.parametrize("decim", [1, 2, 4, 8, 16])
def test_fir_thread_mapping_is_bit_exact(emul, oracle, n_taps, decim):
    rng = np.random.default_rng(100 * n_taps + decim)
    n = 5000 // decim * decim
    x, taps = crand(rng, n), rng.uniform(-1, 1, n_taps).astype(np.float32)
    want = oracle.fir(taps, x, decimate=decim)
    got = np.zeros(n // decim, dtype=np.complex64)
    assert emul.emul_fir(taps, n_taps, decim, 1, 1, x.view(np.float32), got.view(np.float32), n, None) == 0
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    fast = np.zeros_like(got)
    emul.emul_fir(taps, n_taps, decim, 0, 1, x.view(np.float32), fast.view(np.float32), n, None)
    assert np.abs(fast - want).max() <= n_taps * 2.0**-23 * np.abs(taps).sum() * 2


@pytest.mark.parametrize("n_taps", [33, 127, 200])
def test_fir_full_size_tiles_of_long_calls_are_bit_exact(emul, oracle, n_taps):
    """Calls of more than 74 tiles take the 256 x 16 tiles (shorter ones the quarter-size tiles the test above runs)."""
    rng = np.random.default_rng(7 * n_taps)
    n = 75 * 4096 + 1234
    x, taps = crand(rng, n), rng.uniform(-1, 1, n_taps).astype(np.float32)
    want = oracle.fir(taps, x)
    got = np.zeros(n, dtype=np.complex64)
    assert emul.emul_fir(taps, n_taps, 1, 1, 1, x.view(np.float32), got.view(np.float32), n, None) == 0
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("decim", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("complex_stream", [0, 1])
def test_fir_tile_layout_is_bank_conflict_free(emul, decim, complex_stream):
    # window loads of every (half-)warp and the staging writes hit 32 distinct banks (fir_core.cuh TileLayout)
    for n_taps in (33, 64, 127, 255):
        assert emul.emul_fir_bank_conflicts(n_taps, decim, complex_stream) == 0, (n_taps, decim, complex_stream)


def test_fir_real_stream_and_history(emul, oracle):
    rng = np.random.default_rng(1)
    taps = rng.uniform(-1, 1, 127).astype(np.float32)
    x = rng.uniform(-1, 1, 9000).astype(np.float32)
    got = np.zeros_like(x)
    emul.emul_fir(taps, 127, 1, 1, 0, x, got, x.size, None)
    assert np.array_equal(got.view(np.uint32), oracle.fir(taps, x).view(np.uint32))
    # second chunk with the kernel-style history buffer (haloPad = 128 samples, right aligned)
    xc = crand(rng, 8192)
    want = oracle.fir(taps, xc)
    halo = np.zeros(128, dtype=np.complex64)
    halo[2:] = xc[4096 - 126 : 4096]
    second = np.zeros(4096, dtype=np.complex64)
    emul.emul_fir(taps, 127, 1, 1, 1, xc[4096:].view(np.float32).copy(), second.view(np.float32), 4096, halo.view(np.float32).ctypes.data_as(C.c_void_p))
    assert np.array_equal(second.view(np.uint32), want[4096:].view(np.uint32))


FFT_SIZES = [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192]


@pytest.mark.parametrize("windowed", [False, True])
@pytest.mark.parametrize("n", FFT_SIZES)
def test_fft_index_mapping(emul, oracle, n, windowed):
    # every pass of the radix family (gather / twiddle / butterfly / scatter per thread) against a float64 DFT
    rng = np.random.default_rng(2 + n)
    batch = 3
    x = crand(rng, n * batch)
    w = oracle.window("Hann", n) if windowed else None
    got = np.zeros_like(x)
    assert emul.emul_fft(n, x.view(np.float32), got.view(np.float32), batch, w.ctypes.data_as(C.c_void_p) if windowed else None) == 0
    xin = (x.reshape(batch, n) * w).astype(np.complex64).ravel() if windowed else x
    want = oracle.fft_f64(xin, n)
    for b in range(batch):
        sl = slice(b * n, (b + 1) * n)
        assert np.abs(got[sl] - want[sl]).max() <= 2.0e-6 * np.linalg.norm(xin[sl])


@pytest.mark.parametrize("windowed", [False, True])
@pytest.mark.parametrize("n", [16384, 32768, 65536, 131072, 262144])
def test_large_fft_column_passes(emul, oracle, n, windowed):
    # n = n1 * n2 as two passes of column transforms (fft_large.cuh): tile addressing, W_n twiddles, transposed store
    rng = np.random.default_rng(n)
    batch = 2
    x = crand(rng, n * batch)
    w = oracle.window("Hann", n) if windowed else None
    got = np.zeros_like(x)
    assert emul.emul_fft_large(n, x.view(np.float32), got.view(np.float32), batch, w.ctypes.data_as(C.c_void_p) if windowed else None, 0) == 0
    xin = (x.reshape(batch, n) * w).astype(np.complex64).ravel() if windowed else x
    want = oracle.fft_f64(xin, n)
    for b in range(batch):
        sl = slice(b * n, (b + 1) * n)
        assert np.abs(got[sl] - want[sl]).max() <= 2.0e-6 * np.linalg.norm(xin[sl])


@pytest.mark.parametrize("length", [128, 256, 512])
def test_large_fft_column_tiles_are_bank_conflict_free(emul, length):
    # 16 columns per tile at an odd region pitch: scatters, gathers, the natural-order park and the row-rotated read
    assert emul.emul_fft_column_conflict_degree(length) == 1


@pytest.mark.parametrize("n", [16384, 131072])
def test_large_fft_real_input(emul, oracle, n):
    # real samples into the first column pass; DC and Nyquist come out exactly real, the rest within the FFT tolerance
    rng = np.random.default_rng(n + 1)
    x = rng.uniform(-1, 1, n).astype(np.float32)
    got = np.zeros(n, dtype=np.complex64)
    assert emul.emul_fft_large(n, x, got.view(np.float32), 1, None, 1) == 0
    want = np.fft.fft(x.astype(np.float64))
    assert np.abs(got - want).max() <= 2.0e-6 * np.linalg.norm(x)
    assert got[0].imag == 0.0 and got[n // 2].imag == 0.0


@pytest.mark.parametrize("n", FFT_SIZES)
def test_fft_shared_memory_layout_is_bank_conflict_free(emul, n):
    # gathers, scatters, staged input reads and the block-mode parking slots of every (half-/quarter-)warp
    assert emul.emul_fft_conflict_degree(n) == 1


@pytest.mark.parametrize("dphi", [0.62831855, -0.62831855, 0.5, 1.5707964, -1.5707964, 3.0, -3.0, 3.14159, -3.14159, 1e-2, -1e-2, 2 * np.pi / 4096, 1.7, 2.3, -2.3, 1e-3, -1.234567])
def test_mixer_phase_lifting_reproduces_the_float_recurrence(emul, oracle, dphi):
    rng = np.random.default_rng(5)
    n = 1 << 21
    for phi0 in (0.0, 1.0, 6.0, 7.5, -3.0):
        ref, end = oracle.rotator_phases(n, float(np.float32(dphi)), phi0)
        m = np.concatenate([[0, 1, n], np.sort(rng.integers(1, n, 100))]).astype(np.uint64)
        out = np.zeros(m.size, dtype=np.float32)
        states = emul.emul_rotator_phases(float(np.float32(dphi)), phi0, n, m, m.size, out)
        assert states > 0
        want = np.array([np.float32(phi0) if mi == 0 else ref[int(mi) - 1] for mi in m], dtype=np.float32)
        assert np.array_equal(out.view(np.uint32), want.view(np.uint32)), (dphi, phi0)


@pytest.mark.parametrize("dphi,phi0", [(2 * np.pi * 0.1, 0.0), (-2 * np.pi * 0.1, 2.5), (3.0, 0.3), (-3.1, 0.0), (1e-3, 6.0), (0.7, 50.0), (0.25, -7.0), (0.0, 1.0), (0.0, 100.0), (1e-9, 3.0), (5.0, 0.1), (-4.0, 2.0)])
def test_mixer_phase_cycle_reproduces_the_recurrence(emul, oracle, dphi, phi0):
    """rotator.cu replays the phase recurrence once per plan until it closes (it is a map on the float patterns, hence
    eventually periodic) and gathers every call's checkpoints from that table: phase in front of sample m = cycle[m] below
    mu + lambda, cycle[mu + (m - mu) mod lambda] beyond. Checked against a plain replay up to several periods out."""
    dphi = float(np.float32(dphi))
    mu, lam = C.c_ulonglong(0), C.c_ulonglong(0)
    probe = np.zeros(1, dtype=np.uint64)
    out1 = np.zeros(1, dtype=np.float32)
    closed = emul.emul_rotator_cycle(dphi, phi0, probe, 1, out1, C.byref(mu), C.byref(lam))
    if abs(dphi) > np.pi and closed == 0:
        return  # beyond pi the landing states are not bounded by 2^23: the plan then keeps the per-call path
    assert closed == 1, "the recurrence must close within 40 M samples for |dphi| <= pi"
    size = mu.value + lam.value
    assert lam.value >= 1 and size <= 40 << 20
    n = int(min(size * 2 + 1000, 60_000_000))  # a plain replay of up to 60 M steps (C oracle)
    ref, _ = oracle.rotator_phases(n, dphi, phi0)
    rng = np.random.default_rng(5)
    m = np.unique(np.concatenate([[0, 1, 2, 7, 8, 9, n], np.clip([mu.value - 1, mu.value, mu.value + 1, size - 1, size, size + 1, size + lam.value], 0, n), rng.integers(0, n + 1, 4000)])).astype(np.uint64)
    out = np.zeros(m.size, dtype=np.float32)
    assert emul.emul_rotator_cycle(dphi, phi0, m, m.size, out, C.byref(mu), C.byref(lam)) == 1
    want = np.array([np.float32(phi0) if mi == 0 else ref[int(mi) - 1] for mi in m], dtype=np.float32)
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32)), (dphi, phi0, mu.value, lam.value)
    print(f"dphi={dphi} phi0={phi0}: transient {mu.value} samples, period {lam.value} samples")


def test_mixer_out_of_range_increment_takes_serial_path(emul):
    out = np.zeros(1, dtype=np.float32)
    m = np.zeros(1, dtype=np.uint64)
    for dphi in (0.0, 4.0, -5.0, float("nan"), float("inf")):
        assert emul.emul_rotator_phases(dphi, 0.0, 100, m, 1, out) == 0


def test_mixer_sincos_is_the_c_librarys(emul):
    """csrc/sincos_core.cuh restates glibc's sinf / cosf (FMA build) operation by operation; the mixer evaluates it on the
    device's FP64 pipe. Here the same code runs on the host against libm: every float in [0, 8) and (-8, 0] -- the whole
    band the reference's wrapped phase lives in, 2.2e9 arguments -- and every 97th of all 2^32 bit patterns (tiny, the
    4/pi table above 120, inf, NaN). scripts/verify_sincos_all_floats.cu is the exhaustive form (15 s on 8 threads)."""
    with open("/proc/cpuinfo") as f:
        flags = f.read()
    if " fma" not in flags or " avx2" not in flags:
        pytest.skip("this host's libm does not select the FMA build of sinf / cosf")
    assert emul.emul_sincos_mismatches(0x00000000, 0x41000000, 1) == 0
    assert emul.emul_sincos_mismatches(0x80000000, 0xC1000000, 1) == 0
    assert emul.emul_sincos_mismatches(0, 1 << 32, 97) == 0
