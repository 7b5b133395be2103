"""Pins the CPU oracle (oracle/oracle.cpp) three ways:
  1. the reference's own known-answer tests (file:line of the reference test in each docstring),
  2. the committed golden fixtures generated from the reference's compiled sources (tests/golden/*.npz),
  3. live, against oracle/_ref/libgr4ref.so when it is present (skipped otherwise).
Bit-exact for windows, FIR design and FIR; FFT within the tolerance the tests state."""
import os

import numpy as np
import pytest

from tests import _oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FFT_TOL = 2.0e-6  # max |X - X_f64| / ||x||_2


def crand(seed, n):
    rng = np.random.default_rng(seed)
    return (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)


def bit_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


# ---- 1. reference known answers ---------------------------------------------------------------------------------------
def test_window_tables_qa_algorithm_fourier(oracle):
    """algorithm/test/qa_algorithm_fourier.cpp:153-180: 8-point tables for all window types."""
    ref = {
        "Rectangular": [1, 1, 1, 1, 1, 1, 1, 1],
        "Hamming": [0.07672, 0.25053218, 0.64108455, 0.9542833, 0.95428324, 0.6410846, 0.25053206, 0.07672],
        "Hann": [0, 0.1882550991, 0.611260467, 0.950484434, 0.950484434, 0.611260467, 0.1882550991, 0],
        "Blackman": [0, 0.09045342435, 0.4591829575, 0.9203636181, 0.9203636181, 0.4591829575, 0.09045342435, 0],
        "BlackmanHarris": [0.00006, 0.03339172348, 0.3328335043, 0.8893697722, 0.8893697722, 0.3328335043, 0.03339172348, 0.00006],
        "BlackmanNuttall": [0.0003628, 0.03777576895, 0.34272762, 0.8918518611, 0.8918518611, 0.34272762, 0.03777576895, 0.0003628],
        "Exponential": [1, 1.042546905, 1.08690405, 1.133148453, 1.181360413, 1.231623642, 1.284025417, 1.338656724],
        "FlatTop": [0.004, -0.1696424054, 0.04525319348, 3.622389212, 3.622389212, 0.04525319348, -0.1696424054, 0.004],
        "HannExp": [0, 0.611260467, 0.950484434, 0.1882550991, 0.1882550991, 0.950484434, 0.611260467, 0],
        "Nuttall": [0, 0.0311427368, 0.3264168059, 0.8876284573, 0.8876284573, 0.3264168059, 0.0311427368, 0],
        "Kaiser": [0.5714348848, 0.7650986027, 0.9113132365, 0.9899091685, 0.9899091685, 0.9113132365, 0.7650986027, 0.5714348848],
    }
    for name, table in ref.items():
        assert np.allclose(oracle.window(name, 8), np.array(table, dtype=np.float32), atol=1e-5), name
        assert np.allclose(oracle.window(name, 8, dtype=np.float64), table, atol=1e-5), name
    assert np.array_equal(oracle.window("None", 8), np.ones(8, dtype=np.float32))
    assert oracle.window("Hann", 0).size == 0
    with pytest.raises(ValueError):
        oracle.window("Kaiser", 8, beta=-1.0)


def test_unwrap_qa_algorithm_fourier(oracle):
    """qa_algorithm_fourier.cpp:145-151 (numpy.unwrap vector)."""
    phase = [0.2, -1.0, 2.5, -3.1, 0.9, -0.5, 1.2, 0.8, 1.5, -1.2, -2.7, 0.9, -0.8, -1.4, 0.6, 1.1, -1.9, 0.4, 1.3, -0.7]
    want = [0.2, -1.0, -3.78318531, -3.1, -5.38318531, -6.78318531, -5.08318531, -5.48318531, -4.78318531, -7.48318531, -8.98318531, -11.66637061, -13.36637061, -13.96637061, -11.96637061, -11.46637061, -14.46637061, -12.16637061, -11.26637061, -13.26637061]
    assert np.allclose(oracle.unwrap_phase(phase), want, atol=1e-7)


def test_fft_patterns_qa_algorithm_fourier(oracle):
    """qa_algorithm_fourier.cpp:97-143: N = 16 patterns, X[0] and peak amplitude to 1e-5."""
    cases = [(np.zeros(16), 0j, 0.0), (np.ones(16), 16 + 0j, 2.0), (np.ones(16) * (1 + 1j), 16 + 16j, np.sqrt(8.0)), (np.arange(1, 17), 136 + 0j, 17.0), (np.arange(16) % 2, 8 + 0j, 1.0)]
    for signal, x0, amp in cases:
        X = oracle.fft(signal.astype(np.complex64))
        mag = oracle.magnitude(X)
        assert abs(X[0] - x0) < 1e-5 and np.argmax(mag) == 0 and abs(mag[0] - amp) < 1e-5


@pytest.mark.parametrize("n", [256, 512, 4096])
def test_fft_sine_peak_qa_algorithm_fourier(oracle, n):
    """qa_algorithm_fourier.cpp:67-95 and algorithm/benchmarks/bm_fft.cpp:61-62: sine at bin 5 => Im X[5] = -N/2."""
    X = oracle.fft(np.sin(2 * np.pi * 5 * np.arange(n) / n).astype(np.complex64))
    assert abs(X[5].imag + n / 2) < 0.1
    mag = oracle.magnitude(X)
    assert np.argmax(mag[: n // 2]) == 5 and abs(mag[5] - 1.0) < 1e-5


def test_fft_roundtrip_linearity_qa_simdfft(oracle):
    """algorithm/test/qa_SimdFFT.cpp:131,198 (round trip 1e-5 N via conj trick), :421 (linearity 1e-4 N)."""
    n = 2048
    a, b = crand(1, n), crand(2, n)
    Fa, Fb = oracle.fft(a), oracle.fft(b)
    back = np.conj(oracle.fft(np.conj(Fa).astype(np.complex64))) / n
    assert np.abs(back - a).max() < 1e-5 * n
    assert np.abs(oracle.fft((2 * a + 3 * b).astype(np.complex64)) - (2 * Fa + 3 * Fb)).max() < 1e-4 * n
    assert abs(oracle.fft(np.ones(n, dtype=np.complex64))[0] - n) < 1e-4 * n  # DC bin :374-379


def test_fir_step_response_qa_filter(oracle):
    """blocks/filter/test/qa_filter.cpp:54-93: 10-tap boxcar of 0.1 on a unit step settles to 1 +- 1e-3 at index 10."""
    y = oracle.fir(np.full(10, 0.1, dtype=np.float32), np.ones(20, dtype=np.float32))
    assert abs(y[0] - 0.1) < 1e-7 and (np.abs(y[10:] - 1.0) < 1e-3).all() and (y[:9] < 0.95).all()


def test_fir_design_response_qa_filtertool(oracle):
    """algorithm/test/qa_FilterTool.cpp:470-503: {Kaiser,Hamming,Hann} x {LP,HP,BP,BS}, order 4, 1 / 10 Hz, fs 1 kHz."""
    for win in ("Kaiser", "Hamming", "Hann"):
        lp = oracle.fir_design("LOWPASS", 4, 1.0, 10.0, 1000.0, window=win, dtype=np.float64)
        assert lp.size % 2 == 1 and abs(oracle.fir_magnitude_response(lp, 0.0) - 1.0) < 0.01
        assert oracle.fir_magnitude_response(lp, 0.25) < 0.05
        hp = oracle.fir_design("HIGHPASS", 4, 1.0, 10.0, 1000.0, window=win, dtype=np.float64)
        assert abs(oracle.fir_magnitude_response(hp, 0.48) - 1.0) < 0.01 and oracle.fir_magnitude_response(hp, 0.0) < 0.05
        bs = oracle.fir_design("BANDSTOP", 4, 1.0, 10.0, 1000.0, window=win, dtype=np.float64)
        assert abs(oracle.fir_magnitude_response(bs, 0.0) - 1.0) < 0.01


def test_decimating_filter_qa_filter(oracle):
    """qa_filter.cpp:150-265 spirit: designed FIR low-pass passes 50 Hz (>= 0.9), rejects 300 Hz (<= 0.2) at fs = 1 kHz,
    also through the decimate = 5 path on 1000 samples processed in two halves."""
    taps = oracle.fir_design("LOWPASS", 4, 100.0, 200.0, 1000.0, window="Hamming")
    t = np.arange(1000) / 1000.0
    for f, check in ((50.0, lambda a: a >= 0.9), (300.0, lambda a: a <= 0.2)):
        x = np.sin(2 * np.pi * f * t).astype(np.float32)
        state = np.zeros(taps.size - 1, dtype=np.float32)
        y = np.concatenate([oracle.fir(taps, x[:500], state=state, decimate=5), oracle.fir(taps, x[500:], state=state, decimate=5)])
        assert y.size == 200 and check(np.abs(y[100:]).max())
        assert bit_equal(y, oracle.fir(taps, x, decimate=5))  # chunking does not change the result


def test_rotator_qa_rotator(oracle):
    """blocks/math/test/qa_Rotator.cpp:69-92: phase_increment pi/2, 8 samples of (1,0) -> cos/sin((i+1) pi/2) +- 1e-5."""
    y, _ = oracle.rotator(np.ones(8, dtype=np.complex64), np.pi / 2)
    k = np.arange(1, 9)
    assert np.allclose(y.real, np.cos(k * np.pi / 2), atol=1e-5) and np.allclose(y.imag, np.sin(k * np.pi / 2), atol=1e-5)
    assert abs(oracle.rotator_phase_increment(0.25, 1.0) - np.pi / 2) < 1e-6


def test_math_qa_math(oracle):
    """blocks/math/test/qa_Math.cpp:53-151: exact results on small integer-valued complex<float> vectors."""
    x = np.array([1, 2, 8, 17], dtype=np.complex64)
    assert np.array_equal(oracle.mathop_const("add", x, 2), [3, 4, 10, 19])
    assert np.array_equal(oracle.mathop_const("subtract", x, 2), [-1, 0, 6, 15])
    assert np.array_equal(oracle.mathop_const("multiply", x, 2), [2, 4, 16, 34])
    assert np.array_equal(oracle.mathop_const("divide", x, 2), [0.5, 1, 4, 8.5])
    a, b, c = np.array([1, 2, 3], np.complex64), np.array([4, 5, 6], np.complex64), np.array([8, 9, 10], np.complex64)
    assert np.array_equal(oracle.mathop_multi("add", [a, b, c]), [13, 16, 19])
    assert np.array_equal(oracle.mathop_multi("multiply", [a, b, c]), [32, 90, 180])
    assert np.array_equal(oracle.mathop_multi("subtract", [c, b, a]), [3, 2, 1])


def test_fft_block_qa_fourier(oracle):
    """blocks/fourier/test/qa_fourier.cpp:53-109: N = 256 sine at 0.1 fs; peak of the shifted magnitude within one bin."""
    n = 256
    x = np.sin(2 * np.pi * 0.1 * np.arange(n)).astype(np.complex64)
    sig, ranges = oracle.fft_block(x, n, oracle.window("Hann", n))
    peak = np.argmax(sig[0, 0])
    assert abs(abs(peak - n // 2) - 0.1 * n) <= 1.0
    assert np.allclose(ranges[0, :, 0], sig[0].min(axis=1)) and np.allclose(ranges[0, :, 1], sig[0].max(axis=1))
    X = sig[0, 2] + 1j * sig[0, 3]
    assert np.allclose(np.roll(sig[0, 0], -n // 2), np.abs(X) * 2 / n, atol=1e-6)


# ---- 2. golden fixtures generated from the compiled reference -------------------------------------------------------------
def test_golden_windows_and_design(oracle):
    with np.load(os.path.join(GOLDEN, "windows.npz")) as g:
        for key in g.files:
            name, n = key.rsplit("_", 1)
            assert bit_equal(oracle.window(name, int(n)), g[key]), key
    with np.load(os.path.join(GOLDEN, "fir_design.npz")) as g:
        assert bit_equal(oracle.fir_generate(127, "Hamming", 0.1), g["lowpass127_hamming_fc0p1"])
        assert bit_equal(oracle.fir_generate(127, "Hamming", 0.05), g["lowpass127_hamming_fc0p05"])
        for ftype in ("LOWPASS", "HIGHPASS", "BANDPASS", "BANDSTOP"):
            for win in ("Kaiser", "Hamming", "Hann"):
                assert bit_equal(oracle.fir_design(ftype, 4, 1.0, 10.0, 1000.0, window=win), g[f"design_{ftype}_{win}"]), (ftype, win)


def test_golden_fir(oracle):
    with np.load(os.path.join(GOLDEN, "fir.npz")) as g, np.load(os.path.join(GOLDEN, "fir_design.npz")) as d:
        x = crand(int(g["seed"]), int(g["n"]))
        assert bit_equal(oracle.fir(d["lowpass127_hamming_fc0p1"], x), g["y127"])
        assert bit_equal(oracle.fir(d["lowpass127_hamming_fc0p05"], x, decimate=8), g["y127_decim8"])
        assert bit_equal(oracle.fir(np.full(10, 0.1, dtype=np.float32), np.ones(32, dtype=np.float32)), g["step10"])
        for nt in (5, 33, 48, 200):
            assert bit_equal(oracle.fir(g[f"taps_{nt}"], x[:4096]), g[f"y_{nt}"]), nt


def test_golden_fft(oracle):
    with np.load(os.path.join(GOLDEN, "fft.npz")) as g:
        for n in (16, 64, 256, 1024, 4096, 8192):
            batch = int(g[f"batch_{n}"])
            x = crand(int(g["seed"]) + n, n * batch)
            X, want = oracle.fft(x, n), g[f"X_{n}"]
            for b in range(batch):
                sl = slice(b * n, (b + 1) * n)
                assert np.abs(X[sl] - want[sl]).max() <= FFT_TOL * np.linalg.norm(x[sl]), n
        xb = (crand(99, 8192) * 0.1 + np.exp(2j * np.pi * 0.1 * np.arange(8192))).astype(np.complex64)
        sig, ranges = oracle.fft_block(xb, 4096, oracle.window("Hann", 4096))
        want = g["block_signals"]
        scale = np.abs(want[:, 2:]).max()
        assert np.abs(sig[:, 2:] - want[:, 2:]).max() <= FFT_TOL * 64 * scale
        assert np.abs(sig[:, 0] - want[:, 0]).max() <= 1e-5 * want[:, 0].max()
        assert np.abs(ranges[:, 0] - g["block_ranges"][:, 0]).max() <= 1e-5 * want[:, 0].max()


def test_golden_math(oracle):
    with np.load(os.path.join(GOLDEN, "math.npz")) as g:
        xm = crand(77, 4096) * np.exp(np.random.default_rng(78).uniform(-20, 20, 4096)).astype(np.float32)
        for op in ("add", "subtract", "multiply", "divide"):
            assert bit_equal(oracle.mathop_const(op, xm, 0.37 - 1.91j), g[op]), op
        y, phase = oracle.rotator(crand(79, 65536), float(np.float32(2 * np.pi * 0.1)), 0.0)
        assert bit_equal(y[-4096:], g["rotator"]) and np.float32(phase) == g["rotator_end_phase"]


# ---- 3. live against the compiled reference ---------------------------------------------------------------------------------
def test_live_ref_windows_design_fir(oracle, ref):
    for name in _oracle.WINDOWS:
        for n in (1, 2, 8, 63, 1024):
            if name == "Kaiser" and n <= 1:
                continue
            assert np.array_equal(oracle.window(name, n), ref.window(name, n), equal_nan=True), (name, n)  # N = 1: 0/0 like the reference
            assert np.array_equal(oracle.window(name, n, dtype=np.float64), ref.window(name, n, dtype=np.float64), equal_nan=True), (name, n)
    for fc in (0.01, 0.1, 0.25, 0.4):
        for nt in (15, 127, 128, 3072):
            assert bit_equal(oracle.fir_generate(nt, "Kaiser", fc, beta=5.0), ref.fir_generate(nt, "Kaiser", fc, beta=5.0))
    x, xr = crand(5, 6000), np.random.default_rng(6).uniform(-1, 1, 6000).astype(np.float32)
    for nt in (1, 2, 16, 17, 31, 32, 33, 47, 48, 49, 64, 127, 128, 255, 256, 400):
        t = np.random.default_rng(nt).uniform(-1, 1, nt).astype(np.float32)
        assert bit_equal(oracle.fir(t, x), ref.fir(t, x)), nt
        assert bit_equal(oracle.fir(t, xr), ref.fir(t, xr)), nt
        for d in (2, 5, 8):
            assert bit_equal(oracle.fir(t, x[: 6000 // d * d], decimate=d), ref.fir(t, x[: 6000 // d * d], decimate=d)), (nt, d)
    td = np.random.default_rng(9).uniform(-1, 1, 127)
    assert np.array_equal(oracle.fir_f64(td, xr.astype(np.float64)), ref.fir_f64(td, xr.astype(np.float64)))


def test_live_ref_fft_and_block(oracle, ref):
    for n in (16, 32, 128, 256, 2048, 4096, 8192):
        x = crand(n, n * 2)
        X, R, D = oracle.fft(x, n), ref.fft(x, n), oracle.fft_f64(x, n)
        for b in range(2):
            sl = slice(b * n, (b + 1) * n)
            nrm = np.linalg.norm(x[sl])
            assert np.abs(X[sl] - D[sl]).max() <= FFT_TOL * nrm and np.abs(R[sl] - D[sl]).max() <= FFT_TOL * nrm, n
    x = (crand(1, 4096 * 3) * 0.1 + np.exp(2j * np.pi * 0.05 * np.arange(4096 * 3))).astype(np.complex64)
    w = oracle.window("Hann", 4096)
    a, ra = oracle.fft_block(x, 4096, w)
    b, rb = ref.fft_block(x, 4096, w)
    scale = np.abs(b[:, 2:]).max()
    assert np.abs(a[:, 2:] - b[:, 2:]).max() <= FFT_TOL * 64 * scale and np.abs(a[:, 0] - b[:, 0]).max() <= 1e-5 * b[:, 0].max()
    X = crand(3, 4096)
    assert bit_equal(oracle.magnitude(X, shift=True), ref.magnitude(X, shift=True)) and bit_equal(oracle.phase(X, deg=True, shift=True), ref.phase(X, deg=True, shift=True))
    assert bit_equal(oracle.magnitude(X, db=True), ref.magnitude(X, db=True)) and bit_equal(oracle.phase(X, unwrap=True), ref.phase(X, unwrap=True))


def test_live_ref_math_and_mixer(oracle, ref):
    x = crand(8, 20000) * np.exp(np.random.default_rng(8).uniform(-30, 30, 20000)).astype(np.float32)
    x[::97] = np.inf
    x[5::101] = complex(np.nan, 1)
    for op in ("add", "subtract", "multiply", "divide"):
        for v in (0.37 - 1.91j, 0j, complex(np.inf, 0), 2 + 0j):
            a, b = oracle.mathop_const(op, x, v), ref.mathop_const(op, x, v)
            same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a.view(np.float32)) & np.isnan(b.view(np.float32)))
            assert same.all(), (op, v)
        ins = [crand(s, 3000) + np.complex64(0.25) for s in (1, 2, 3)]
        assert bit_equal(oracle.mathop_multi(op, ins), ref.mathop_multi(op, ins))
    for dphi in (0.6283185, -0.6283185, 1.5707964, 3.0, 1e-3):
        a, pa = oracle.rotator(crand(4, 50000), dphi, 0.3)
        b, pb = ref.rotator(crand(4, 50000), dphi, 0.3)
        assert bit_equal(a, b) and np.float32(pa) == np.float32(pb)


def test_own_definitions_channelizer_and_resampler_against_float64(oracle):
    """PARITY UNPINNED blocks (no reference implementation): the oracle's own definitions agree with a direct float64
    evaluation of the stated formulas."""
    rng = np.random.default_rng(77)
    # resampler: y[m] = sum_k h[(m M) mod L + k L] x[floor(m M / L) - k]
    for interp, decim, n_taps in ((3, 2, 25), (2, 5, 17), (1, 1, 9)):
        taps = rng.uniform(-1, 1, n_taps).astype(np.float32)
        x = (rng.uniform(-1, 1, decim * 40) + 1j * rng.uniform(-1, 1, decim * 40)).astype(np.complex64)
        got = oracle.resampler(taps, interp, decim, x)
        per_phase = -(-n_taps // interp)
        want = np.zeros(got.size, dtype=np.complex128)
        for m in range(got.size):
            p, q = (m * decim) % interp, (m * decim) // interp
            for k in range(per_phase):
                if p + k * interp < n_taps and q - k >= 0:
                    want[m] += float(taps[p + k * interp]) * complex(x[q - k])
        assert np.abs(got - want).max() < 1e-5
    # equivalent to zero-stuffing, filtering and decimating
    interp, decim = 3, 2
    taps = rng.uniform(-1, 1, 30).astype(np.float32)
    x = (rng.uniform(-1, 1, 200) + 1j * rng.uniform(-1, 1, 200)).astype(np.complex64)
    up = np.zeros(x.size * interp, dtype=np.complex128)
    up[::interp] = x
    full = np.convolve(up, taps.astype(np.float64))[: up.size]
    assert np.abs(oracle.resampler(taps, interp, decim, x) - full[::decim]).max() < 1e-5


@pytest.mark.parametrize("nfft", [16, 256, 2048])
def test_fft_block_on_real_input_matches_the_compiled_reference(oracle, ref, nfft):
    """FFT<float> block (half spectrum): the oracle's restatement against the statements of processBulk / createDataset run
    on the reference's own FFT<float>, computeMagnitudeSpectrum and computePhaseSpectrum (oracle/_ref): same plane sizes
    (N/2), no fft-shift, Re / Im taken from the LAST N/2 bins of the spectrum."""
    rng = np.random.default_rng(nfft)
    x = rng.uniform(-1, 1, nfft * 7).astype(np.float32)
    w = oracle.window("Hann", nfft)
    for db, deg, unwrap in ((False, False, False), (True, True, False), (False, False, True)):
        a, ra = oracle.fft_block_real(x, nfft, w, db, deg, unwrap)
        b, rb = ref.fft_block_real(x, nfft, w, db, deg, unwrap)
        assert a.shape == b.shape == (7, 4, nfft // 2)
        scale = np.abs(b[:, 2:]).max()
        assert np.abs(a[:, 2:] - b[:, 2:]).max() <= 2e-6 * np.sqrt(nfft) * scale
        lin = ref.fft_block_real(x, nfft, w, want_ranges=False)
        strong = lin[:, 0] > 1e-3 * lin[:, 0].max()
        assert np.abs(a[:, 0] - b[:, 0])[strong].max() <= (1e-3 if db else 1e-5 * lin[:, 0].max())
        period = 360.0 if deg else 2 * np.pi
        d = np.abs((a[:, 1] - b[:, 1] + period / 2) % period - period / 2)
        assert d[strong].max() <= (0.2 if deg else 3e-3)
    # the structure itself: magnitude / phase are those of bins [0, N/2) of the full transform, Re / Im of bins [N/2, N)
    X = np.fft.fft((x[:nfft] * w).astype(np.float64))
    sig = ref.fft_block_real(x[:nfft], nfft, w, want_ranges=False)[0]
    assert np.allclose(sig[0], np.abs(X[: nfft // 2]) * 2 / nfft, atol=1e-5)
    assert np.allclose(sig[2], X[nfft // 2 :].real, atol=2e-4) and np.allclose(sig[3], X[nfft // 2 :].imag, atol=2e-4)


# ---- sample-format converters (blocks/basic/test/qa_Converter.cpp:242-268) -----------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.int16, np.int8])
def test_complex_interleaved_round_trip_golden(oracle, dtype):
    x = np.array([1 + 2j, 3 + 4j, 5 + 6j], dtype=np.complex64)
    items = oracle.complex_to_interleaved(x, dtype)
    assert items.dtype == dtype and items.tolist() == [1, 2, 3, 4, 5, 6]
    assert np.array_equal(oracle.interleaved_to_complex(items), x)


def test_complex_to_interleaved_truncates_toward_zero_and_is_position_independent(oracle):
    # static_cast semantics: truncation toward zero in range; out-of-range values come out the same wherever they sit in
    # the buffer (scalar and vectorised loop iterations of the compiled restatement agree)
    x = np.array([1.9 - 1.9j, -0.5 + 0.5j, 32767.99 - 32768.0j, 127.5 - 128.9j], dtype=np.complex64)
    assert oracle.complex_to_interleaved(x, np.int16).tolist() == [1, -1, 0, 0, 32767, -32768, 127, -128]
    assert oracle.complex_to_interleaved(x[[0, 1, 3]], np.int8).tolist() == [1, -1, 0, 0, 127, -128]
    rng = np.random.default_rng(3)
    wild = (rng.standard_normal(4099) * 1e5 + 1j * rng.standard_normal(4099) * 1e11).astype(np.complex64)
    wild[7] = np.nan + 1j * np.inf
    for dtype in (np.int16, np.int8):
        whole = oracle.complex_to_interleaved(wild, dtype)
        single = np.concatenate([oracle.complex_to_interleaved(wild[i : i + 1], dtype) for i in range(0, wild.size, 97)])
        assert np.array_equal(whole.reshape(-1, 2)[::97].ravel(), single)
