"""GPU parity tests: the CUDA path, called through the C ABI (gnuradio4_b200 -> libgr4b200.so), against the CPU oracle on
the same seeded inputs. Bit-exact for FIR / add / subtract / multiply / divide / decimate / mixer (phase AND output); a stated
tolerance for the FFT. Run on the B200 box: pytest -m gpu."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def gr4():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import gnuradio4_b200 as g

    g.load()  # raises if libgr4b200.so is missing: no fallback
    return g


def crandn(rng, n, scale=1.0):
    return (scale * (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n))).astype(np.complex64)


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def bits(x):
    x = np.ascontiguousarray(x)
    return x.view(np.uint32)


def assert_bit_equal(got, want, what=""):
    got, want = np.ascontiguousarray(got), np.ascontiguousarray(want)
    assert got.shape == want.shape, what
    same = bits(got) == bits(want)
    both_nan = np.isnan(got.view(np.float32)) & np.isnan(want.view(np.float32))
    bad = ~(same | both_nan)
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} words differ, first at {np.argmax(bad)}"


# ---- elementwise math (reference tests: blocks/math/test/qa_Math.cpp:53-151, exact equality) ---------------------------
@pytest.mark.parametrize("op,cls", [("add", "AddConst"), ("subtract", "SubtractConst"), ("multiply", "MultiplyConst"), ("divide", "DivideConst")])
@pytest.mark.parametrize("n", [0, 1, 2, 7, 4096, 100003])
def test_mathop_const_bit_exact(gr4, oracle, op, cls, n):
    rng = np.random.default_rng(11 + n)
    x = crandn(rng, n) * np.exp(rng.uniform(-30, 30, n)).astype(np.float32)
    value = 0.37 - 1.91j
    block = getattr(gr4, cls)(value=value)
    if n == 0:
        assert block.process_bulk(dev(x)).numel() == 0
        return
    got = block.process_bulk(dev(x)).cpu().numpy()
    assert_bit_equal(got, oracle.mathop_const(op, x, value), f"{cls} n={n}")


@pytest.mark.parametrize("op,cls", [("multiply", "MultiplyConst"), ("divide", "DivideConst"), ("add", "AddConst")])
def test_mathop_const_special_values(gr4, oracle, op, cls):
    specials = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e38, -1e38, 1e-45, 3.5], dtype=np.float32)
    re, im = np.meshgrid(specials, specials)
    x = (re + 1j * im).astype(np.complex64).ravel()
    for value in [2 + 0j, 0j, complex(np.inf, 1.0), 1e30 + 1e30j, complex(0.0, -3.0)]:
        got = getattr(gr4, cls)(value=value).process_bulk(dev(x)).cpu().numpy()
        want = oracle.mathop_const(op, x, value)
        # NaN payload/sign is not part of the contract; everything else is bit-exact
        assert_bit_equal(got, want, f"{cls} value={value}")


def test_mathop_const_qa_math_vectors(gr4):
    """qa_Math.cpp known answers: {1,2,8,17} op 2 on complex<float>."""
    x = np.array([1, 2, 8, 17], dtype=np.complex64)
    assert np.array_equal(gr4.AddConst(value=2).process_bulk(dev(x)).cpu().numpy(), np.array([3, 4, 10, 19], dtype=np.complex64))
    assert np.array_equal(gr4.SubtractConst(value=2).process_bulk(dev(x)).cpu().numpy(), np.array([-1, 0, 6, 15], dtype=np.complex64))
    assert np.array_equal(gr4.MultiplyConst(value=2).process_bulk(dev(x)).cpu().numpy(), np.array([2, 4, 16, 34], dtype=np.complex64))
    assert np.array_equal(gr4.DivideConst(value=2).process_bulk(dev(x)).cpu().numpy(), np.array([0.5, 1, 4, 8.5], dtype=np.complex64))


def test_mathop_misaligned_view(gr4, oracle):
    rng = np.random.default_rng(5)
    x = crandn(rng, 10001)
    xd = dev(x)
    got = gr4.MultiplyConst(value=1.5 - 2j).process_bulk(xd[1:].contiguous() if False else xd[1:]).cpu().numpy()  # 8-byte aligned only
    assert_bit_equal(got, oracle.mathop_const("multiply", x[1:], 1.5 - 2j))


@pytest.mark.parametrize("op,cls", [("add", "Add"), ("subtract", "Subtract"), ("multiply", "Multiply"), ("divide", "Divide")])
@pytest.mark.parametrize("n_inputs", [1, 2, 3, 5])
def test_mathop_multi_bit_exact(gr4, oracle, op, cls, n_inputs):
    rng = np.random.default_rng(23)
    ins = [crandn(rng, 5000) + np.complex64(0.1) for _ in range(n_inputs)]
    got = getattr(gr4, cls)(n_inputs=n_inputs).process_bulk([dev(i) for i in ins]).cpu().numpy()
    assert_bit_equal(got, oracle.mathop_multi(op, ins), f"{cls} x{n_inputs}")


def test_mathop_multi_limits(gr4):
    with pytest.raises(gr4.Gr4b200Error):
        gr4.Add(n_inputs=33)


def test_decimator(gr4, oracle):
    """qa_filter.cpp:267-293: 100 -> 10 samples with decim 10."""
    x = (np.arange(100) + 1j * np.arange(100)).astype(np.complex64)
    got = gr4.Decimator(decim=10).process_bulk(dev(x)).cpu().numpy()
    assert got.size == 10
    assert np.array_equal(got, oracle.decimate(x, 10))


# ---- mixer (reference test: blocks/math/test/qa_Rotator.cpp:69-92) ------------------------------------------------------
def max_ulp_distance(got, want):
    """largest distance in units of the last place between two float32 arrays (0 = bit-identical; NaN == NaN)"""
    g = np.ascontiguousarray(got).view(np.float32).astype(np.float64)
    w = np.ascontiguousarray(want).view(np.float32).astype(np.float64)
    both_nan = np.isnan(g) & np.isnan(w)
    ulp = np.spacing(np.abs(np.ascontiguousarray(want).view(np.float32))).astype(np.float64)
    d = np.where(both_nan, 0.0, np.abs(g - w) / ulp)
    return float(np.nanmax(d)) if d.size else 0.0


def assert_mixer_bits(got, want, what):
    """north_star: within 1 ulp for elementwise work. cos/sin are the C library's operation sequence on the FP64 pipe
    (csrc/sincos_core.cuh), the phase is the reference's recurrence, the product rounds like __mulsc3: the bar is 0 ulp."""
    print(f"{what}: max distance {max_ulp_distance(got, want)} ulp over {np.asarray(got).size} samples")
    assert_bit_equal(got, want, what)


def test_rotator_qa_known_answer(gr4):
    x = np.ones(8, dtype=np.complex64)
    got = gr4.Rotator(phase_increment=np.pi / 2).process_bulk(dev(x)).cpu().numpy()
    k = np.arange(1, 9)
    assert np.allclose(got.real, np.cos(k * np.pi / 2), atol=1e-5)
    assert np.allclose(got.imag, np.sin(k * np.pi / 2), atol=1e-5)
    rot = gr4.Rotator(sample_rate=1.0, frequency_shift=0.25)
    assert abs(rot.phase_increment - np.pi / 2) < 1e-6
    with pytest.raises(gr4.Gr4b200Error):
        gr4.Rotator(frequency_shift=0.1, phase_increment=0.1)


@pytest.mark.parametrize("dphi", [2 * np.pi * 0.1, -2 * np.pi * 0.1, np.pi / 2, 3.0, -3.1, 1e-3, 2 * np.pi / 4096, 0.0, 5.0])
@pytest.mark.parametrize("phi0", [0.0, 2.5])
def test_rotator_matches_reference_recurrence(gr4, oracle, dphi, phi0):
    rng = np.random.default_rng(31)
    n = 300000 if abs(dphi) <= np.pi and dphi != 0.0 else 20000  # serial fallback path for the out-of-range increments
    x = crandn(rng, n)
    rot = gr4.Rotator(phase_increment=dphi, initial_phase=phi0)
    xd = dev(x)
    # three calls of ragged sizes: the accumulated phase must carry across chunks exactly
    cuts = [0, 4097, n // 2 + 3, n]
    got = np.concatenate([rot.process_bulk(xd[a:b].clone()).cpu().numpy() for a, b in zip(cuts[:-1], cuts[1:])])
    want, end_phase = oracle.rotator(x, float(np.float32(dphi)), phi0)
    assert np.float32(rot.accumulated_phase) == np.float32(end_phase), "phase accumulator is not bit-identical"
    assert_mixer_bits(got, want, f"Rotator dphi={dphi} phi0={phi0}")


def test_rotator_per_call_lookup_path_still_matches(gr4, oracle):
    """The checkpoints normally come from the plan's phase cycle (rotator.cu findPhaseCycle); the landing-table look-ups are
    the fallback for recurrences that do not close. GR4B200_ROTATOR_CYCLE=0 forces the fallback (read once per process,
    hence a child process): mixer and fused DDC must give the same bits either way."""
    import subprocess
    import sys
    import textwrap

    code = textwrap.dedent("""
        import sys, numpy as np, torch
        sys.path.insert(0, %r)
        import gnuradio4_b200 as gr4
        rng = np.random.default_rng(3)
        x = (rng.uniform(-1, 1, 400000) + 1j * rng.uniform(-1, 1, 400000)).astype(np.complex64)
        xd = torch.from_numpy(x).cuda()
        taps = gr4.fir_generate(127, "Hamming", 0.05)
        out = []
        for dphi, phi0 in ((0.6283185, 0.0), (-1.9, 2.5), (1e-3, 6.0)):
            rot = gr4.Rotator(phase_increment=dphi, initial_phase=phi0)
            out.append(torch.cat([rot.process_bulk(xd[:123456]), rot.process_bulk(xd[123456:])]).cpu().numpy())
            ddc = gr4.DDC(gr4.Rotator(phase_increment=dphi, initial_phase=phi0), gr4.fir_filter(b=taps, decimate=8))
            out.append(torch.cat([ddc.process_bulk(xd[:80000]), ddc.process_bulk(xd[80000:])]).cpu().numpy())
        np.savez(sys.argv[1], *out)
    """) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    results = []
    for cycle in ("1", "0"):
        path = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"gr4b200_rotator_cycle_{cycle}_{os.getpid()}.npz")
        subprocess.run([sys.executable, "-c", code, path], check=True, env={**os.environ, "GR4B200_ROTATOR_CYCLE": cycle}, timeout=300)
        with np.load(path) as data:
            results.append([data[k] for k in data.files])
        os.remove(path)
    for a, b in zip(*results):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    x = np.random.default_rng(3)
    xs = (x.uniform(-1, 1, 400000) + 1j * x.uniform(-1, 1, 400000)).astype(np.complex64)
    want, _ = oracle.rotator(xs, float(np.float32(0.6283185)), 0.0)
    assert_mixer_bits(results[0][0], want, "phase-cycle path against the oracle")


@pytest.mark.parametrize("phi0", [-0.0, 1e-5, -3e-4, 0.7853, 100.0, -119.9, 121.0, 1e6, -3e9, 1e30, float("inf"), float("nan")])
def test_rotator_sincos_argument_ranges(gr4, oracle, phi0):
    """every branch of the library's sinf / cosf: tiny, below pi/4, one-step reduction, the 4/pi table above 120, inf, NaN"""
    rng = np.random.default_rng(33)
    x = crandn(rng, 5000)
    dphi = float(np.float32(1e-4))
    got = gr4.Rotator(phase_increment=dphi, initial_phase=phi0).process_bulk(dev(x)).cpu().numpy()
    want, _ = oracle.rotator(x, dphi, phi0)
    assert_mixer_bits(got, want, f"Rotator phi0={phi0}")


def test_rotator_every_phase_of_a_revolution(gr4, oracle):
    """2^23 consecutive phases of a slow mixer (every float between two wraps of a 2 pi / 2^22 increment is visited)"""
    n = 1 << 23
    dphi = float(np.float32(2 * np.pi / (1 << 22)))
    x = np.ones(n, dtype=np.complex64)
    x.imag = -0.5
    got = gr4.Rotator(phase_increment=dphi, initial_phase=0.0).process_bulk(dev(x)).cpu().numpy()
    want, _ = oracle.rotator(x, dphi, 0.0)
    assert_mixer_bits(got, want, "Rotator, one revolution in 2^22 steps")


def test_rotator_long_run_phase_is_bit_exact(gr4, oracle):
    """2^24 samples: the drifting float phase of the reference is reproduced bit for bit (closed forms would not)."""
    n = 1 << 24
    dphi = float(np.float32(2 * np.pi * 0.1))
    rot = gr4.Rotator(phase_increment=dphi)
    x = torch.ones(n, dtype=torch.complex64, device="cuda")
    y = rot.process_bulk(x)
    _, end_phase = oracle.rotator_phases(n, dphi, 0.0, want=False)
    assert np.float32(rot.accumulated_phase) == np.float32(end_phase)
    phases, _ = oracle.rotator_phases(4096, dphi, 0.0)
    head = y[:4096].cpu().numpy()
    assert np.allclose(head.real, np.cos(phases.astype(np.float64)), atol=3e-7)
    tail = y[n - 65536 :].cpu().numpy()  # the last 64 Ki outputs against the oracle run over the whole stream
    want_tail, _ = oracle.rotator(np.ones(n, dtype=np.complex64), dphi, 0.0)
    assert_mixer_bits(tail, want_tail[n - 65536 :], "Rotator after 2^24 samples")
    # drift check: the ideal phase n*dphi is far from the float recurrence by now
    ideal = (n * np.float64(dphi)) % (2 * np.pi)
    assert abs(ideal - end_phase) > 1e-3


# ---- FIR (reference tests: blocks/filter/test/qa_filter.cpp:54-93, 150-265) ---------------------------------------------
def test_fir_step_response_qa_filter(gr4):
    taps = np.full(10, 0.1, dtype=np.float32)
    x = np.ones(20, dtype=np.float32)
    y = gr4.fir_filter(b=taps).process_bulk(dev(x)).cpu().numpy()
    assert abs(y[0] - 0.1) < 1e-6 and abs(y[10] - 1.0) < 1e-3 and (np.abs(y[10:] - 1.0) < 1e-3).all()


@pytest.mark.parametrize("n_taps", [1, 2, 16, 32, 33, 47, 48, 64, 127, 128, 129, 255, 1000])
@pytest.mark.parametrize("decimate", [1, 2, 4, 8, 16, 5])
def test_fir_cf32_bit_exact(gr4, oracle, n_taps, decimate):
    rng = np.random.default_rng(1000 * n_taps + decimate)
    n = 3 * 4096 * decimate // decimate + 40 * decimate
    n = n // decimate * decimate
    x = crandn(rng, n)
    taps = rng.uniform(-1, 1, n_taps).astype(np.float32)
    got = gr4.fir_filter(b=taps, decimate=decimate).process_bulk(dev(x)).cpu().numpy()
    assert_bit_equal(got, oracle.fir(taps, x, decimate=decimate), f"fir taps={n_taps} D={decimate}")


@pytest.mark.parametrize("n_taps", [2, 33, 48, 127, 129, 255, 257, 1000])
def test_fir_cf32_bit_exact_long_calls(gr4, oracle, n_taps):
    """Full-rate calls of more than 74 tiles take the 256 x 16 tiles (the test above, like every short call, runs the
    quarter-size tiles); up to 256 taps travel as kernel parameters (uniform registers), longer filters as shared-memory
    tables."""
    rng = np.random.default_rng(31 * n_taps)
    n = 75 * 4096 + 777
    x = crandn(rng, n)
    taps = rng.uniform(-1, 1, n_taps).astype(np.float32)
    block = gr4.fir_filter(b=taps)
    xd = dev(x)
    got = torch.cat([block.process_bulk(xd[: n - 5000]), block.process_bulk(xd[n - 5000 :])]).cpu().numpy()  # long call, then a short one
    assert_bit_equal(got, oracle.fir(taps, x), f"fir taps={n_taps} long call")


@pytest.mark.parametrize("first,second,keeps", [(100, 127, True), (127, 33, True), (20, 32, True), (20, 40, False), (127, 129, False), (5, 300, False)])
@pytest.mark.parametrize("decimate", [1, 8])
def test_fir_new_coefficients_mid_stream_keep_the_history_like_the_reference(gr4, oracle, first, second, keeps, decimate):
    """fir_filter::settingsChanged (time_domain_filter.hpp:39-43) replaces -- and thereby zeroes -- the history buffer only
    when the new `b` is longer than its capacity (32 samples, then bit_ceil(b.size())); otherwise the first outputs after
    the change are formed from the OLD stream's samples. Same here, bit for bit (gr4b200_fir_plan_set_taps)."""
    rng = np.random.default_rng(first * 1000 + second)
    n1, n2 = 16 * 700, 16 * 900
    x = crandn(rng, n1 + n2)
    a, b = rng.uniform(-1, 1, first).astype(np.float32), rng.uniform(-1, 1, second).astype(np.float32)
    block = gr4.fir_filter(b=a, decimate=decimate)
    xd = dev(x)
    got1 = block.process_bulk(xd[:n1]).cpu().numpy()
    block.settings_changed(b=b)
    got2 = block.process_bulk(xd[n1:]).cpu().numpy()
    assert_bit_equal(got1, oracle.fir(a, x[:n1], decimate=decimate), "before the change")
    if keeps:  # the whole stream filtered by the new taps, seen from the second chunk on
        want2 = oracle.fir(b, x, decimate=decimate)[n1 // decimate :]
    else:  # a fresh history buffer: zeros in front of the second chunk
        want2 = oracle.fir(b, x[n1:], decimate=decimate)
    assert_bit_equal(got2, want2, f"after the change {first} -> {second} taps (history kept: {keeps})")


def test_fir_127_tap_lowpass_streaming_seams(gr4, oracle):
    """BASELINE config #2 at test size: designed 127-tap Hamming low-pass, ragged chunking, history across calls."""
    rng = np.random.default_rng(7)
    taps = gr4.fir_generate(127, "Hamming", 0.1)
    assert np.array_equal(taps, oracle.fir_generate(127, "Hamming", 0.1))
    n = 1 << 18
    x = crandn(rng, n)
    want = oracle.fir(taps, x)
    block = gr4.fir_filter(b=taps)
    xd = dev(x)
    cuts = [0, 1, 100, 127, 4096, 4097, 70001, 200000, n]
    got = np.concatenate([block.process_bulk(xd[a:b]).cpu().numpy() for a, b in zip(cuts[:-1], cuts[1:])])
    assert_bit_equal(got, want, "127-tap streaming")
    block.reset()
    assert_bit_equal(block.process_bulk(xd[:5000]).cpu().numpy(), want[:5000], "after reset")


def test_fir_decimating_streaming(gr4, oracle):
    rng = np.random.default_rng(8)
    taps = gr4.fir_generate(127, "Hamming", 0.05)
    n = 1 << 17
    x = crandn(rng, n)
    state = np.zeros(126 * 2, dtype=np.float32)
    block = gr4.fir_filter(b=taps, decimate=8)
    xd = dev(x)
    cuts = [0, 8, 4096, 4104, 65536 + 24, n]
    got, want = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        got.append(block.process_bulk(xd[a:b]).cpu().numpy())
        want.append(oracle.fir(taps, x[a:b], state=state, decimate=8))
    assert_bit_equal(np.concatenate(got), np.concatenate(want), "decimating streaming")
    with pytest.raises(gr4.Gr4b200Error):
        block.process_bulk(xd[:13])


def test_fir_real_stream_and_fast_mode(gr4, oracle):
    rng = np.random.default_rng(9)
    taps = gr4.fir_generate(127, "Hamming", 0.1)
    xr = rng.uniform(-1, 1, 50000).astype(np.float32)
    assert_bit_equal(gr4.fir_filter(b=taps).process_bulk(dev(xr)).cpu().numpy(), oracle.fir(taps, xr), "real stream")
    x = crandn(rng, 50000)
    fast = gr4.fir_filter(b=taps, exact=False).process_bulk(dev(x)).cpu().numpy()
    want = oracle.fir(taps, x)
    bound = 127 * 2.0**-24 * np.abs(taps).sum() * np.abs(x).max() * 2  # gamma_n * sum|b||x|
    assert np.abs(fast - want).max() <= bound


def test_fir_misaligned_and_special(gr4, oracle):
    rng = np.random.default_rng(10)
    taps = rng.uniform(-1, 1, 127).astype(np.float32)
    x = crandn(rng, 20001)
    x[100] = np.inf
    x[5000] = complex(np.nan, 1.0)
    x[7000] = -0.0
    xd = dev(x)
    got = gr4.fir_filter(b=taps).process_bulk(xd[1:]).cpu().numpy()
    assert_bit_equal(got, oracle.fir(taps, x[1:]), "misaligned / special values")


def test_fir_passband_stopband(gr4):
    """qa_filter.cpp:150-265 style: designed low-pass passes 0.05 fs, rejects 0.3 fs."""
    taps = gr4.fir_generate(127, "Hamming", 0.1)
    n = 20000
    t = np.arange(n)
    block = gr4.fir_filter(b=taps)
    lo = block.process_bulk(dev(np.exp(2j * np.pi * 0.05 * t).astype(np.complex64))).cpu().numpy()
    block.reset()
    hi = block.process_bulk(dev(np.exp(2j * np.pi * 0.3 * t).astype(np.complex64))).cpu().numpy()
    assert np.abs(lo[1000:]).min() > 0.9 and np.abs(hi[1000:]).max() < 0.01


# ---- FFT (reference tests: algorithm/test/qa_algorithm_fourier.cpp:67-143, qa_SimdFFT.cpp, blocks/fourier/test/qa_fourier.cpp)
FFT_TOL = 2.0e-6  # max_k |X_gpu[k] - X_f64[k]| <= FFT_TOL * ||x||_2 ; the compiled reference itself measures ~0.8e-6 at N=4096


@pytest.mark.parametrize("nfft", [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_fft_c2c_against_f64(gr4, oracle, nfft):
    rng = np.random.default_rng(nfft)
    batch = 5
    x = crandn(rng, nfft * batch)
    got = gr4.FFT(fftSize=nfft).compute(dev(x)).cpu().numpy()
    want = oracle.fft_f64(x, nfft)
    for b in range(batch):
        sl = slice(b * nfft, (b + 1) * nfft)
        err = np.abs(got[sl] - want[sl]).max() / np.linalg.norm(x[sl])
        assert err <= FFT_TOL, f"N={nfft} transform {b}: {err}"


@pytest.mark.parametrize("nfft", [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072, 262144])
def test_fft_many_transforms_persistent_loop_and_ragged_tail(gr4, oracle, nfft):
    """Enough transforms that every persistent CTA loops many times, plus a ragged tail (batch not a multiple of the
    transforms a CTA holds): windowed spectrum against a float64 transform of the same input (torch.fft in complex128 as
    the checker), and the block-mode planes against that spectrum, on every transform."""
    batch = (1 << 23) // nfft + 3
    g = torch.Generator(device="cuda")
    g.manual_seed(nfft)
    x = torch.empty(batch * nfft, dtype=torch.complex64, device="cuda")
    torch.view_as_real(x).uniform_(-1.0, 1.0, generator=g)
    block = gr4.FFT(fftSize=nfft, window="Hann")
    w = torch.from_numpy(oracle.window("Hann", nfft)).cuda()
    xw = (x.view(batch, nfft) * w).to(torch.complex64)  # the float product the kernel forms (fft.hpp:155-162)
    want = torch.fft.fft(xw.to(torch.complex128), dim=1)
    norm = torch.linalg.vector_norm(xw.to(torch.complex128), dim=1, keepdim=True)
    got = block.compute(x, windowed=True).view(batch, nfft)
    err = ((got.to(torch.complex128) - want).abs() / norm).max().item()
    assert err <= FFT_TOL, f"N={nfft}: spectrum error {err}"
    sig, ranges = block.process_bulk(x, want_ranges=True)
    assert torch.equal(sig[:, 2, :], got.real) and torch.equal(sig[:, 3, :], got.imag), "Re/Im planes are the spectrum itself"
    mag = torch.roll(want.abs() * (2.0 / nfft), nfft // 2, dims=1)
    assert ((sig[:, 0, :].double() - mag).abs().max() / mag.max()).item() < 1e-5
    strong = mag > 1e-3 * mag.max()
    dphi = torch.angle(torch.exp(1j * (sig[:, 1, :].double() - torch.roll(torch.angle(want), nfft // 2, dims=1))))
    assert dphi[strong].abs().max().item() < 2e-3
    assert torch.equal(ranges[:, :, 0], sig.amin(dim=2)) and torch.equal(ranges[:, :, 1], sig.amax(dim=2)), "per-signal min / max"


def test_fft_large_sizes_against_f64_and_slicing(gr4, oracle):
    """N > 8192 runs as two passes of column transforms through a scratch buffer that holds one slice of transforms
    (2^26 samples): the float64 oracle on a few transforms, a batch that spans slices against the same transforms alone,
    and the block's unwrapped-phase and dB / degree variants against the small-size kernel's definition of them."""
    nfft = 65536
    rng = np.random.default_rng(65)
    x = crandn(rng, nfft * 3)
    got = gr4.FFT(fftSize=nfft).compute(dev(x)).cpu().numpy()
    want = oracle.fft_f64(x, nfft)
    for b in range(3):
        sl = slice(b * nfft, (b + 1) * nfft)
        assert np.abs(got[sl] - want[sl]).max() / np.linalg.norm(x[sl]) <= FFT_TOL
    batch = (1 << 26) // nfft + 5  # one full slice plus a ragged second one
    g = torch.Generator(device="cuda")
    g.manual_seed(99)
    xs = torch.empty(batch * nfft, dtype=torch.complex64, device="cuda")
    torch.view_as_real(xs).uniform_(-1.0, 1.0, generator=g)
    fft = gr4.FFT(fftSize=nfft)
    whole = fft.compute(xs).view(batch, nfft)
    for b in (0, batch - 6, batch - 5, batch - 1):
        alone = fft.compute(xs[b * nfft : (b + 1) * nfft].clone())
        assert torch.equal(whole[b], alone), f"transform {b} depends on its position in the batch"
    del whole
    # the block's planes with every post-processing flag against the oracle
    tone = (0.7 * np.exp(2j * np.pi * 1234.0 * np.arange(nfft) / nfft)).astype(np.complex64)
    frames = np.concatenate([tone + 0.01 * x[:nfft], x[nfft : 2 * nfft]])
    window = oracle.window("Hann", nfft)
    lin = oracle.fft_block(frames, nfft, window, want_ranges=False)
    strong = lin[:, 0] > 1e-3 * lin[:, 0].max()
    for db, deg, unwrap in [(False, False, False), (True, True, False), (False, False, True)]:
        block = gr4.FFT(fftSize=nfft, window="Hann", outputInDb=db, outputInDeg=deg, unwrapPhase=unwrap)
        sig, ranges = block.process_bulk(dev(frames), want_ranges=True)
        sig, ranges = sig.cpu().numpy(), ranges.cpu().numpy()
        want_sig, _ = oracle.fft_block(frames, nfft, window, db=db, deg=deg, unwrap=unwrap)
        assert np.abs(sig[:, 2:] - want_sig[:, 2:]).max() <= FFT_TOL * np.sqrt(nfft) * np.abs(want_sig[:, 2:]).max()
        if db:
            assert np.abs(sig[:, 0] - want_sig[:, 0])[strong].max() <= 1e-3
        else:
            assert np.abs(sig[:, 0] - want_sig[:, 0]).max() <= 1e-5 * want_sig[:, 0].max() + 1e-7
        period = 360.0 if deg else 2 * np.pi
        d = np.abs((sig[:, 1] - want_sig[:, 1] + period / 2) % period - period / 2)
        assert d[strong].max() <= (0.2 if deg else 3e-3)
        assert np.array_equal(ranges[:, :, 0], sig.min(axis=2)) and np.array_equal(ranges[:, :, 1], sig.max(axis=2))


@pytest.mark.parametrize("nfft", [256, 1024, 4096, 8192])
def test_fft_input_that_is_only_8_byte_aligned(gr4, oracle, nfft):
    """Bulk staging needs 16-byte aligned sources; an edge span that starts on an odd sample takes the direct-load
    variant of the same kernel and must give the very same bits."""
    rng = np.random.default_rng(nfft + 5)
    batch = 37
    x = crandn(rng, nfft * batch + 1)
    f = gr4.FFT(fftSize=nfft, window="Hann")
    xd = dev(x)
    aligned = xd[1:].clone()  # same samples, 16-byte aligned copy
    assert xd[1:].data_ptr() % 16 == 8 and aligned.data_ptr() % 16 == 0
    got, want = f.compute(xd[1:], windowed=True), f.compute(aligned, windowed=True)
    assert torch.equal(torch.view_as_real(got).view(torch.int32), torch.view_as_real(want).view(torch.int32))
    sig_a, sig_b = f.process_bulk(xd[1:]), f.process_bulk(aligned)
    assert torch.equal(sig_a.view(torch.int32), sig_b.view(torch.int32))
    ref = oracle.fft_f64((x[1:].reshape(batch, nfft) * oracle.window("Hann", nfft)).astype(np.complex64).ravel(), nfft)
    assert np.abs(got.cpu().numpy() - ref).max() <= FFT_TOL * np.linalg.norm(x[1 : 1 + nfft]) * 2


def test_fft_block_eight_tones_peak_bins_and_amplitudes(gr4):
    """eight tones on top of noise, Hann window on: every tone shows up in its (fft-shifted) bin of
    the magnitude plane with amplitude * coherent gain (0.5 for Hann), in every transform of the batch."""
    nfft, batch = 4096, 64
    rng = np.random.default_rng(8)
    bins = np.array([5, 100, 777, 1500, 2047, 2500, 3333, 4000])
    amps = np.linspace(0.5, 4.0, 8)
    n = np.arange(nfft * batch)
    x = 1e-3 * (rng.standard_normal(n.size) + 1j * rng.standard_normal(n.size))
    for k, a in zip(bins, amps):
        x = x + a * np.exp(2j * np.pi * k * n / nfft)
    sig = gr4.FFT(fftSize=nfft, window="Hann").process_bulk(dev(x.astype(np.complex64))).cpu().numpy()
    shifted = (bins + nfft // 2) % nfft  # the block rotates the spectrum so that negative frequencies come first
    for b in range(batch):
        mag = sig[b, 0]
        # magnitude = |X| * 2 / N; a bin-centred tone of amplitude a under a Hann window gives a in its bin, a / 2 next to it
        assert np.allclose(mag[shifted], amps, rtol=2e-3), f"transform {b}"
        assert np.allclose(mag[(shifted + 1) % nfft], amps / 2, rtol=1e-2) and np.allclose(mag[(shifted - 1) % nfft], amps / 2, rtol=1e-2)
        rest = np.ones(nfft, dtype=bool)
        for d in (-1, 0, 1):
            rest[(shifted + d) % nfft] = False
        assert mag[rest].max() < 0.01, "nothing but the noise floor away from the tones"


@pytest.mark.parametrize("nfft", [16, 64, 256, 1024, 4096, 8192, 16384, 262144])
def test_fft_real_input_full_spectrum(gr4, oracle, nfft):
    """gr::algorithm::FFT<float>::compute (fft.hpp:214-258): real samples in, the full N-bin spectrum out; DC and Nyquist
    real, upper half the conjugate mirror (within rounding), everything within the FFT tolerance of a float64 transform."""
    rng = np.random.default_rng(nfft + 17)
    batch = 41
    x = rng.uniform(-1, 1, nfft * batch).astype(np.float32)
    f = gr4.FFT(fftSize=nfft, window="Hann")
    got = f.compute_real(dev(x)).cpu().numpy().reshape(batch, nfft)
    want = np.fft.fft(x.astype(np.float64).reshape(batch, nfft), axis=1)
    norm = np.linalg.norm(x.reshape(batch, nfft), axis=1, keepdims=True)
    assert (np.abs(got - want) / norm).max() <= FFT_TOL
    assert np.all(got[:, 0].imag == 0) and np.all(got[:, nfft // 2].imag == 0)
    assert (np.abs(got[:, 1 : nfft // 2] - np.conj(got[:, : nfft // 2 : -1])) / norm).max() <= FFT_TOL
    as_complex = f.compute(dev(x.astype(np.complex64))).cpu().numpy().reshape(batch, nfft)
    assert np.abs(got[:, 1 : nfft // 2] - as_complex[:, 1 : nfft // 2]).max() == 0  # the very same kernel arithmetic
    windowed = f.compute_real(dev(x), windowed=True).cpu().numpy().reshape(batch, nfft)
    w = oracle.window("Hann", nfft)
    want_w = np.fft.fft((x.reshape(batch, nfft) * w).astype(np.float32).astype(np.float64), axis=1)
    assert (np.abs(windowed - want_w) / norm).max() <= FFT_TOL
    # qa_algorithm_fourier.cpp:67-95 style peak check: a real sine at bin 5 shows -N/2 in Im X[5] and +N/2 in Im X[N-5]
    sine = np.sin(2 * np.pi * 5 * np.arange(nfft) / nfft).astype(np.float32)
    X = f.compute_real(dev(sine)).cpu().numpy()
    assert abs(X[5].imag + nfft / 2) < 1e-3 * nfft and abs(X[nfft - 5].imag - nfft / 2) < 1e-3 * nfft


@pytest.mark.parametrize("nfft", [16, 128, 256, 1024, 4096, 16384, 131072])
@pytest.mark.parametrize("db,deg,unwrap", [(False, False, False), (True, True, False), (False, False, True)])
def test_fft_block_on_a_real_stream(gr4, oracle, nfft, db, deg, unwrap):
    """FFT<float> (fft.hpp:147-250, half spectrum): planes of N/2 values -- magnitude and phase of bins [0, N/2) without the
    fft-shift, Re / Im of the last N/2 bins of the spectrum -- against the oracle's restatement (pinned to the compiled
    reference in tests/test_oracle.py)."""
    rng = np.random.default_rng(nfft + 3 * db + unwrap)
    batch = 23
    x = rng.uniform(-1, 1, nfft * batch).astype(np.float32)
    x[:nfft] += 2.0 * np.sin(2 * np.pi * 3 * np.arange(nfft) / nfft).astype(np.float32)  # a strong line in the first chunk
    block = gr4.FFT(fftSize=nfft, window="Hann", outputInDb=db, outputInDeg=deg, unwrapPhase=unwrap)
    got, got_ranges = block.process_bulk_real(dev(x), want_ranges=True)
    got, got_ranges = got.cpu().numpy(), got_ranges.cpu().numpy()
    want, _ = oracle.fft_block_real(x, nfft, oracle.window("Hann", nfft), db=db, deg=deg, unwrap=unwrap)
    assert got.shape == (batch, 4, nfft // 2)
    scale = np.abs(want[:, 2:]).max()
    assert np.abs(got[:, 2:] - want[:, 2:]).max() <= FFT_TOL * np.sqrt(nfft) * scale
    lin = oracle.fft_block_real(x, nfft, oracle.window("Hann", nfft), want_ranges=False)
    strong = lin[:, 0] > 1e-3 * lin[:, 0].max()
    if db:
        assert np.abs(got[:, 0] - want[:, 0])[strong].max() <= 1e-3
    else:
        assert np.abs(got[:, 0] - want[:, 0]).max() <= 1e-5 * want[:, 0].max() + 1e-7
    if not unwrap:
        period = 360.0 if deg else 2 * np.pi
        d = np.abs((got[:, 1] - want[:, 1] + period / 2) % period - period / 2)
        assert d[strong].max() <= (0.2 if deg else 3e-3)
    else:
        steps = np.abs(np.diff(got[:, 1], axis=1))
        assert steps.max() <= np.pi + 1e-3  # unwrapped: no jump larger than pi between neighbouring bins
        d = np.abs((got[:, 1] - want[:, 1] + np.pi) % (2 * np.pi) - np.pi)
        assert d[strong].max() <= 3e-3
    assert np.array_equal(got_ranges[:, :, 0], got.min(axis=2)) and np.array_equal(got_ranges[:, :, 1], got.max(axis=2))


def test_fft_pattern_known_answers(gr4):
    """qa_algorithm_fourier.cpp:97-143 (N = 16) and bm_fft.cpp:61-62 (sine at bin 5 => Im X[5] = -N/2)."""
    fft16 = gr4.FFT(fftSize=16)
    for signal, x0, amp in [(np.ones(16), 16 + 0j, 2.0), (np.ones(16) * (1 + 1j), 16 + 16j, np.sqrt(8.0)), (np.arange(1, 17), 136 + 0j, 17.0), (np.arange(16) % 2, 8 + 0j, 1.0)]:
        X = fft16.compute(dev(signal.astype(np.complex64))).cpu().numpy()
        assert abs(X[0] - x0) < 1e-5 * 16
        mag = np.abs(X) * 2 / 16
        assert np.argmax(mag) == 0 and abs(mag[0] - amp) < 1e-5
    n = 4096
    X = gr4.FFT(fftSize=n).compute(dev(np.sin(2 * np.pi * 5 * np.arange(n) / n).astype(np.complex64))).cpu().numpy()
    assert abs(X[5].imag + n / 2) < 0.1 and np.argmax(np.abs(X[: n // 2])) == 5


def test_fft_linearity_and_roundtrip_property(gr4):
    """qa_SimdFFT.cpp:131,421: linearity < 1e-4 N; Parseval as the size-independent property at full batch."""
    rng = np.random.default_rng(3)
    n, batch = 4096, 64
    a, b = crandn(rng, n * batch), crandn(rng, n * batch)
    f = gr4.FFT(fftSize=n)
    Fa, Fb, Fab = (f.compute(dev(v)).cpu().numpy() for v in (a, b, (2 * a + 3 * b).astype(np.complex64)))
    assert np.abs(Fab - (2 * Fa + 3 * Fb)).max() < 1e-4 * n
    for k in range(batch):
        sl = slice(k * n, (k + 1) * n)
        assert abs(np.vdot(Fa[sl], Fa[sl]).real / n - np.vdot(a[sl], a[sl]).real) < 1e-4 * n


@pytest.mark.parametrize("nfft", [256, 4096, 1024])
@pytest.mark.parametrize("db,deg", [(False, False), (True, True)])
def test_fft_block_signals(gr4, oracle, nfft, db, deg):
    """FFT block DataSet planes vs the oracle (tolerances of blocks/fourier/test/qa_fourier.cpp:77-94: 1e-4)."""
    rng = np.random.default_rng(nfft + db)
    batch = 3
    t = np.arange(nfft * batch)
    x = (crandn(rng, nfft * batch) * 0.1 + np.exp(2j * np.pi * 0.1 * t)).astype(np.complex64)
    block = gr4.FFT(fftSize=nfft, window="Hann", outputInDb=db, outputInDeg=deg)
    sig, ranges = block.process_bulk(dev(x), want_ranges=True)
    sig, ranges = sig.cpu().numpy(), ranges.cpu().numpy()
    want, want_ranges = oracle.fft_block(x, nfft, oracle.window("Hann", nfft), db=db, deg=deg)
    scale = np.abs(want[:, 2:]).max()
    assert np.abs(sig[:, 2] - want[:, 2]).max() <= FFT_TOL * np.sqrt(nfft) * scale and np.abs(sig[:, 3] - want[:, 3]).max() <= FFT_TOL * np.sqrt(nfft) * scale
    if db:
        strong = want[:, 0] > want[:, 0].max() - 60
        assert np.abs(sig[:, 0] - want[:, 0])[strong].max() < 1e-2
    else:
        assert np.abs(sig[:, 0] - want[:, 0]).max() <= 1e-5 * want[:, 0].max() + 1e-7
    # phase only where the bin is well above the rounding floor
    mag_lin = np.hypot(want[:, 2], want[:, 3])
    mag_shift = np.roll(mag_lin, nfft // 2, axis=1)
    strong = mag_shift > 1e-3 * mag_lin.max()
    dphi = np.abs(sig[:, 1] - want[:, 1])[strong]
    period = 360.0 if deg else 2 * np.pi
    dphi = np.minimum(dphi, np.abs(dphi - period))
    assert dphi.max() < (1e-2 if deg else 2e-4)
    assert np.abs(ranges[:, 2:] - want_ranges[:, 2:]).max() <= FFT_TOL * np.sqrt(nfft) * scale
    assert np.array_equal(block.frequency_axis()[[0, nfft // 2]], np.array([-0.5, 0.0], dtype=np.float32))


def test_fft_block_unwrap(gr4, oracle):
    rng = np.random.default_rng(77)
    nfft = 256
    x = (np.exp(2j * np.pi * 0.05 * np.arange(nfft)) * 5 + 5).astype(np.complex64)  # strong, smooth spectrum
    sig = gr4.FFT(fftSize=nfft, window="None", unwrapPhase=True).process_bulk(dev(x)).cpu().numpy()
    want = oracle.fft_block(x, nfft, oracle.window("None", nfft), unwrap=True, want_ranges=False)
    mag_shift = np.roll(np.hypot(want[0, 2], want[0, 3]), nfft // 2)
    assert sig.shape == want.shape and np.isfinite(sig).all()
    # unwrapped phase accumulates data-dependent 2 pi jumps at noise-level bins; compare modulo 2 pi on strong bins
    strong = mag_shift > 1e-2 * mag_shift.max()
    d = (sig[0, 1] - want[0, 1])[strong]
    assert np.abs(d - 2 * np.pi * np.round(d / (2 * np.pi))).max() < 1e-3


def test_fft_rejects_bad_sizes(gr4):
    with pytest.raises(gr4.Gr4b200Error):
        gr4.FFT(fftSize=0)
    with pytest.raises(gr4.Gr4b200Error):
        gr4.FFT(fftSize=131073)  # not a power of two and beyond the chirp-z limit
    with pytest.raises(gr4.Gr4b200Error):
        gr4.FFT(fftSize=4096).compute(torch.zeros(100, dtype=torch.complex64, device="cuda"))


BLUESTEIN_TOL = 4.0e-6  # max |X - X_f64| / ||x||_2 for the chirp-z sizes (three transforms and two chirp products deep)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 12, 17, 96, 100, 160, 1000, 1009, 4095, 4097, 6000, 10007, 50000, 100003, 131072 - 1])
def test_fft_any_size_against_float64(gr4, n):
    """gr::algorithm::FFT::compute accepts every size (fft.hpp:113-153: SimdFFT for {2,3,4,5}-smooth multiples of 16,
    Bluestein otherwise; bm_fft.cpp times N = 1009). Here: chirp-z over the power-of-two kernels, natural order,
    unnormalised forward DFT."""
    rng = np.random.default_rng(n)
    batch = 3 if n <= 10007 else 1
    x = crandn(rng, n * batch)
    got = gr4.FFT(fftSize=n).compute(dev(x)).cpu().numpy().reshape(batch, n)
    want = np.fft.fft(x.astype(np.complex128).reshape(batch, n), axis=1)
    for b in range(batch):
        err = np.abs(got[b] - want[b]).max() / max(np.linalg.norm(x.reshape(batch, n)[b]), 1e-30)
        assert err <= BLUESTEIN_TOL, f"N={n}: {err}"


@pytest.mark.parametrize("n", [1000, 1009, 96, 250])
def test_fft_block_any_size_against_oracle(gr4, oracle, n):
    """The FFT block at sizes that are not a power of two: window, planes, fft-shift by floor(N/2) (std::rotate), ranges."""
    rng = np.random.default_rng(7 * n)
    k = np.arange(3 * n)
    x = (crandn(rng, 3 * n) * 0.05 + np.exp(2j * np.pi * 0.123 * k)).astype(np.complex64)
    sig, ranges = gr4.FFT(fftSize=n, window="Hann").process_bulk(dev(x), want_ranges=True)
    sig, ranges = sig.cpu().numpy(), ranges.cpu().numpy()
    want, want_ranges = oracle.fft_block(x, n, oracle.window("Hann", n))
    scale = np.abs(want[:, 2:]).max()
    assert np.abs(sig[:, 2:] - want[:, 2:]).max() <= BLUESTEIN_TOL * 64 * scale
    assert np.abs(sig[:, 0] - want[:, 0]).max() <= 2e-5 * want[:, 0].max()
    strong = np.roll(np.hypot(want[:, 2], want[:, 3]), -(n - n // 2), axis=1) > 1e-2 * scale  # bin k sits at (k + n - n/2) mod n
    d = np.angle(np.exp(1j * (sig[:, 1] - want[:, 1])))
    assert np.abs(d[strong]).max() <= 1e-3
    assert np.abs(ranges[:, 0] - want_ranges[:, 0]).max() <= 2e-5 * want[:, 0].max()
    peak = int(np.argmax(sig[0, 0]))
    assert abs(peak - (int(round(0.123 * n)) + n // 2) % n) <= 1  # the tone, fft-shifted


def test_fft_real_input_any_size(gr4, oracle):
    """FFT<float> at N = 1000: full Hermitian spectrum from compute_real, half-spectrum planes from the block."""
    n = 1000
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, 2 * n).astype(np.float32)
    X = gr4.FFT(fftSize=n).compute_real(dev(x)).cpu().numpy().reshape(2, n)
    want = np.fft.fft(x.astype(np.float64).reshape(2, n), axis=1)
    assert np.abs(X - want).max() <= BLUESTEIN_TOL * np.linalg.norm(x[:n]) * 1.5
    assert (X[:, 0].imag == 0).all() and (X[:, n // 2].imag == 0).all()
    sig = gr4.FFT(fftSize=n, window="Hann").process_bulk_real(dev(x)).cpu().numpy()
    want_sig = oracle.fft_block_real(x, n, oracle.window("Hann", n), want_ranges=False)
    assert sig.shape == want_sig.shape == (2, 4, n // 2)
    assert np.abs(sig[:, 2:] - want_sig[:, 2:]).max() <= BLUESTEIN_TOL * 64 * np.abs(want_sig[:, 2:]).max()
    assert np.abs(sig[:, 0] - want_sig[:, 0]).max() <= 2e-5 * want_sig[:, 0].max()


# ---- chains --------------------------------------------------------------------------------------------------------------
def test_fir_to_fft_flowgraph_against_oracle(gr4, oracle):
    """The north-star flowgraph at test size through Graph / Simple with host buffers (H2D, rings, D2H inside)."""
    rng = np.random.default_rng(99)
    nfft, n = 4096, 4096 * 40
    taps = gr4.fir_generate(127, "Hamming", 0.1)
    g = gr4.Graph()
    fir = g.emplaceBlock(gr4.fir_filter, b=taps, compute_domain="gpu:cuda:0")
    fft = g.emplaceBlock(gr4.FFT, fftSize=nfft, window="Hann", compute_domain="gpu:cuda:0")
    assert g.connect(fir, fft)
    sched = gr4.Simple(g, chunk_items=4096 * 8)
    src = gr4.HostBuffer(n, np.complex64)
    dst = gr4.HostBuffer(n // nfft * 4 * nfft, np.float32)
    x = crandn(rng, n)
    src.array[:] = x
    nbytes = sched.runAndWait(src.array, dst.array)
    assert nbytes == dst.nbytes and sched.launches == 2 * 5
    got = dst.array.reshape(-1, 4, nfft).copy()
    y = oracle.fir(taps, x)
    want = oracle.fft_block(y, nfft, oracle.window("Hann", nfft), want_ranges=False)
    scale = np.abs(want[:, 2:]).max()
    assert np.abs(got[:, 2:] - want[:, 2:]).max() <= FFT_TOL * np.sqrt(nfft) * scale
    assert np.abs(got[:, 0] - want[:, 0]).max() <= 1e-5 * want[:, 0].max() + 1e-7
    sched.close()


@pytest.mark.parametrize("exact", [True, False])
@pytest.mark.parametrize("n_taps", [127, 33, 2, 255])
def test_fused_fir_fft_equals_the_two_blocks(gr4, oracle, n_taps, exact):
    """Merge<fir_filter, FFT> as one kernel: the very bits of the two kernels back to back, over several calls (the FIR
    history carries over), with and without a 16-byte aligned input; and against the oracle for the exact 127-tap case."""
    rng = np.random.default_rng(900 + n_taps)
    nfft, frames = 4096, 37
    taps = gr4.fir_generate(n_taps, "Hamming", 0.1) if n_taps > 2 else np.array([0.75, -0.25], dtype=np.float32)
    x = crandn(rng, nfft * frames + 1)
    for offset in (0, 1):  # offset 1: the input is only 8-byte aligned -> cooperative staging instead of bulk copies
        xd = dev(x)[offset : offset + nfft * frames]
        fused = gr4.FirFft(gr4.fir_filter(b=taps, exact=exact), gr4.FFT(fftSize=nfft, window="Hann"))
        fir, fft = gr4.fir_filter(b=taps, exact=exact), gr4.FFT(fftSize=nfft, window="Hann")
        cuts = [0, nfft * 5, nfft * 6, nfft * frames]
        got = torch.cat([fused.process_bulk(xd[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]).cpu().numpy()
        want = torch.cat([fft.process_bulk(fir.process_bulk(xd[a:b])) for a, b in zip(cuts[:-1], cuts[1:])]).cpu().numpy()
        assert_bit_equal(got, want, f"fused FIR->FFT taps={n_taps} exact={exact} offset={offset}")
    if exact and n_taps == 127:
        y = oracle.fir(taps, x[1 : 1 + nfft * frames])
        ref = oracle.fft_block(y, nfft, oracle.window("Hann", nfft), want_ranges=False)
        scale = np.abs(ref[:, 2:]).max()
        assert np.abs(got[:, 2:] - ref[:, 2:]).max() <= FFT_TOL * np.sqrt(nfft) * scale
        assert np.abs(got[:, 0] - ref[:, 0]).max() <= 1e-5 * ref[:, 0].max() + 1e-7
    with pytest.raises(gr4.Gr4b200Error):
        gr4.FirFft(gr4.fir_filter(b=taps, decimate=2), gr4.FFT(fftSize=nfft))
    with pytest.raises(gr4.Gr4b200Error):
        gr4.FirFft(gr4.fir_filter(b=taps), gr4.FFT(fftSize=1024))


def test_ddc_chain_against_oracle(gr4, oracle):
    """BASELINE config #4 at test size: Rotator -> decimating FIR (x8) -> FFT 4096."""
    rng = np.random.default_rng(4)
    n = 4096 * 8 * 6
    x = crandn(rng, n)
    dphi = float(np.float32(2 * np.pi * 0.1))
    taps = gr4.fir_generate(127, "Hamming", 0.05)
    mixer, fir = gr4.Rotator(phase_increment=dphi), gr4.fir_filter(b=taps, decimate=8)
    y = gr4.DDC(mixer, fir).process_bulk(dev(x))
    X = gr4.FFT(fftSize=4096).compute(y).cpu().numpy()
    mixed, _ = oracle.rotator(x, dphi)
    want_y = oracle.fir(taps, mixed, decimate=8)
    assert_mixer_bits(y.cpu().numpy(), want_y, "DDC (mixer -> FIR /8) against the oracle chain")
    want_X = oracle.fft_f64(want_y, 4096)
    for b in range(6):
        sl = slice(b * 4096, (b + 1) * 4096)
        assert np.abs(X[sl] - want_X[sl]).max() <= 2 * FFT_TOL * np.linalg.norm(want_y[sl]) + 1e-6


@pytest.mark.parametrize("decimate", [2, 4, 8, 16, 5])
@pytest.mark.parametrize("exact", [True, False])
def test_ddc_fused_equals_the_two_blocks_back_to_back(gr4, decimate, exact):
    """gr4b200_ddc_cf32 (mixer rotated in shared memory inside the decimating FIR kernel) must give, bit for bit and across
    ragged chunk seams, what Rotator -> fir_filter give as separate device calls (decimate = 5 takes the unfused route)."""
    rng = np.random.default_rng(40 + decimate)
    n = decimate * 16 * 1500
    x = dev(crandn(rng, n))
    dphi = float(np.float32(2 * np.pi * 0.0731))
    taps = gr4.fir_generate(127, "Hamming", 0.05)
    fused = gr4.DDC(gr4.Rotator(phase_increment=dphi), gr4.fir_filter(b=taps, decimate=decimate, exact=exact))
    mixer, fir = gr4.Rotator(phase_increment=dphi), gr4.fir_filter(b=taps, decimate=decimate, exact=exact)
    cuts = [0, decimate, 3 * decimate, 11 * decimate, 11 * decimate + 16 * decimate * 400, n - decimate, n]  # chunks shorter than the 126-sample history included
    for a, b in zip(cuts[:-1], cuts[1:]):
        got = fused.process_bulk(x[a:b]).cpu().numpy()
        want = fir.process_bulk(mixer.process_bulk(x[a:b])).cpu().numpy()
        assert_bit_equal(got, want, f"fused DDC D={decimate} chunk [{a},{b})")
    assert fused.mixer.accumulated_phase == mixer.accumulated_phase


@pytest.mark.parametrize("dphi,phi0", [(2 * np.pi * 0.0731, 0.0), (-2 * np.pi * 0.21, 0.5), (1e-3, 6.0), (0.7, 50.0)])
def test_ddc_fused_carried_halo_over_long_tile_ranges(gr4, dphi, phi0):
    """Calls long enough that every CTA of the fused kernel owns several consecutive tiles: from the second tile on the
    halo is the mixed tail of the previous tile moved inside shared memory (not fetched and rotated again), the wrap test
    is the one-sided one for the sign of dphi. Same bits as Rotator -> fir_filter, ragged end and a second call included."""
    rng = np.random.default_rng(91)
    n = 5120 * 148 * 5 * 2 * 3 + 8 * 777  # three tiles per CTA of the two-wave grid, then a partial tile
    x = dev(crandn(rng, n))
    dphi = float(np.float32(dphi))
    taps = gr4.fir_generate(127, "Hamming", 0.05)
    fused = gr4.DDC(gr4.Rotator(phase_increment=dphi, initial_phase=phi0), gr4.fir_filter(b=taps, decimate=8))
    mixer, fir = gr4.Rotator(phase_increment=dphi, initial_phase=phi0), gr4.fir_filter(b=taps, decimate=8)
    for a, b in ((0, n - 8 * 4000), (n - 8 * 4000, n)):
        got = fused.process_bulk(x[a:b])
        want = fir.process_bulk(mixer.process_bulk(x[a:b]))
        assert torch.equal(got.view(torch.int32), want.view(torch.int32)), f"fused DDC dphi={dphi} chunk [{a},{b})"
    assert fused.mixer.accumulated_phase == mixer.accumulated_phase


def test_ddc_fused_special_values_and_far_start_phase(gr4):
    """The fused kernel's straight-line mixer must hand non-finite products (Annex G recovery) and phases outside its
    fast sin/cos range to the checked path: same bits as the separate blocks."""
    rng = np.random.default_rng(77)
    n = 8 * 16 * 3000
    x = crandn(rng, n)
    x[[5000, 5001, 123456, 200001]] = [complex(np.inf, 1.0), complex(np.nan, np.inf), complex(-np.inf, np.inf), complex(0.0, np.nan)]
    xd = dev(x)
    taps = gr4.fir_generate(127, "Hamming", 0.05)
    for dphi, phi0 in ((0.3, 0.0), (-2.9, 1000.0), (3.1, -777.0)):
        fused = gr4.DDC(gr4.Rotator(phase_increment=dphi, initial_phase=phi0), gr4.fir_filter(b=taps, decimate=8))
        mixer, fir = gr4.Rotator(phase_increment=dphi, initial_phase=phi0), gr4.fir_filter(b=taps, decimate=8)
        got = fused.process_bulk(xd).cpu().numpy()
        want = fir.process_bulk(mixer.process_bulk(xd)).cpu().numpy()
        assert_bit_equal(got, want, f"fused DDC specials dphi={dphi} phi0={phi0}")


def test_polyphase_channelizer_against_own_oracle(gr4, oracle):
    """Config #5 building block. PARITY UNPINNED: the reference has no channelizer; the oracle is our own definition."""
    rng = np.random.default_rng(6)
    m, p, frames = 256, 12, 300
    proto = gr4.fir_generate(m * p, "Kaiser", 1.0 / (2 * m), beta=8.0)
    x = crandn(rng, m * frames)
    chan = gr4.PolyphaseChannelizer(proto, m)
    xd = dev(x)
    got = torch.cat([chan.process_bulk(xd[: m * 100].clone()), chan.process_bulk(xd[m * 100 :].clone())]).cpu().numpy()
    state = np.zeros((p - 1) * m, dtype=np.complex64)
    want = np.concatenate([oracle.pfb_channelizer(proto, m, x[: m * 100], state), oracle.pfb_channelizer(proto, m, x[m * 100 :], state)])
    u_norm = np.abs(want).max() * np.sqrt(m)
    assert np.abs(got - want).max() <= 4 * FFT_TOL * u_norm + 1e-7
    # forward-DFT convention of our definition: channel k is centred at -k/M, so a tone at +37/M lands in channel M-37
    tone = np.exp(2j * np.pi * (37 / m) * np.arange(m * 200)).astype(np.complex64)
    chan2 = gr4.PolyphaseChannelizer(proto, m)
    Y = chan2.process_bulk(dev(tone)).cpu().numpy()
    power = np.abs(Y[100:]).mean(axis=0)
    assert np.argmax(power) == m - 37 and power[m - 37] > 0.9 and np.delete(power, m - 37).max() < 1e-3


@pytest.mark.parametrize("m,p,frames", [(256, 12, 5000), (256, 4, 700), (64, 8, 3000), (1000, 12, 300), (16, 16, 9000), (48, 5, 500), (256, 24, 200)])
def test_polyphase_filter_stage_bit_exact_and_streaming(gr4, oracle, m, p, frames):
    """Stage 1 alone (register-ring kernel for P in {4, 8, 12, 16}, generic kernel otherwise) against the oracle's own
    definition, bit for bit, in three chunks so that the carried history and the stretch seams are exercised."""
    rng = np.random.default_rng(m * 31 + p)
    proto = rng.uniform(-1, 1, m * p).astype(np.float32)
    x = crandn(rng, m * frames)
    chan = gr4.PolyphaseChannelizer(proto, m)
    cuts = [0, m * (frames // 7), m * (frames // 2), m * frames]
    got = torch.cat([chan.filter_stage(dev(x[a:b])) for a, b in zip(cuts[:-1], cuts[1:])]).cpu().numpy()
    want = oracle.pfb_filter(proto, m, x)
    assert_bit_equal(got, want, f"pfb filter stage M={m} P={p}")


@pytest.mark.parametrize("p", [4, 8, 12])
def test_fused_channelizer_equals_the_two_stages(gr4, oracle, p):
    """256 channels: filter bank + FFT in one kernel (the bank outputs never reach HBM) gives the very bits of the two
    kernels back to back, across ragged chunks (frame counts that are not multiples of the 16-frame transform batch)."""
    rng = np.random.default_rng(40 + p)
    m, frames = 256, 3000 + p
    proto = gr4.fir_generate(m * p, "Kaiser", 1.0 / (2 * m), beta=8.0)
    x = crandn(rng, m * frames)
    fused, staged = gr4.PolyphaseChannelizer(proto, m), gr4.PolyphaseChannelizer(proto, m)
    assert fused.fused
    cuts = [0, m * 7, m * 1501, m * frames]
    got = torch.cat([fused.process_bulk(dev(x[a:b]), fused=True) for a, b in zip(cuts[:-1], cuts[1:])]).cpu().numpy()
    want = torch.cat([staged.process_bulk(dev(x[a:b]), fused=False) for a, b in zip(cuts[:-1], cuts[1:])]).cpu().numpy()
    assert_bit_equal(got, want, f"fused channelizer P={p}")
    ref = oracle.pfb_channelizer(proto, m, x[: m * 64])
    assert np.abs(got[:64] - ref).max() <= 4 * FFT_TOL * np.abs(ref).max() * np.sqrt(m) + 1e-7
    assert not gr4.PolyphaseChannelizer(gr4.fir_generate(64 * 8, "Kaiser", 1 / 128, beta=8.0), 64).fused


@pytest.mark.parametrize("interp,decim,n_taps", [(1, 1, 31), (3, 2, 72), (2, 3, 49), (160, 147, 160 * 12), (1, 8, 127), (7, 1, 70), (5, 4, 3), (4099, 4096, 4099 * 4)])
def test_polyphase_resampler_bit_exact_and_streaming(gr4, oracle, interp, decim, n_taps):
    """Rational resampler (own definition, PARITY UNPINNED: no reference block): bit for bit against our oracle, in three
    chunks (history carry-over, tile seams), and a tone keeps its frequency scaled by M/L."""
    rng = np.random.default_rng(interp * 1000 + decim)
    taps = (gr4.fir_generate(n_taps, "Kaiser", 0.45 / max(interp, decim), beta=6.0) * interp).astype(np.float32) if n_taps > 8 else rng.uniform(-1, 1, n_taps).astype(np.float32)
    n = decim * 20000
    x = crandn(rng, n)
    rs = gr4.PolyphaseResampler(taps, interp, decim)
    cuts = [0, decim * 3, decim * 9001, n]
    got = torch.cat([rs.process_bulk(dev(x[a:b])) for a, b in zip(cuts[:-1], cuts[1:])]).cpu().numpy()
    want = oracle.resampler(taps, interp, decim, x)
    assert_bit_equal(got, want, f"resampler L={interp} M={decim}")
    if n_taps > 8 and interp != decim:
        f_in = 0.02
        tone = np.exp(2j * np.pi * f_in * np.arange(n)).astype(np.complex64)
        y = gr4.PolyphaseResampler(taps, interp, decim).process_bulk(dev(tone)).cpu().numpy()[2000:]
        spec = np.abs(np.fft.fft(y[: 1 << 14] * np.hanning(1 << 14)))
        peak = np.fft.fftfreq(1 << 14)[np.argmax(spec)]
        assert abs(peak - f_in * decim / interp) < 2e-4
    if decim > 1:
        with pytest.raises(gr4.Gr4b200Error):
            rs.process_bulk(dev(x[: decim + 1]))


def test_ring_two_readers(gr4):
    """SPMC: space becomes writable only when the slowest reader has consumed it; each reader has its own cursor."""
    lib = gr4.load()
    ring = lib.gr4b200_ring_create(0, 1 << 20, 0)
    second = lib.gr4b200_ring_add_reader(ring)
    assert second == 1
    p = lib.gr4b200_ring_reserve(ring, 1 << 19, None)
    assert p and lib.gr4b200_ring_publish(ring, 1 << 19, None) == 0
    assert lib.gr4b200_ring_available_for(ring, 0) == 1 << 19 and lib.gr4b200_ring_available_for(ring, 1) == 1 << 19
    assert lib.gr4b200_ring_add_reader(ring) < 0  # readers join before the first publish
    q0 = lib.gr4b200_ring_get_for(ring, 0, 1 << 19, None)
    assert q0 == p and lib.gr4b200_ring_consume_for(ring, 0, 1 << 19, None) == 0
    assert lib.gr4b200_ring_writable(ring) == 1 << 19  # reader 1 still holds the first half
    assert lib.gr4b200_ring_reserve(ring, 1 << 19, None) and lib.gr4b200_ring_publish(ring, 1 << 19, None) == 0
    assert lib.gr4b200_ring_writable(ring) == 0
    q1 = lib.gr4b200_ring_get_for(ring, 1, 1 << 18, None)
    assert q1 == p and lib.gr4b200_ring_consume_for(ring, 1, 1 << 18, None) == 0
    assert lib.gr4b200_ring_writable(ring) == 1 << 18
    assert lib.gr4b200_ring_available_for(ring, 1) == (1 << 20) - (1 << 18) and lib.gr4b200_ring_available_for(ring, 0) == 1 << 19
    assert not lib.gr4b200_ring_get_for(ring, 2, 16, None)
    assert lib.gr4b200_ring_destroy(ring) == 0


def test_ring_cursor_protocol(gr4):
    import ctypes as C

    lib = gr4.load()
    ring = lib.gr4b200_ring_create(0, 1 << 20, 0)
    assert ring and lib.gr4b200_ring_capacity(ring) == 1 << 20
    assert lib.gr4b200_ring_available(ring) == 0 and lib.gr4b200_ring_writable(ring) == 1 << 20
    p = lib.gr4b200_ring_reserve(ring, 3 << 18, None)
    assert p and lib.gr4b200_ring_publish(ring, 3 << 18, None) == 0
    assert lib.gr4b200_ring_available(ring) == 3 << 18
    assert not lib.gr4b200_ring_reserve(ring, 1 << 19, None)  # only 1<<18 left
    q = lib.gr4b200_ring_get(ring, 1 << 18, None)
    assert q == p and lib.gr4b200_ring_consume(ring, 1 << 18, None) == 0
    assert lib.gr4b200_ring_writable(ring) == 1 << 18  # contiguous part up to the end of the ring
    assert lib.gr4b200_ring_destroy(ring) == 0


def test_ring_history_stays_in_front_of_every_span(gr4):
    """A ring with history: the bytes behind the reader's cursor are not handed back to the writer, and when the reader
    leaves the end of the ring the tail is copied in front of the base -- every span finds the stream's past at p[-h..-1]."""
    import ctypes as C

    lib = gr4.load()
    cap, h, piece = 4096, 1024, 1024
    ring = lib.gr4b200_ring_create(0, cap, h)
    assert lib.gr4b200_ring_writable(ring) == cap - h  # the (still empty) history counts as in use
    assert lib.gr4b200_ring_add_reader(ring) < 0  # the reader maintains the history: exactly one
    stream_bytes = np.arange(16 * piece, dtype=np.uint32).view(np.uint8)  # 64 KiB of distinct words
    position, seen = 0, 0
    back = np.zeros(h + piece, dtype=np.uint8)
    for step in range(40):
        if lib.gr4b200_ring_writable(ring) >= piece and position + piece <= stream_bytes.size:
            dst = lib.gr4b200_ring_reserve(ring, piece, None)
            assert dst
            chunk = np.ascontiguousarray(stream_bytes[position : position + piece])
            assert lib.gr4b200_copy_h2d(C.c_void_p(dst), chunk.ctypes.data_as(C.c_void_p), piece, None) == 0
            lib.gr4b200_stream_synchronize(None)
            assert lib.gr4b200_ring_publish(ring, piece, None) == 0
            position += piece
        if lib.gr4b200_ring_available(ring) >= piece:
            src = lib.gr4b200_ring_get(ring, piece, None)
            assert src
            assert lib.gr4b200_copy_d2h(back.ctypes.data_as(C.c_void_p), C.c_void_p(src - h), h + piece, None) == 0
            lib.gr4b200_stream_synchronize(None)
            want = np.zeros(h + piece, dtype=np.uint8)
            lo = max(0, seen - h)
            want[h - (seen - lo) :] = stream_bytes[lo : seen + piece]
            assert np.array_equal(back, want), f"span at stream offset {seen}: history or data wrong"
            assert lib.gr4b200_ring_consume(ring, piece, None) == 0
            seen += piece
    assert seen >= 12 * piece  # several turns of the ring
    assert lib.gr4b200_ring_destroy(ring) == 0


# ---- sample-format converters (reference tests: blocks/basic/test/qa_Converter.cpp:242-268) ------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.int16, torch.int8])
@pytest.mark.parametrize("n", [0, 1, 2, 3, 1000, 65537, (1 << 22) + 5])
def test_interleaved_to_complex_is_exact(gr4, oracle, dtype, n):
    rng = np.random.default_rng(n + 1)
    np_dtype = {torch.float32: np.float32, torch.int16: np.int16, torch.int8: np.int8}[dtype]
    if dtype == torch.float32:
        items = rng.uniform(-1, 1, 2 * n).astype(np.float32)
    else:
        info = np.iinfo(np_dtype)
        items = rng.integers(info.min, info.max, 2 * n, dtype=np_dtype, endpoint=True)
    got = gr4.InterleavedToComplex(dtype).process_bulk(torch.from_numpy(items).cuda()).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), oracle.interleaved_to_complex(items).view(np.uint32))


@pytest.mark.parametrize("dtype", [torch.float32, torch.int16, torch.int8])
def test_complex_to_interleaved_matches_the_cast(gr4, oracle, dtype):
    """static_cast<R>(float): truncation toward zero; halves, negative values, both ends of the range, values far outside
    it, infinities and NaN -- all as the oracle's compiled cast produces them, at every position of a long buffer."""
    np_dtype = {torch.float32: np.float32, torch.int16: np.int16, torch.int8: np.int8}[dtype]
    rng = np.random.default_rng(9)
    n = (1 << 20) + 3
    scale = 40000.0 if dtype == torch.int16 else 200.0
    x = (rng.uniform(-scale, scale, n) + 1j * rng.uniform(-scale, scale, n)).astype(np.complex64)
    special = np.array([0.5, -0.5, 1.5, -1.5, 127.99, -128.99, 32767.5, -32768.5, 65536.0, -65537.0, 2147483520.0, -2147483648.0, 3e9, -3e9, 1e30, np.inf, -np.inf, np.nan, -0.0, 255.9], dtype=np.float32)
    x[100 : 100 + special.size].real, x[100 : 100 + special.size].imag = special, special[::-1]
    x[-special.size :].real, x[-special.size :].imag = special[::-1], -special
    want = oracle.complex_to_interleaved(x, np_dtype)
    block = gr4.ComplexToInterleaved(dtype)
    got = block.process_bulk(dev(x)).cpu().numpy()
    if dtype == torch.float32:
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    else:
        assert np.array_equal(got, want)
    # a view that starts on an odd sample (8-byte aligned only): the item-wise kernel, same values
    shifted = block.process_bulk(dev(x)[1:]).cpu().numpy()
    assert np.array_equal(shifted.view(np.uint8), want[2:].view(np.uint8))
    # round trip through the integer format and back is the identity on integer-valued samples
    if dtype != torch.float32:
        info = np.iinfo(np_dtype)
        items = rng.integers(info.min, info.max, 2 * 4097, dtype=np_dtype, endpoint=True)
        there = gr4.InterleavedToComplex(dtype).process_bulk(torch.from_numpy(items).cuda())
        assert np.array_equal(block.process_bulk(there).cpu().numpy(), items)


def test_fir_history_from_the_stream_equals_the_carried_state(gr4, oracle):
    """gr4b200_fir_cf32_contiguous reads its past samples in front of the span it is given (an HBM ring with history keeps
    them there) instead of the plan's state: same bits as the stateful call, chunk after chunk, full rate and /8."""
    import ctypes as C

    lib = gr4.load()
    rng = np.random.default_rng(91)
    taps = gr4.fir_generate(127, "Hamming", 0.1)
    for decimate in (1, 8):
        plan = lib.gr4b200_fir_plan_create(taps.ctypes.data_as(C.c_void_p), taps.size, decimate, 1)
        h = lib.gr4b200_fir_plan_history_items(plan)
        assert h == 128
        n = 8 * 5000
        x = crandn(rng, n)
        padded = dev(np.concatenate([np.zeros(h, dtype=np.complex64), x]))  # the stream with h zeros of "before the start"
        out = torch.empty(n // decimate, dtype=torch.complex64, device="cuda")
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        cuts = [0, 8 * 7, 8 * 300, 8 * 2049, n]
        for a, b in zip(cuts[:-1], cuts[1:]):  # every chunk finds its history in front of it: no state is carried
            rc = lib.gr4b200_fir_cf32_contiguous(plan, stream, C.c_void_p(padded.data_ptr() + 8 * (h + a)), C.c_void_p(out.data_ptr() + 8 * (a // decimate)), b - a)
            assert rc == 0
        torch.cuda.synchronize()
        lib.gr4b200_fir_plan_destroy(plan)
        assert_bit_equal(out.cpu().numpy(), oracle.fir(taps, x, decimate=decimate), f"contiguous FIR, decimate {decimate}")


@pytest.mark.parametrize("n_taps", [2, 33, 127, 128, 1000, 2049])
def test_fir_overlap_save_mode_within_its_tolerance(gr4, oracle, n_taps):
    """The opt-in tolerance mode (y = IFFT(FFT(x) . FFT(b)) on blocks of 4096): against the exact mode's result, ragged
    chunks with the history carried across them. Stated bound: 2e-6 * sum|b| * max|x| (float transforms of 4096 points);
    the measured maximum is printed."""
    rng = np.random.default_rng(n_taps)
    n = 3 * 4096 * 5 + 777
    x = crandn(rng, n)
    taps = gr4.fir_generate(n_taps, "Hamming", 0.1) if n_taps > 2 else np.array([0.75, -0.25], dtype=np.float32)
    want = oracle.fir(taps, x)
    f = gr4.fir_filter(b=taps, overlap_save=True)
    xd = dev(x)
    cuts = [0, 1, 130, 5000, 5000 + 4096 * 3, n]
    got = np.concatenate([f.process_bulk(xd[a:b].clone()).cpu().numpy() for a, b in zip(cuts[:-1], cuts[1:])])
    bound = 2e-6 * np.abs(taps).sum() * np.abs(x).max()
    err = np.abs(got - want).max()
    print(f"overlap-save, {n_taps} taps: max error {err:.3g} (bound {bound:.3g}, {err / (np.abs(taps).sum() * np.abs(x).max()):.3g} of sum|b| max|x|)")
    assert err <= bound


def test_fir_overlap_save_mode_limits(gr4):
    taps = gr4.fir_generate(127, "Hamming", 0.1)
    with pytest.raises(gr4.Gr4b200Error):
        gr4.fir_filter(b=taps, decimate=8, overlap_save=True)
    with pytest.raises(gr4.Gr4b200Error):
        gr4.fir_filter(b=gr4.fir_generate(2050, "Hamming", 0.1), overlap_save=True)
    with pytest.raises(gr4.Gr4b200Error):
        gr4.fir_filter(b=taps, overlap_save=True).process_bulk(torch.zeros(4096, dtype=torch.float32, device="cuda"))


def test_grc_flowgraph_runs_like_the_hand_built_one(gr4, oracle):
    """A .grc document (Graph_yaml_importer.hpp) of the FIR -> FFT flowgraph, loaded and run through Graph / Simple."""
    taps = gr4.fir_generate(127, "Hamming", 0.1)
    text = "blocks:\n  - id: gr::filter::fir_filter<complex64>\n    parameters:\n      name: lowpass\n      b: [" + ", ".join(repr(float(t)) for t in taps) + "]\n"
    text += "  - id: gr::blocks::fft::FFT<complex64>\n    parameters:\n      name: spectrum\n      fftSize: !!uint32 4096\n      window: Hann\nconnections:\n  - [lowpass, 0, spectrum, 0]\n"
    graph = gr4.load_grc(text)
    assert [type(b).__name__ for b in graph.blocks] == ["fir_filter", "FFT"] and graph.blocks[0].name == "lowpass"
    n = 4096 * 24
    x = crandn(np.random.default_rng(12), n)
    src, dst = gr4.HostBuffer(n, np.complex64), gr4.HostBuffer(n * 4, np.float32)
    src.array[:] = x
    sched = gr4.Simple(graph, chunk_items=4096 * 5)
    sched.runAndWait(src.array, dst.array)
    got = dst.array.reshape(n // 4096, 4, 4096).copy()
    y = gr4.fir_filter(b=taps).process_bulk(dev(x))
    want = gr4.FFT(fftSize=4096, window="Hann").process_bulk(y).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    again = gr4.parse_grc(gr4.save_grc(graph))
    assert [b[0] for b in again[0]] == ["gr::filter::fir_filter", "gr::blocks::fft::FFT"] and again[1] == [("lowpass", 0, "spectrum", 0, None)]
    sched.close(), src.close(), dst.close()
