"""Generates tests/golden/*.npz from the REFERENCE's own code (oracle/_ref/libgr4ref.so = /root/reference sources
compiled in place, see oracle/ref_harness.cpp). Run in the build container only:  python tests/golden/make_golden.py
The fixtures pin the oracle (tests/test_oracle.py) and the GPU path (tests/test_gpu_golden.py) where /root/reference and
oracle/_ref do not exist. Inputs are regenerated from the seeds stored in the files."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import _oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def crand(seed, n):
    rng = np.random.default_rng(seed)
    return (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)


def main():
    ref = _oracle.load_ref()
    assert ref is not None, "oracle/_ref/libgr4ref.so missing: run `make -C oracle ref` where /root/reference exists"
    # windows: all 12 types at N = 8 (the reference's own table size), 127 and 4096
    windows = {f"{name}_{n}": ref.window(name, n) for name in _oracle.WINDOWS for n in (8, 127, 4096)}
    np.savez_compressed(os.path.join(HERE, "windows.npz"), **windows)
    # FIR design
    design = {"lowpass127_hamming_fc0p1": ref.fir_generate(127, "Hamming", 0.1), "lowpass127_hamming_fc0p05": ref.fir_generate(127, "Hamming", 0.05)}
    for ftype in ("LOWPASS", "HIGHPASS", "BANDPASS", "BANDSTOP"):
        for win in ("Kaiser", "Hamming", "Hann"):
            design[f"design_{ftype}_{win}"] = ref.fir_design(ftype, 4, 1.0, 10.0, 1000.0, window=win)
    np.savez_compressed(os.path.join(HERE, "fir_design.npz"), **design)
    # FIR outputs (seeded inputs): 127-tap complex, decimate 8, real 10-tap step
    taps = design["lowpass127_hamming_fc0p1"]
    x = crand(1234, 16384)
    fir = {"seed": 1234, "n": 16384, "y127": ref.fir(taps, x), "y127_decim8": ref.fir(design["lowpass127_hamming_fc0p05"], x, decimate=8), "step10": ref.fir(np.full(10, 0.1, dtype=np.float32), np.ones(32, dtype=np.float32))}
    for nt in (5, 33, 48, 200):
        t = np.random.default_rng(nt).uniform(-1, 1, nt).astype(np.float32)
        fir[f"taps_{nt}"] = t
        fir[f"y_{nt}"] = ref.fir(t, x[:4096])
    np.savez_compressed(os.path.join(HERE, "fir.npz"), **fir)
    # FFT: 4 transforms of 4096, 8 of 256, one each of the other sizes
    fft = {"seed": 4321}
    for n, batch in ((16, 4), (64, 2), (256, 8), (1024, 2), (4096, 4), (8192, 1)):
        fft[f"X_{n}"] = ref.fft(crand(4321 + n, n * batch), n)
        fft[f"batch_{n}"] = batch
    xb = (crand(99, 4096 * 2) * 0.1 + np.exp(2j * np.pi * 0.1 * np.arange(8192))).astype(np.complex64)
    sig, ranges = ref.fft_block(xb, 4096, ref.window("Hann", 4096))
    fft["block_signals"], fft["block_ranges"] = sig, ranges
    sig_db, _ = ref.fft_block(xb, 4096, ref.window("Hann", 4096), db=True, deg=True)
    fft["block_signals_db_deg"] = sig_db
    np.savez_compressed(os.path.join(HERE, "fft.npz"), **fft)
    # math / mixer
    xm = crand(77, 4096) * np.exp(np.random.default_rng(78).uniform(-20, 20, 4096)).astype(np.float32)
    math = {"seed": 77}
    for op in ("add", "subtract", "multiply", "divide"):
        math[op] = ref.mathop_const(op, xm, 0.37 - 1.91j)
    rot, phase = ref.rotator(crand(79, 65536), float(np.float32(2 * np.pi * 0.1)), 0.0)
    math["rotator"], math["rotator_end_phase"] = rot[-4096:], np.float32(phase)
    np.savez_compressed(os.path.join(HERE, "math.npz"), **math)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
