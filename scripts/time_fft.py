"""CUDA-event timing of the FFT family (GPU box): every size, spectrum and block mode, and the A/B variants
(GR4B200_FFT_TMA=0: direct loads instead of bulk staging)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnuradio4_b200 as gr4

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
peak = 6547.5
torch.manual_seed(1234)
x = torch.empty(n, dtype=torch.complex64, device="cuda")
torch.view_as_real(x).uniform_(-1, 1)
y = torch.empty_like(x)
sig = torch.empty(4 * n, dtype=torch.float32, device="cuda")


def timeit(name, fn, bytes_per_sample, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    gbs = bytes_per_sample * n / ms / 1e6
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "GS/s": round(n / ms / 1e6, 2), "GB/s": round(gbs, 1), "frac_hbm": round(gbs / peak, 4)}), flush=True)


only_large = os.environ.get("TIME_FFT_ONLY_LARGE") == "1"
for size in (16384, 32768, 65536, 131072, 262144):  # two passes of column transforms: 32 algorithmic bytes per sample for the spectrum
    f = gr4.FFT(fftSize=size, window="Hann")
    timeit(f"fft{size} c2c [two column passes]", lambda: f.compute(x, out=y), 32)
    timeit(f"fft{size} c2c windowed [two column passes]", lambda: f.compute(x, out=y, windowed=True), 32)
    timeit(f"fft{size} block [two column passes]", lambda: f.process_bulk(x, signals=sig.view(n // size, 4, size)), 40)
if only_large:
    sys.exit(0)
variants = [("radix", {}), ("radix, direct loads", {"GR4B200_FFT_TMA": "0"})]
for size in (4096, 1024, 2048, 8192, 256, 512, 128, 64, 32, 16):
    for label, env in variants:
        if label == "radix, direct loads" and size < 1024:
            continue
        for k in ("GR4B200_FFT_TMA",):
            os.environ.pop(k, None)
        os.environ.update(env)
        f = gr4.FFT(fftSize=size, window="Hann")
        s = sig.view(n // size, 4, size)
        timeit(f"fft{size} c2c [{label}]", lambda: f.compute(x, out=y), 16)
        if size in (4096, 1024, 256):
            timeit(f"fft{size} c2c windowed [{label}]", lambda: f.compute(x, out=y, windowed=True), 16)
        timeit(f"fft{size} block [{label}]", lambda: f.process_bulk(x, signals=s), 24)
for k in ("GR4B200_FFT_TMA",):
    os.environ.pop(k, None)
xr = torch.view_as_real(x).reshape(-1)[:n].contiguous()  # n real samples
for size in (4096, 1024, 256):
    f = gr4.FFT(fftSize=size, window="Hann")
    timeit(f"fft{size} real input -> full spectrum", lambda: f.compute_real(xr, out=y), 12)
# sizes that are not a power of two: chirp-z over the power-of-two kernels (fft_bluestein.cuh); n_used = whole transforms
for size in (1000, 1009, 96, 3000):
    f = gr4.FFT(fftSize=size, window="Hann")
    m = (n // 8) // size * size  # an eighth of the buffer: the padded transforms need scratch of their own
    xs, ys = x[:m], y[:m]
    for _ in range(2):
        f.compute(xs, out=ys)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        f.compute(xs, out=ys)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(json.dumps({"kernel": f"fft{size} c2c [chirp-z]", "ms": round(ms, 4), "GS/s": round(m / ms / 1e6, 2), "samples": m}), flush=True)
t = torch.empty_like(x)
timeit("copy (torch)", lambda: t.copy_(x), 16)
