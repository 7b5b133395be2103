"""CUDA-event timing of every hot kernel at a given size (GPU box). Prints one JSON line per kernel."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnuradio4_b200 as gr4

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None  # optional comma-separated substrings of kernel names
peak = 6547.5
torch.manual_seed(1234)
x = torch.empty(n, dtype=torch.complex64, device="cuda")
torch.view_as_real(x).uniform_(-1, 1)
y = torch.empty_like(x)
taps = gr4.fir_generate(127, "Hamming", 0.1)


def timeit(name, fn, bytes_per_sample, reps=5):
    if only is not None and not any(o in name for o in only):
        return
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
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "GS/s": round(n / ms / 1e6, 2), "GB/s": round(gbs, 1), "frac_hbm": round(gbs / peak, 4), "tune": os.environ.get("GR4B200_FIR_TUNE", "0"), "fir_grid_mult": os.environ.get("GR4B200_FIR_GRID_MULT", "default")}))


def checksum(t):  # order-independent bit-level fingerprint (compares tuning variants of an exact kernel)
    return int(torch.view_as_real(t).view(torch.int32).to(torch.int64).sum().item())


f_exact, f_fast = gr4.fir_filter(b=taps), gr4.fir_filter(b=taps, exact=False)
timeit("fir127 exact", lambda: f_exact.process_bulk(x, out=y), 16)
if only is not None and "fir127 exact" in only:
    print(json.dumps({"checksum fir127 exact": checksum(y)}))
timeit("fir127 fast", lambda: f_fast.process_bulk(x, out=y), 16)
f_ols = gr4.fir_filter(b=taps, overlap_save=True)
timeit("fir127 overlap-save (tolerance mode)", lambda: f_ols.process_bulk(x, out=y), 16)
for d in (2, 4, 8, 16):
    fd = gr4.fir_filter(b=taps, decimate=d)
    yd = torch.empty(n // d, dtype=torch.complex64, device="cuda")
    timeit(f"fir127 decim{d} exact", lambda: fd.process_bulk(x, out=yd), 8 + 8 / d)
    if only is not None and f"fir127 decim{d} exact" in only:
        print(json.dumps({f"checksum fir127 decim{d} exact": checksum(yd)}))
    fdf = gr4.fir_filter(b=taps, decimate=d, exact=False)
    timeit(f"fir127 decim{d} fast", lambda: fdf.process_bulk(x, out=yd), 8 + 8 / d)
fft = gr4.FFT(fftSize=4096, window="Hann")
sig = torch.empty((n // 4096, 4, 4096), dtype=torch.float32, device="cuda")
timeit("fft4096 c2c", lambda: fft.compute(x, out=y), 16)
timeit("fft4096 c2c windowed", lambda: fft.compute(x, out=y, windowed=True), 16)
timeit("fft4096 block", lambda: fft.process_bulk(x, signals=sig), 24)
fused = gr4.FirFft(gr4.fir_filter(b=taps), gr4.FFT(fftSize=4096, window="Hann"))
timeit("fir127 exact -> fft4096 block, two kernels", lambda: fft.process_bulk(f_exact.process_bulk(x, out=y), signals=sig), 40)
timeit("fir127 exact -> fft4096 block, fused", lambda: fused.process_bulk(x, signals=sig), 24)
timeit("fir127 overlap-save -> fft4096 block, two kernels", lambda: fft.process_bulk(f_ols.process_bulk(x, out=y), signals=sig), 40)
fused_fast = gr4.FirFft(gr4.fir_filter(b=taps, exact=False), gr4.FFT(fftSize=4096, window="Hann"))
timeit("fir127 fast -> fft4096 block, fused", lambda: fused_fast.process_bulk(x, signals=sig), 24)
f256 = gr4.FFT(fftSize=256, window="Hann")
timeit("fft256 c2c", lambda: f256.compute(x, out=y), 16)
f1024 = gr4.FFT(fftSize=1024, window="Hann")
timeit("fft1024 c2c (generic)", lambda: f1024.compute(x, out=y), 16)
rot = gr4.Rotator(phase_increment=0.6283185)
timeit("rotator", lambda: rot.process_bulk(x, out=y), 16)
for cls in ("AddConst", "MultiplyConst", "DivideConst"):
    m = getattr(gr4, cls)(value=2 + 1j)
    timeit(cls, lambda: m.process_bulk(x, out=y), 16)
ddc = gr4.DDC(gr4.Rotator(phase_increment=0.6283185), gr4.fir_filter(b=taps, decimate=8))
yd = torch.empty(n // 8, dtype=torch.complex64, device="cuda")
timeit("ddc (mixer + fir/8), fused", lambda: ddc.process_bulk(x, out=yd), 9)
proto = gr4.fir_generate(256 * 12, "Kaiser", 1 / 512, beta=8.0)
ch = gr4.PolyphaseChannelizer(proto, 256)
timeit("pfb filter stage", lambda: ch.filter_stage(x, out=y), 16)
f256p = gr4.FFT(fftSize=256, window="Hann")
scratch = torch.empty_like(x)
timeit("pfb fft stage (fft256 c2c)", lambda: f256p.compute(y, out=scratch), 16)
timeit("pfb channelizer, two stages", lambda: ch.process_bulk(x, out=y, fused=False), 32)
timeit("pfb channelizer, fused", lambda: ch.process_bulk(x, out=y, fused=True), 16)
rtaps = (gr4.fir_generate(160 * 12, "Kaiser", 0.45 / 160, beta=6.0) * 160).astype("float32")
for interp, decim in ((160, 147), (3, 2), (2, 3), (1, 1)):
    rs = gr4.PolyphaseResampler(rtaps[: interp * 12] if interp < 160 else rtaps, interp, decim)
    n_in = (n // 2) // decim * decim
    xin = x[:n_in]
    rout = torch.empty(n_in // decim * interp, dtype=torch.complex64, device="cuda")
    name = f"resampler {interp}/{decim} (12 taps per phase)"
    if only is None or any(o in name for o in only):
        for _ in range(2):
            rs.process_bulk(xin, out=rout)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            rs.process_bulk(xin, out=rout)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        gbs = 8.0 * (n_in + rout.numel()) / ms / 1e6
        print(json.dumps({"kernel": name, "ms": round(ms, 4), "GS/s in": round(n_in / ms / 1e6, 2), "GS/s out": round(rout.numel() / ms / 1e6, 2), "GB/s": round(gbs, 1), "frac_hbm": round(gbs / peak, 4)}))
    del rout
t = torch.empty_like(x)
i16 = torch.randint(-32768, 32767, (2 * n,), dtype=torch.int16, device="cuda")
i8 = torch.randint(-128, 127, (2 * n,), dtype=torch.int8, device="cuda")
widen16, widen8 = gr4.InterleavedToComplex(torch.int16), gr4.InterleavedToComplex(torch.int8)
narrow16, narrow8 = gr4.ComplexToInterleaved(torch.int16), gr4.ComplexToInterleaved(torch.int8)
timeit("int16 I/Q -> complex<float>", lambda: widen16.process_bulk(i16, out=y), 12)
timeit("int8 I/Q -> complex<float>", lambda: widen8.process_bulk(i8, out=y), 10)
timeit("complex<float> -> int16 I/Q", lambda: narrow16.process_bulk(x, out=i16), 12)
timeit("complex<float> -> int8 I/Q", lambda: narrow8.process_bulk(x, out=i8), 10)
timeit("copy (torch)", lambda: t.copy_(x), 16)
