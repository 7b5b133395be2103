"""Short driver for ncu captures (GPU box): runs each hot kernel a few times on 2^26 samples."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnuradio4_b200 as gr4

n = 1 << 26
which = sys.argv[1] if len(sys.argv) > 1 else "all"
x = torch.empty(n, dtype=torch.complex64, device="cuda")
torch.view_as_real(x).uniform_(-1, 1)
y = torch.empty_like(x)
taps = gr4.fir_generate(127, "Hamming", 0.1)
reps = 3
if which in ("all", "fir"):
    f = gr4.fir_filter(b=taps)
    for _ in range(reps):
        f.process_bulk(x, out=y)
if which in ("all", "firfast"):
    f = gr4.fir_filter(b=taps, exact=False)
    for _ in range(reps):
        f.process_bulk(x, out=y)
if which in ("all", "firdecim"):
    f = gr4.fir_filter(b=taps, decimate=8)
    yd = torch.empty(n // 8, dtype=torch.complex64, device="cuda")
    for _ in range(reps):
        f.process_bulk(x, out=yd)
if which in ("all", "ddc"):
    d = gr4.DDC(gr4.Rotator(phase_increment=0.6283185), gr4.fir_filter(b=taps, decimate=8))
    yd = torch.empty(n // 8, dtype=torch.complex64, device="cuda")
    for _ in range(reps):
        d.process_bulk(x, out=yd)
if which in ("all", "fft"):
    f = gr4.FFT(fftSize=4096, window="Hann")
    sig = torch.empty((n // 4096, 4, 4096), dtype=torch.float32, device="cuda")
    for _ in range(reps):
        f.compute(x, out=y)
    for _ in range(reps):
        f.process_bulk(x, signals=sig)
if which == "fftc2c":
    f = gr4.FFT(fftSize=4096, window="Hann")
    for _ in range(reps):
        f.compute(x, out=y, windowed=True)
if which == "fftblock":
    f = gr4.FFT(fftSize=4096, window="Hann")
    sig = torch.empty((n // 4096, 4, 4096), dtype=torch.float32, device="cuda")
    for _ in range(reps):
        f.process_bulk(x, signals=sig)
if which == "fftlarge":
    f = gr4.FFT(fftSize=65536, window="Hann")
    for _ in range(reps):
        f.compute(x, out=y)
if which == "pfb":
    proto = gr4.fir_generate(256 * 12, "Kaiser", 1 / 512, beta=8.0)
    ch = gr4.PolyphaseChannelizer(proto, 256)
    for _ in range(reps):
        ch.filter_stage(x, out=y)
if which == "channelizer":
    proto = gr4.fir_generate(256 * 12, "Kaiser", 1 / 512, beta=8.0)
    ch = gr4.PolyphaseChannelizer(proto, 256)
    for _ in range(reps):
        ch.process_bulk(x, out=y, fused=True)
if which == "firfft":
    f = gr4.FirFft(gr4.fir_filter(b=taps), gr4.FFT(fftSize=4096, window="Hann"))
    sig = torch.empty((n // 4096, 4, 4096), dtype=torch.float32, device="cuda")
    for _ in range(reps):
        f.process_bulk(x, signals=sig)
if which == "resampler":
    rtaps = (gr4.fir_generate(160 * 12, "Kaiser", 0.45 / 160, beta=6.0) * 160).astype("float32")
    rs = gr4.PolyphaseResampler(rtaps, 160, 147)
    xin = x[: n // 147 * 147]
    for _ in range(reps):
        rs.process_bulk(xin)
if which in ("all", "rot"):
    r = gr4.Rotator(phase_increment=0.6283185)
    for _ in range(reps):
        r.process_bulk(x, out=y)
if which in ("all", "math"):
    m = gr4.MultiplyConst(value=2 + 1j)
    for _ in range(reps):
        m.process_bulk(x, out=y)
torch.cuda.synchronize()
print("done", which)
