"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck) on the GPU box."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gnuradio4_b200 as gr4

torch.manual_seed(0)
n = 4096 * 24
x = torch.empty(n, dtype=torch.complex64, device="cuda")
torch.view_as_real(x).uniform_(-1, 1)
taps = gr4.fir_generate(127, "Hamming", 0.1)
for size in (16, 64, 256, 512, 1024, 2048, 4096, 8192):
    f = gr4.FFT(fftSize=size, window="Hann")
    f.compute(x[: size * 11], windowed=True)
    f.process_bulk(x[: size * 11], want_ranges=True)
    os.environ["GR4B200_FFT_TMA"] = "0"
    g = gr4.FFT(fftSize=size, window="Hann")
    g.compute(x[: size * 11], windowed=True)
    os.environ.pop("GR4B200_FFT_TMA")
big = torch.empty(262144 * 2, dtype=torch.complex64, device="cuda")
torch.view_as_real(big).uniform_(-1, 1)
for size in (16384, 32768, 131072, 262144):  # two passes of column transforms (128-, 256- and 512-point columns)
    f = gr4.FFT(fftSize=size, window="Hann")
    f.compute(big[: size * 2], windowed=True)
    f.process_bulk(big[: size * 2], want_ranges=True)
xr = torch.view_as_real(x).reshape(-1)[: 4096 * 7].contiguous()
for size in (64, 1024, 4096):
    f = gr4.FFT(fftSize=size, window="Hann")
    f.compute_real(xr[: size * 7])
    f.process_bulk_real(xr[: size * 7], want_ranges=True)
for dtype in (torch.int16, torch.int8):
    items = gr4.ComplexToInterleaved(dtype).process_bulk(x[:100001] * 100.0)
    gr4.InterleavedToComplex(dtype).process_bulk(items)
    gr4.ComplexToInterleaved(dtype).process_bulk((x * 100.0)[1:100000])  # 8-byte aligned view: item-wise kernels
gr4.fir_filter(b=taps).process_bulk(x)
gr4.fir_filter(b=taps, exact=False).process_bulk(x)
for d in (2, 4, 8, 16, 5):
    gr4.fir_filter(b=taps, decimate=d).process_bulk(x[: n // 80 * 80])
gr4.fir_filter(b=taps, decimate=8, exact=False).process_bulk(x[: n // 80 * 80])
gr4.fir_filter(b=gr4.fir_generate(1001, "Hamming", 0.1), decimate=8).process_bulk(x[: n // 80 * 80])  # long filter: CTA-wide tiles, shared tap pairs
long_call = torch.empty(75 * 4096 + 100, dtype=torch.complex64, device="cuda")  # more than 74 tiles: the 256 x 16 tiles with parameter taps
torch.view_as_real(long_call).uniform_(-1, 1)
gr4.fir_filter(b=taps).process_bulk(long_call)
gr4.FirFft(gr4.fir_filter(b=taps), gr4.FFT(fftSize=4096, window="Hann")).process_bulk(x)
gr4.Rotator(phase_increment=0.6283185).process_bulk(x)
gr4.DDC(gr4.Rotator(phase_increment=0.6283185), gr4.fir_filter(b=taps, decimate=8)).process_bulk(x)
# one resident wave (GR4B200_FIR_GRID_MULT=1, set by gpu_sanitize.sh) and more tiles than CTAs: every CTA of the fused kernel
# owns several consecutive tiles, so the halo carried in shared memory and the one-sided phase step are exercised
carry = torch.empty(1280 * 148 * 20 * 2 + 8 * 100, dtype=torch.complex64, device="cuda")
torch.view_as_real(carry).uniform_(-1, 1)
for dphi in (0.6283185, -1.9):
    gr4.DDC(gr4.Rotator(phase_increment=dphi), gr4.fir_filter(b=taps, decimate=8)).process_bulk(carry)
del carry
gr4.MultiplyConst(value=2 + 1j).process_bulk(x)
gr4.Multiply(n_inputs=3).process_bulk([x, x, x])
proto = gr4.fir_generate(256 * 12, "Kaiser", 1 / 512, beta=8.0)
ch = gr4.PolyphaseChannelizer(proto, 256)
ch.process_bulk(x[: 256 * 333], fused=True)
ch.process_bulk(x[: 256 * 333], fused=False)
gr4.PolyphaseResampler(taps, 3, 2).process_bulk(x)
gr4.PolyphaseResampler(taps, 160, 147).process_bulk(x[: 147 * 600])
torch.cuda.synchronize()
print("sanitize run done")
