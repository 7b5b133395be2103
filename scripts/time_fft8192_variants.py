import os, sys, json
sys.path.insert(0, os.getcwd())
import torch
import gnuradio4_b200 as gr4
n = 1 << 28
x = torch.empty(n, dtype=torch.complex64, device="cuda"); torch.view_as_real(x).uniform_(-1, 1)
y = torch.empty_like(x); sig = torch.empty(4 * n, dtype=torch.float32, device="cuda")
def timeit(name, fn, bps, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "GS/s": round(n / ms / 1e6, 2), "frac_hbm": round(bps * n / ms / 1e6 / 6547.5, 4)}), flush=True)
size = 8192
f = gr4.FFT(fftSize=size, window="Hann")
tag = f"TMA={os.environ.get('GR4B200_FFT_TMA','1')} MULT={os.environ.get('GR4B200_FFT_GRID_MULT','default')}"
timeit(f"fft8192 c2c [{tag}]", lambda: f.compute(x, out=y), 16)
timeit(f"fft8192 block [{tag}]", lambda: f.process_bulk(x, signals=sig.view(n // size, 4, size)), 24)
