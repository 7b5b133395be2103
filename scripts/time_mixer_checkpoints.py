"""Fused DDC and mixer at several phase increments (GPU box): what the checkpoint pass (GR4B200_CHECKPOINT_STRETCH) costs."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnuradio4_b200 as gr4

n = 1 << 28
x = torch.empty(n, dtype=torch.complex64, device="cuda")
torch.view_as_real(x).uniform_(-1, 1)
y = torch.empty_like(x)
yd = torch.empty(n // 8, dtype=torch.complex64, device="cuda")
taps = gr4.fir_generate(127, "Hamming", 0.1)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for dphi in (0.6283185, -1.9, 0.01, 1e-3):
    ddc = gr4.DDC(gr4.Rotator(phase_increment=dphi), gr4.fir_filter(b=taps, decimate=8))
    rot = gr4.Rotator(phase_increment=dphi)
    ms_d, ms_r = timeit(lambda: ddc.process_bulk(x, out=yd)), timeit(lambda: rot.process_bulk(x, out=y))
    print(json.dumps({"stretch": os.environ.get("GR4B200_CHECKPOINT_STRETCH", "default"), "dphi": dphi, "ddc_ms": round(ms_d, 4), "ddc_GS/s": round(n / ms_d / 1e6, 1), "rotator_ms": round(ms_r, 4), "rotator_GS/s": round(n / ms_r / 1e6, 1)}), flush=True)
