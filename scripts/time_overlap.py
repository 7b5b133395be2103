"""Does running the FIR of chunk k+1 next to the FFT block of chunk k (two streams) beat the serial step? (GPU box)"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnuradio4_b200 as gr4

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 29
nfft = 4096
x = torch.empty(n, dtype=torch.complex64, device="cuda")
torch.view_as_real(x).uniform_(-1, 1)
y = torch.empty_like(x)
sig = torch.empty((n // nfft, 4, nfft), dtype=torch.float32, device="cuda")
taps = gr4.fir_generate(127, "Hamming", 0.1)
fft = gr4.FFT(fftSize=nfft, window="Hann")


def run(chunks, overlap):
    fir = gr4.fir_filter(b=taps)
    # OVERLAP_PRIORITY=1: the FFT's stream gets the higher priority, so that its CTAs are dispatched into free slots ahead of
    # the FIR's pending waves (the block scheduler otherwise drains the earlier grid first)
    a, b = torch.cuda.Stream(), torch.cuda.Stream(priority=-1 if os.environ.get("OVERLAP_PRIORITY") == "1" else 0)
    c = n // chunks
    per = c // nfft

    def step():
        done = []
        for k in range(chunks):
            with torch.cuda.stream(a):
                fir.process_bulk(x[k * c : (k + 1) * c], out=y[k * c : (k + 1) * c])
                e = torch.cuda.Event()
                e.record()
            with torch.cuda.stream(b if overlap else a):
                if overlap:
                    b.wait_event(e)
                fft.process_bulk(y[k * c : (k + 1) * c], signals=sig[k * per : (k + 1) * per])
        torch.cuda.current_stream().wait_stream(a)
        torch.cuda.current_stream().wait_stream(b)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(3):
        step()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 3
    print(json.dumps({"chunks": chunks, "two_streams": overlap, "ms": round(ms, 3), "GS/s": round(n / ms / 1e6, 2)}), flush=True)
    return sig[:: max(1, per)].clone()


ref = run(1, False)
for chunks in (8, 32) if os.environ.get("OVERLAP_CHUNKS") is None else [int(c) for c in os.environ["OVERLAP_CHUNKS"].split(",")]:
    for overlap in (False, True):
        got = run(chunks, overlap)
print(json.dumps({"planes_identical_across_variants": bool(torch.equal(ref[0], got[0]))}))
