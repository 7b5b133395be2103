"""An edge between two GPUs as part of the producing kernel: the kernel on cuda:0 stores its output straight into a
buffer in cuda:1's HBM (peer access over NVLink) -- against the same kernel writing locally, and against kernel +
cudaMemcpyPeerAsync. One process, two GPUs (gpurun --gpus 2)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnuradio4_b200 as gr4
from gnuradio4_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
lib = _lib.load()
assert torch.cuda.device_count() >= 2, "needs two GPUs"
assert lib.gr4b200_peer_enable(0, 1) == 0, "no peer access between cuda:0 and cuda:1"
torch.cuda.set_device(0)
x = torch.empty(n, dtype=torch.complex64, device="cuda:0")
torch.view_as_real(x).uniform_(-1, 1)
local = torch.empty_like(x)
remote = torch.empty(n, dtype=torch.complex64, device="cuda:1")
taps = gr4.fir_generate(127, "Hamming", 0.1)
proto = gr4.fir_generate(256 * 12, "Kaiser", 1 / 512, beta=8.0)


def timeit(name, fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize(0)
    ms = a.elapsed_time(b) / reps
    print(json.dumps({"case": name, "ms": round(ms, 4), "GS/s": round(n / ms / 1e6, 2), "edge_GB/s": round(8 * n / ms / 1e6, 1)}), flush=True)


stream = torch.cuda.current_stream(0).cuda_stream
blocks = {
    "MultiplyConst": gr4.MultiplyConst(value=2 + 1j),
    "fir127 exact": gr4.fir_filter(b=taps),
    "pfb filter stage": None,
}
ch = gr4.PolyphaseChannelizer(proto, 256)
for name, block in blocks.items():
    if block is None:
        launch = lambda out: ch.filter_stage(x, out=out)  # noqa: E731
    else:
        launch = lambda out, block=block: block.launch(stream, x.data_ptr(), out.data_ptr(), n)  # noqa: E731
    timeit(f"{name} -> local HBM", lambda: launch(local))
    timeit(f"{name} -> peer HBM (stores over NVLink)", lambda: launch(remote))

    def two_steps():
        launch(local)
        lib.gr4b200_peer_copy(remote.data_ptr(), 1, local.data_ptr(), 0, 8 * n, stream)

    timeit(f"{name} -> local HBM, then cudaMemcpyPeerAsync", two_steps)
    # same bits either way
    launch(local)
    launch(remote)
    torch.cuda.synchronize(0)
    skip = 8192  # past the filter history the two consecutive launches do not share
    same = torch.equal(torch.view_as_real(local[skip:]).contiguous().view(torch.int32).cpu(), torch.view_as_real(remote[skip:]).contiguous().view(torch.int32).cpu())
    print(json.dumps({"case": f"{name}: peer result bit-identical", "ok": bool(same)}), flush=True)
timeit("cudaMemcpyPeerAsync alone", lambda: lib.gr4b200_peer_copy(remote.data_ptr(), 1, x.data_ptr(), 0, 8 * n, stream))
