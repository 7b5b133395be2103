"""Host <-> device link ceiling of this box, measured the way the end-to-end flowgraph uses it: pinned host buffers,
cudaMemcpyAsync through the C ABI (gr4b200_copy_h2d / _d2h), one stream per direction, both directions at once.

    python scripts/time_host_link.py                      # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/time_host_link.py   # N ranks at once

Prints one JSON line (rank 0): per-GPU and whole-box GB/s for H2D alone, D2H alone, and the concurrent mix `--mix a:b`
(default 1:2 = the FIR -> FFT flowgraph's 8 B in + 16 B out per sample), plus the samples/s ceiling that mix implies.
`measure()` is imported by bench.py, which puts the ceiling next to its end-to-end figure."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(lib, device_index, h2d_bytes, d2h_bytes, reps=4, chunk_bytes=64 << 20, barrier=None, host_in=None, host_out=None):
    """Returns {"h2d_gbs", "d2h_gbs", "both_gbs", "both_ms"} for this process's GPU. The copies go out in `chunk_bytes`
    pieces alternating between the two streams' queues, as a streaming flowgraph issues them. `barrier` (callable) lines
    several ranks up so that they load the host at the same time."""
    import torch

    torch.cuda.set_device(device_index)
    vp = C.c_void_p
    own_in, own_out = host_in is None, host_out is None
    host_in = vp(lib.gr4b200_malloc_host(h2d_bytes)) if own_in else host_in
    host_out = vp(lib.gr4b200_malloc_host(d2h_bytes)) if own_out else host_out
    dev_in = torch.empty(h2d_bytes, dtype=torch.uint8, device="cuda")
    dev_out = torch.zeros(d2h_bytes, dtype=torch.uint8, device="cuda")
    if own_in:  # (a caller's buffers keep their content: bench.py measures with the arrays its flowgraph run uses)
        C.memset(host_in, 1, h2d_bytes)
    if own_out:
        C.memset(host_out, 0, d2h_bytes)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def copies(direction_in, direction_out):
        offs_in = list(range(0, h2d_bytes, chunk_bytes)) if direction_in else []
        offs_out = list(range(0, d2h_bytes, chunk_bytes)) if direction_out else []
        k_in = k_out = 0
        while k_in < len(offs_in) or k_out < len(offs_out):  # interleave proportionally
            if k_in < len(offs_in) and (k_out >= len(offs_out) or k_in * max(len(offs_out), 1) <= k_out * max(len(offs_in), 1)):
                o = offs_in[k_in]
                lib.gr4b200_copy_h2d(vp(dev_in.data_ptr() + o), vp(host_in.value + o), min(chunk_bytes, h2d_bytes - o), vp(s_in.cuda_stream))
                k_in += 1
            else:
                o = offs_out[k_out]
                lib.gr4b200_copy_d2h(vp(host_out.value + o), vp(dev_out.data_ptr() + o), min(chunk_bytes, d2h_bytes - o), vp(s_out.cuda_stream))
                k_out += 1

    def timed(direction_in, direction_out):
        copies(direction_in, direction_out)  # warm-up
        torch.cuda.synchronize()
        if barrier is not None:
            barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(s_in)
        s_out.wait_event(e0)
        for _ in range(reps):
            copies(direction_in, direction_out)
        e1.record(s_in)
        e2.record(s_out)
        torch.cuda.synchronize()
        return max(e0.elapsed_time(e1), e0.elapsed_time(e2)) / reps

    ms_in, ms_out, ms_both = timed(True, False), timed(False, True), timed(True, True)
    if own_in:
        lib.gr4b200_free_host(host_in)
    if own_out:
        lib.gr4b200_free_host(host_out)
    return {"h2d_gbs": h2d_bytes / ms_in / 1e6, "d2h_gbs": d2h_bytes / ms_out / 1e6, "both_gbs": (h2d_bytes + d2h_bytes) / ms_both / 1e6, "both_ms": ms_both}


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--mbytes", type=int, default=768, help="total MiB per repetition and GPU, split by --mix")
    p.add_argument("--mix", default="1:2", help="H2D:D2H byte ratio of the concurrent leg")
    p.add_argument("--reps", type=int, default=4)
    p.add_argument("--bytes-per-sample", type=float, default=24.0, help="H2D + D2H bytes one sample of the workload moves")
    args = p.parse_args()
    import torch

    import gnuradio4_b200 as gr4

    lib = gr4.load()
    rank, world, local_rank = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    barrier = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

        def barrier():
            dist.barrier()
            torch.cuda.synchronize()

    a, b = (float(v) for v in args.mix.split(":"))
    total = args.mbytes << 20
    h2d_bytes = int(total * a / (a + b)) // (1 << 20) * (1 << 20)
    d2h_bytes = total - h2d_bytes
    r = measure(lib, local_rank, h2d_bytes, d2h_bytes, reps=args.reps, barrier=barrier)
    vals = torch.tensor([r["h2d_gbs"], r["d2h_gbs"], r["both_gbs"]], dtype=torch.float64, device="cuda")
    slowest = torch.tensor([r["both_ms"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals)  # sums over the ranks running at the same time
        dist.all_reduce(slowest, op=dist.ReduceOp.MAX)
    if rank == 0:
        box_both = world * (h2d_bytes + d2h_bytes) / slowest.item() / 1e6  # whole box, the slowest rank defines the step
        print(json.dumps({"n_gpus": world, "mix_h2d_d2h": args.mix, "bytes_per_rep_per_gpu": total, "h2d_alone_gbs_sum": vals[0].item(), "d2h_alone_gbs_sum": vals[1].item(), "both_gbs_sum": vals[2].item(), "both_gbs_slowest_rank": box_both,
                          "link_ceiling_gbs": box_both, "ceiling_msamples_per_s": box_both * 1e3 / args.bytes_per_sample, "cpus": os.cpu_count(), "affinity": sorted(os.sched_getaffinity(0))[:4] + ["..."] if len(os.sched_getaffinity(0)) > 4 else sorted(os.sched_getaffinity(0))}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
