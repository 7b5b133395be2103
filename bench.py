#!/usr/bin/env python
"""bench.py -- complex<float> MSamples/s through the FIR -> FFT flowgraph (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload ("fir127_fft4096_flowgraph", BASELINE.json configs[1]+[2] chained = the metric's flowgraph):
    fir_filter(127-tap Hamming low-pass, fc = 0.1, EXACT reference summation order) on complex<float>
      -> FFT block (N = 4096, Hann window, DataSet planes magnitude/phase/Re/Im)
One step = one pass of that flowgraph over one batch of synthetic samples per GPU (default 2^30, i.e. 8 GiB in,
8 GiB FIR edge, 16 GiB FFT planes, all resident in HBM; far larger than the 126 MB L2, so no flush is needed).
Independent channels replicate the flowgraph one per GPU (no collective): `value` = samples of all ranks / max time.

JSON keys beyond the base contract:
  roofline      dominant kernel = the exact FIR, which is bound by the fp32 pipe (508 separately rounded lane results per
                sample): achieved / peak in TFLOP/s of non-fused fp32 operations; the HBM view of the same kernel and the
                HBM-bound FFT block kernel are in `roofline.hbm` and `kernels`
  kernels       every kernel of the step, CUDA events on the launching stream
  e2e           the same flowgraph from pinned HOST arrays to pinned HOST arrays through the C++ host layer (gr::Graph /
                gr::scheduler::Simple, tests/cpp/bm_flowgraph.cpp called in-process), copies inside the timed region;
                `link_ceiling_gbs` is this box's host link measured in the same run with the same traffic mix;
                `variants` = the Python mirror of the same graph, int16 I/Q input, magnitude-plane-only output
  workloads     the other BASELINE configs in the same run: #1 host plumbing, #4 DDC (one channel per GPU), #5 the
                256-channel polyphase channelizer pipelined over the GPUs (N >= 2), and the streaming chunk sweep
  cpu_baseline  the reference's own CPU code (oracle/_ref) timed on this box's cores (N = 1 only)
"""
import argparse
import ctypes as C
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "complex<float> MSamples/s through FIR->FFT flowgraph"
UNIT = "MSamples/s"
NFFT = 4096
NTAPS = 127
HBM_FALLBACK_GBS = 6650.0
# dram__bytes_read + dram__bytes_write of the FIR kernel from the committed `ncu --set full` capture
# (profiles/r01z_fir127_exact_ncu.md: 537.0 MB + 490.5 MB for 2^26 samples), per sample; algorithmic = 16 B/sample
FIR_DRAM_BYTES_PER_SAMPLE = (536.965632e6 + 490.491392e6) / (1 << 26)
FP32_LANES_PER_SM = 128


def workload_config(samples_per_gpu, world):
    """The `config` object, identical in both arms (the reference arm runs bounded samples of this workload)."""
    return {"workload": "fir127_fft4096_flowgraph", "fir_taps": NTAPS, "fir_mode": "exact(reference summation order)", "fft_size": NFFT, "window": "Hann", "fft_output": "DataSet planes mag/phase/re/im",
            "samples_per_gpu_per_step": samples_per_gpu, "parallelism": f"{world} independent channel(s), one flowgraph per GPU, no collective", "l2": "inputs (8 GiB/GPU) exceed L2; no flush needed"}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            peaks = json.load(f)
        return float(peaks.get("hbm_gbs", HBM_FALLBACK_GBS)), "measured (MEASURED_PEAKS.json)", float(peaks.get("sm_max_mhz", 1965.0))
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)", 1965.0


def load_script(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "scripts", name + ".py"))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device = device
        self.rows = []
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-i", str(self.device)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=10)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None, "reasons": reasons, "samples": len(self.rows)}


def synthetic_input_torch(n, device, seed):
    """re, im ~ U(-1, 1): counter-based (torch Philox) generator, regenerable from (seed, n)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    x = torch.empty(n, dtype=torch.complex64, device=device)
    torch.view_as_real(x).uniform_(-1.0, 1.0, generator=g)
    return x


# --------------------------------------------------------------------------------------------------------------------
# CPU reference arm / baseline: the reference's own code (oracle/_ref) or, if that was never built, the oracle port
# --------------------------------------------------------------------------------------------------------------------
def cpu_flowgraph(samples_per_thread, threads, budget_s=0.0):
    """FIR(127) -> FFT block(4096, Hann) on `threads` independent channels, each thread repeating its pass until
    `budget_s` seconds are used; returns (MSamples/s, kind, seconds, samples processed)."""
    from tests import _oracle

    ref = _oracle.load_ref()
    lib, kind = (ref, "reference") if ref is not None else (_oracle.load_oracle(), "port")
    oracle = _oracle.load_oracle()
    taps = oracle.fir_generate(NTAPS, "Hamming", 0.1)
    window = oracle.window("Hann", NFFT)
    rng = np.random.default_rng(0)
    n = samples_per_thread // NFFT * NFFT
    inputs = [(rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64) for _ in range(threads)]

    passes = [0] * threads

    def work(i, deadline):
        while True:
            y = lib.fir(taps, inputs[i])
            lib.fft_block(y, NFFT, window, want_ranges=False)
            passes[i] += 1
            if time.perf_counter() >= deadline:
                break

    lib.fft_block(lib.fir(taps, inputs[0][: NFFT * 4]), NFFT, window, want_ranges=False)  # warm-up: plans, page faults
    t0 = time.perf_counter()
    pool = [threading.Thread(target=work, args=(i, t0 + budget_s)) for i in range(threads)]
    for t in pool:
        t.start()
    for t in pool:
        t.join()
    seconds = time.perf_counter() - t0
    samples = n * sum(passes)
    return samples / seconds / 1e6, kind, seconds, samples


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    per_thread = 1 << 21
    times, counts, kind = [], [], "port"
    for _ in range(args.warmup):
        cpu_flowgraph(per_thread // 8, cores)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        _, kind, seconds, samples = cpu_flowgraph(per_thread, cores, budget_s=2.0)  # one step = a bounded ~2-4 s sample
        times.append(seconds)
        counts.append(samples)
    total = time.perf_counter() - t_all
    n = sum(counts) // max(len(counts), 1)
    value = sum(counts) / sum(times) / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / max(len(times), 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.samples // NFFT * NFFT, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": f"the reference's own FIR / FFT / window / magnitude / phase code (oracle/_ref, release flags -O2) on {cores} independent channels, one per host thread, each repeating passes of {per_thread // NFFT * NFFT} samples for >= 2 s per step ({n} samples per step on average) -- a bounded sample of the workload in `config`"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": total,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------------------------
def partition_cores(local_rank, world):
    """Several ranks on one host: every rank gets its own slice of the cores next to its GPU (or of all allowed cores when
    the topology puts every GPU on one node), so that eight launcher threads and their pinned first-touch pages do not
    pile up on the same cores. Returns the CPU list, or None if the affinity could not be changed."""
    if world <= 1:
        return None
    try:
        import torch

        allowed = sorted(os.sched_getaffinity(0))
        near = set(allowed)
        try:
            props = torch.cuda.get_device_properties(local_rank)
            path = f"/sys/bus/pci/devices/{getattr(props, 'pci_domain_id', 0):04x}:{props.pci_bus_id:02x}:{getattr(props, 'pci_device_id', 0):02x}.0/local_cpulist"
            with open(path) as f:
                text = f.read().strip()
            cpus = set()
            for part in text.split(","):
                if "-" in part:
                    lo, hi = part.split("-")
                    cpus.update(range(int(lo), int(hi) + 1))
                elif part:
                    cpus.add(int(part))
            if cpus & near:
                near &= cpus
        except Exception:
            pass
        pool = sorted(near)
        per = max(1, len(pool) // world)
        mine = pool[(local_rank * per) % len(pool) : (local_rank * per) % len(pool) + per] or pool
        os.sched_setaffinity(0, set(mine))
        return mine
    except Exception:
        return None


class FlowgraphLibrary:
    """tests/cpp/bm_flowgraph.cpp as a shared library: the C++ host layer's FIR -> FFT flowgraph, called in-process."""

    def __init__(self):
        path = os.path.join(ROOT, "build", "cpp", "libbm_flowgraph.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp")], check=True, capture_output=True)
        self.lib = C.CDLL(path)
        self.lib.bm_flowgraph_host.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]
        self.lib.bm_flowgraph_device.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]

    def host(self, device, variant, host_in, n_samples, host_out, chunk):
        seconds, setup, frames, err = C.c_double(), C.c_double(), C.c_size_t(), C.create_string_buffer(512)
        rc = self.lib.bm_flowgraph_host(device, variant, host_in, n_samples, host_out, chunk, C.byref(seconds), C.byref(setup), C.byref(frames), err, 512)
        if rc != 0:
            raise RuntimeError("bm_flowgraph_host: " + err.value.decode())
        return seconds.value, frames.value

    def device(self, device, capture_ptr, capture_size, n_samples, chunk):
        seconds, setup, frames, err = C.c_double(), C.c_double(), C.c_size_t(), C.create_string_buffer(512)
        rc = self.lib.bm_flowgraph_device(device, capture_ptr, capture_size, n_samples, chunk, C.byref(seconds), C.byref(setup), C.byref(frames), err, 512)
        if rc != 0:
            raise RuntimeError("bm_flowgraph_device: " + err.value.decode())
        return seconds.value, frames.value


def run_ours(args):
    import torch

    import gnuradio4_b200 as gr4
    from gnuradio4_b200 import multigpu

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dom = f"gpu:cuda:{local_rank}"
    lib = gr4.load()
    cores = partition_cores(local_rank, world)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def slowest(v):
        return multigpu.max_over_ranks(v, device)

    if args.workload == "ddc_fft":  # config #4 alone, as its own line
        ddc = measure_ddc(args, gr4, torch, rank, world, device, dom, barrier, slowest, local_rank)
        if rank == 0:
            print(json.dumps(ddc["line"]))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- headline: the two kernels of the flowgraph over one resident batch ------------------------------------------------
    n = args.samples // NFFT * NFFT
    taps = gr4.fir_generate(NTAPS, "Hamming", 0.1)
    fir = gr4.fir_filter(b=taps, exact=not args.fast_fir, compute_domain=dom)
    fft = gr4.FFT(fftSize=NFFT, window="Hann", compute_domain=dom)
    x = synthetic_input_torch(n, device, seed=0x67723462 + rank)  # channel = rank: independent streams
    y = torch.empty_like(x)
    sig = torch.empty((n // NFFT, 4, NFFT), dtype=torch.float32, device=device)

    def step():
        fir.process_bulk(x, out=y)
        fft.process_bulk(y, signals=sig)

    for _ in range(args.warmup):
        step()
    barrier()

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        launches_before = lib.gr4b200_launch_count()
        start.record()
        for k in range(args.steps):
            ev[k][0].record()
            fir.process_bulk(x, out=y)
            ev[k][1].record()
            fft.process_bulk(y, signals=sig)
            ev[k][2].record()
        stop.record()
        launches = lib.gr4b200_launch_count() - launches_before
        barrier()
    total_ms = slowest(start.elapsed_time(stop))
    fir_ms = slowest(sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps)
    fft_ms = slowest(sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps)

    # the same flowgraph with the two blocks merged into ONE kernel (gr4b200_fir_fft_block_cf32, the reference's
    # compile-time Merge applied on the device): reported next to the two-kernel step, not used for `value`
    merged = gr4.FirFft(gr4.fir_filter(b=taps, exact=not args.fast_fir, compute_domain=dom), fft)
    for _ in range(2):
        merged.process_bulk(x, signals=sig)
    barrier()
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m0.record()
    for _ in range(args.steps):
        merged.process_bulk(x, signals=sig)
    m1.record()
    barrier()
    merged_ms = slowest(m0.elapsed_time(m1) / args.steps)
    ms_per_step = total_ms / args.steps
    value = n * world / (ms_per_step * 1e-3) / 1e6
    # the opt-in tolerance mode of the FIR (overlap-save through the 4096-point transform, float-transform accuracy instead of
    # the reference's rounding): the same flowgraph with it, reported next to the headline, never as `value`
    ols = gr4.fir_filter(b=taps, overlap_save=True, compute_domain=dom)
    for _ in range(2):
        ols.process_bulk(x, out=y)
        fft.process_bulk(y, signals=sig)
    barrier()
    o0, o1, o2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    o0.record()
    for _ in range(args.steps):
        ols.process_bulk(x, out=y)
    o1.record()
    for _ in range(args.steps):
        ols.process_bulk(x, out=y)
        fft.process_bulk(y, signals=sig)
    o2.record()
    barrier()
    ols_ms, ols_flow_ms = slowest(o0.elapsed_time(o1) / args.steps), slowest(o1.elapsed_time(o2) / args.steps)
    del merged, ols, y, sig

    # ---- streaming through the C++ scheduler, device resident: what chunking costs next to one launch per batch ----------
    graphs = FlowgraphLibrary()
    streaming = []
    for chunk in (1 << 16, 1 << 18, 1 << 20, 1 << 22, 1 << 24):  # 2^16: the reference's default edge size (Graph.hpp:102)
        n_stream = min(n, max(chunk * 1024, 1 << 26))
        graphs.device(local_rank, x.data_ptr(), 2 * chunk, n_stream, chunk)  # warm-up; the source fills its two-chunk ring from x once
        barrier()
        seconds = slowest(min(graphs.device(local_rank, x.data_ptr(), 2 * chunk, n_stream, chunk)[0] for _ in range(3)))
        streaming.append({"chunk_samples": chunk, "samples": n_stream, "value": n_stream * world / seconds / 1e6, "unit": UNIT, "us_per_chunk": seconds * 1e6 / (n_stream / chunk), "frac_of_one_launch_per_batch": n_stream * world / seconds / 1e6 / value})
    del x
    torch.cuda.empty_cache()

    # ---- end to end: pinned host arrays in, pinned host arrays out -----------------------------------------------------------
    e2e_n = min(n, args.e2e_samples) // NFFT * NFFT
    src = gr4.HostBuffer(e2e_n, np.complex64)
    dst = gr4.HostBuffer(e2e_n * 4, np.float32)
    rng = np.random.default_rng(rank)
    src.array.view(np.float32)[:] = rng.uniform(-1, 1, 2 * e2e_n).astype(np.float32)
    src16 = gr4.HostBuffer(2 * e2e_n, np.int16)
    src16.array[:] = (src.array.view(np.float32) * 32767.0).astype(np.int16)

    def timed_e2e(run_once):
        run_once()  # warm-up
        total = 0.0
        for _ in range(args.steps):
            barrier()
            total += run_once()
        return slowest(total) / args.steps

    # the box's link with this traffic mix (8 B up + 16 B down per sample), all ranks at once, same pinned buffers
    link = load_script("time_host_link").measure(lib, local_rank, 8 * e2e_n, 16 * e2e_n, reps=3, barrier=barrier if world > 1 else None, host_in=C.c_void_p(src.ptr), host_out=C.c_void_p(dst.ptr))
    link_ms = slowest(link["both_ms"])
    link_gbs = world * 24.0 * e2e_n / link_ms / 1e6

    e2e_s = timed_e2e(lambda: graphs.host(local_rank, 0, src.ptr, e2e_n, dst.ptr, args.e2e_chunk)[0])
    checksum = float(np.abs(dst.array[: 4 * NFFT]).sum())
    e2e_value = e2e_n * world / e2e_s / 1e6
    variants = {}
    for name, variant, source, up, down in (("int16_iq_in", 1, src16, 4, 16), ("magnitude_plane_out", 2, src, 8, 4), ("int16_iq_in_magnitude_plane_out", 3, src16, 4, 4)):
        s = timed_e2e(lambda: graphs.host(local_rank, variant, source.ptr, e2e_n, dst.ptr, args.e2e_chunk)[0])
        variants[name] = {"value": e2e_n * world / s / 1e6, "unit": UNIT, "api": "c++", "h2d_bytes_per_step": up * e2e_n, "d2h_bytes_per_step": down * e2e_n, "ms_per_step": s * 1e3, "link_gbs": world * (up + down) * e2e_n / s / 1e9}
    # the Python mirror of the same graph (gnuradio4_b200.Graph / Simple), what round 1 reported
    g = gr4.Graph()
    b1 = g.emplaceBlock(gr4.fir_filter, b=taps, exact=not args.fast_fir, compute_domain=dom)
    b2 = g.emplaceBlock(gr4.FFT, fftSize=NFFT, window="Hann", compute_domain=dom)
    g.connect(b1, b2)
    sched = gr4.Simple(g, chunk_items=args.e2e_chunk, device=local_rank)

    def python_once():
        t0 = time.perf_counter()
        sched.runAndWait(src.array, dst.array)
        return time.perf_counter() - t0

    py_s = timed_e2e(python_once)
    variants["python_mirror"] = {"value": e2e_n * world / py_s / 1e6, "unit": UNIT, "api": "gnuradio4_b200.Graph / Simple.runAndWait", "h2d_bytes_per_step": 8 * e2e_n, "d2h_bytes_per_step": 16 * e2e_n, "ms_per_step": py_s * 1e3, "launches_per_step": sched.launches}
    sched.close()
    src.close(), dst.close(), src16.close()

    # ---- the other BASELINE configs ------------------------------------------------------------------------------------------
    workloads = {"streaming_chunk_sweep": {"api": "c++ gr::Graph / gr::scheduler::Simple, capture in HBM -> fir_filter -> FFT -> device sink", "rows": streaming}}
    ddc = measure_ddc(args, gr4, torch, rank, world, device, dom, barrier, slowest, local_rank)
    if world >= 2:
        pipeline = load_script("bench_pipeline").run_pipeline(rank, world, local_rank, device, chunks=32, chunk_samples=1 << 24)
    else:
        pipeline = None

    if rank == 0:
        hbm_peak, peak_source, sm_max_mhz = load_peaks()
        sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        lane_rate = sms * FP32_LANES_PER_SM * sm_max_mhz * 1e6          # fp32 lane results per second = non-fused flop/s
        fir_gbs = 16.0 * n / (fir_ms * 1e-3) / 1e9                       # 8 B read + 8 B written per sample
        fft_gbs = 24.0 * n / (fft_ms * 1e-3) / 1e9                       # 8 B read + 16 B (4 float planes) written per sample
        # exact mode rounds product and sum separately: 254 mul + 254 add per complex sample = 508 lane results (flop);
        # fast mode fuses them: 254 FMA lane results (508 flop against the 2 flop/lane FMA peak)
        fir_lane_ops = (2 * NTAPS if args.fast_fir else 4 * NTAPS) * n / (fir_ms * 1e-3)
        fir_tflops, fir_peak_tflops = 4 * NTAPS * n / (fir_ms * 1e-3) / 1e12, (2 if args.fast_fir else 1) * lane_rate / 1e12
        kernels = [
            {"name": "firKernel<float2,256,16,%s>" % ("fast" if args.fast_fir else "exact"), "ms": fir_ms, "bound": "fp32", "algorithmic_flop": 4.0 * NTAPS * n, "achieved_tflops": fir_tflops, "peak_tflops": fir_peak_tflops, "frac_fp32": fir_tflops / fir_peak_tflops,
             "fp32_lane_results_per_s": fir_lane_ops, "algorithmic_bytes": 16.0 * n, "achieved_gbs": fir_gbs, "frac_hbm": fir_gbs / hbm_peak, "note": "AI 31.75 flop/B > ridge 11.4: separately rounded mul and add, the reference's arithmetic, cannot fuse"},
            {"name": "fftRadixKernel<4096,Block,staged>", "ms": fft_ms, "bound": "hbm", "algorithmic_bytes": 24.0 * n, "achieved_gbs": fft_gbs, "frac_hbm": fft_gbs / hbm_peak},
        ]
        config = workload_config(n, world)
        if args.fast_fir:
            config["fir_mode"] = "fast(fma)"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "roofline": {"bound": "fp32", "kernel": kernels[0]["name"], "achieved": fir_tflops, "peak": fir_peak_tflops, "unit": "TFLOP/s", "frac": fir_tflops / fir_peak_tflops, "traffic": FIR_DRAM_BYTES_PER_SAMPLE * n,
                         "peak_source": f"{sms} SMs x {FP32_LANES_PER_SM} fp32 lanes x {sm_max_mhz:.0f} MHz (sm_max_mhz from MEASURED_PEAKS.json), one separately rounded operation per lane and clock; packed FMUL2/FFMA2 issue measured at this rate, profiles/r01_ubench_f32x2.txt",
                         "hbm": {"achieved": fir_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": fir_gbs / hbm_peak, "peak_source": peak_source},
                         "note": "the dominant kernel (direct-form 127-tap FIR with the reference's rounding: 508 flop per 16 B) is bound by the fp32 pipe, not by HBM; kernels[1], the FFT block kernel, is the HBM-bound one"},
            "kernels": kernels,
            "merged_fir_fft_kernel": {"ms_per_step": merged_ms, "value": n * world / (merged_ms * 1e-3) / 1e6, "unit": UNIT, "algorithmic_bytes": 24.0 * n, "note": "FIR and FFT block as one kernel (filtered stream stays in shared memory, bit-identical planes): HBM traffic 24 instead of 40 B/sample, but the FFT's arithmetic then competes for the fp32 pipe the FIR is bound by"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * e2e_n, "d2h_bytes_per_step": 16 * e2e_n, "samples_per_step": e2e_n, "ms_per_step": e2e_s * 1e3, "steps": args.steps, "api": "c++",
                    "path": "gr::Graph / gr::scheduler::Simple::runAndWait (tests/cpp/bm_flowgraph.cpp, in-process): pinned host array -> cuda::HostSource -> fir_filter -> FFT -> cuda::HostSink -> pinned host array, 3 streams",
                    "link_ceiling_gbs": link_gbs, "link_gbs": world * 24.0 * e2e_n / e2e_s / 1e9, "frac_of_link": (world * 24.0 * e2e_n / e2e_s / 1e9) / link_gbs,
                    "link_ceiling_note": "cudaMemcpyAsync of the same pinned buffers, 8 B up + 16 B down per sample, both directions at once on two streams, all ranks at the same time (scripts/time_host_link.py); a separate measurement next to the e2e run: the two repeat to about 1 %, so frac_of_link can read slightly above 1",
                    "cores_per_rank": len(cores) if cores else None, "checksum": checksum, "variants": variants},
            "workloads": workloads,
            "gpu_launches": int(launches),  # counted by the library at every kernel launch inside the timed region
            "clocks": clocks.summary(),
        }
        workloads["fir_overlap_save_mode"] = {"note": "opt-in tolerance mode of fir_filter (y = IFFT(FFT(x) . FFT(b)) on blocks of 4096): bound by memory traffic and the transform passes instead of the fp32 pipe; agrees with the exact mode to ~1.4e-7 * sum|b| * max|x| (tests/test_gpu_parity.py), not bit-identical, hence not the headline",
                                              "fir_kernel": {"name": "firOverlapSaveKernel", "ms": ols_ms, "value": n * world / (ols_ms * 1e-3) / 1e6, "unit": UNIT, "achieved_gbs": 16.0 * n / (ols_ms * 1e-3) / 1e9, "frac_hbm": 16.0 * n / (ols_ms * 1e-3) / 1e9 / hbm_peak},
                                              "flowgraph": {"ms_per_step": ols_flow_ms, "value": n * world / (ols_flow_ms * 1e-3) / 1e6, "unit": UNIT}}
        workloads["ddc"] = ddc["summary"]
        if pipeline is not None:
            workloads["pipeline"] = pipeline
        plumbing = measure_plumbing()
        if plumbing is not None:
            workloads["plumbing"] = plumbing
        if world == 1 and not args.no_cpu_baseline:
            ncores = os.cpu_count() or 1
            cpu_value, kind, seconds, samples = cpu_flowgraph(1 << 21, ncores, budget_s=12.0)
            line["cpu_baseline"] = {"value": cpu_value, "unit": UNIT, "cores": ncores, "kind": kind, "sample": f"{ncores} independent channels (one host thread each) repeating FIR(127)->FFT(4096) passes of {(1 << 21) // NFFT * NFFT} samples: {samples} samples in {seconds:.1f} s"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def measure_plumbing():
    """BASELINE config #1: NullSource -> MultiplyConst -> CountingSink, 1 000 448 complex<float>, host only, through the C++
    host scheduler (tests/cpp/qa_plumbing.cpp prints the best of 10 runs)."""
    binary = os.path.join(ROOT, "build", "cpp", "qa_plumbing")
    try:
        out = subprocess.run([binary], capture_output=True, text=True, timeout=120).stdout
        for row in out.splitlines():
            if "config1_plumbing_msamples_per_s=" in row:
                v = float(row.split("config1_plumbing_msamples_per_s=")[1].split()[0])
                return {"workload": "null_source->multiply_const->counting_sink, 1000448 complex<float>, host only, gr::scheduler::Simple", "value": v, "unit": UNIT, "reference_published": "87-162 MS/s for float chains (docs/USER_API_Connecting_Blocks.md:208-209)"}
    except Exception:
        pass
    return None


def measure_ddc(args, gr4, torch, rank, world, device, dom, barrier, slowest, local_rank):
    """BASELINE config #4: one DDC channel per GPU, Rotator(channel c: 2 pi (0.05 + 0.01 c)) -> decimating FIR (127 taps, /8,
    exact) as ONE fused kernel -> FFT block 4096. Throughput is counted in INPUT samples."""
    n = args.samples // (8 * NFFT) * (8 * NFFT)
    taps = gr4.fir_generate(NTAPS, "Hamming", 0.05)
    dphi = float(np.float32(2 * np.pi * (0.05 + 0.01 * rank)))
    ddc = gr4.DDC(gr4.Rotator(phase_increment=dphi, compute_domain=dom), gr4.fir_filter(b=taps, decimate=8, compute_domain=dom))
    fft = gr4.FFT(fftSize=NFFT, window="Hann", compute_domain=dom)
    x = synthetic_input_torch(n, device, seed=0x67723462 + rank)
    z = torch.empty(n // 8, dtype=torch.complex64, device=device)
    sig = torch.empty((n // 8 // NFFT, 4, NFFT), dtype=torch.float32, device=device)
    steps = args.steps
    for _ in range(args.warmup):
        ddc.process_bulk(x, out=z)
        fft.process_bulk(z, signals=sig)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        start.record()
        for k in range(steps):
            ev[k][0].record()
            ddc.process_bulk(x, out=z)
            ev[k][1].record()
            fft.process_bulk(z, signals=sig)
            ev[k][2].record()
        stop.record()
        barrier()
    total_ms = slowest(start.elapsed_time(stop))
    ddc_ms = slowest(sum(e[0].elapsed_time(e[1]) for e in ev) / steps)
    fft_ms = slowest(sum(e[1].elapsed_time(e[2]) for e in ev) / steps)
    ms_per_step = total_ms / steps
    hbm_peak, peak_source, sm_max_mhz = load_peaks()
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    lane_rate = sms * FP32_LANES_PER_SM * sm_max_mhz * 1e6
    ddc_gbs = 9.0 * n / (ddc_ms * 1e-3) / 1e9  # 8 B read + 1 B written per input sample (fused: the mixed stream never reaches HBM)
    fft_gbs = 24.0 * (n // 8) / (fft_ms * 1e-3) / 1e9
    fir_flop = 4.0 * NTAPS / 8.0 * n  # 127 separately rounded mul + add per kept output (re and im), one output per 8 inputs
    value = n * world / (ms_per_step * 1e-3) / 1e6
    roofline = {"bound": "hbm", "kernel": "firDecimKernel<Mix> (fused mixer + FIR/8)", "achieved": ddc_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ddc_gbs / hbm_peak, "traffic": None, "peak_source": peak_source,
                "fp32": {"achieved_tflops": fir_flop / (ddc_ms * 1e-3) / 1e12, "peak_tflops": lane_rate / 1e12, "frac": fir_flop / (ddc_ms * 1e-3) / lane_rate},
                "note": "9 algorithmic bytes per input sample; the kernel also carries the bit-exact phase replay, the library-exact sin/cos on the FP64 pipe and 63.5 non-fused fp32 flop per input sample, see DESIGN.md"}
    kernels = [{"name": "fused DDC (firDecimKernel<Mix>)", "ms": ddc_ms, "achieved_gbs": ddc_gbs, "frac_hbm": ddc_gbs / hbm_peak}, {"name": "fftRadixKernel<4096,Block,staged>", "ms": fft_ms, "achieved_gbs": fft_gbs, "frac_hbm": fft_gbs / hbm_peak}]
    config = {"workload": "ddc_mixer_fir127_decim8_fft4096", "input_samples_per_gpu_per_step": n, "parallelism": f"{world} independent channel(s), one per GPU, no collective", "l2": "inputs (8 GiB/GPU) exceed L2; no flush needed"}
    line = {"metric": "complex<float> input MSamples/s through the DDC flowgraph (mixer -> FIR/8 -> FFT 4096)", "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "roofline": roofline, "kernels": kernels, "gpu_launches": None, "clocks": clocks.summary()}
    summary = {"workload": config["workload"], "value": value, "unit": "input " + UNIT, "ms_per_step": ms_per_step, "input_samples_per_gpu_per_step": n, "roofline": roofline, "kernels": kernels}
    del x, z, sig
    torch.cuda.empty_cache()
    return {"line": line, "summary": summary}


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--samples", type=int, default=1 << 30, help="complex samples per GPU per step")
    p.add_argument("--e2e-samples", type=int, default=1 << 28)
    p.add_argument("--e2e-chunk", type=int, default=1 << 22, help="samples per work chunk of the end-to-end flowgraph run")
    p.add_argument("--fast-fir", action="store_true", help="FMA FIR (tolerance mode) instead of the bit-exact default")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--workload", default="fir_fft", choices=["fir_fft", "ddc_fft"], help="fir_fft = the metric's flowgraph (default, carries the other configs under `workloads`); ddc_fft = BASELINE config #4 as its own line")
    args = p.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
