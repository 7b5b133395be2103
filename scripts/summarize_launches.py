"""Turns the ncu launch list of the bench command (gpurun_out/verify_launches.csv, written by scripts/gpu_verify.sh) into
profiles/<round>_bench_launches.md: per-kernel totals and shares, then every launch."""
import csv
import re
import sys
from collections import OrderedDict

src, dst = sys.argv[1], sys.argv[2]
fir_share, fft_share = (sys.argv[3], sys.argv[4]) if len(sys.argv) > 4 else ("?", "?")
lines = [l for l in open(src) if l.startswith('"')]
rows = list(csv.DictReader(lines))


def short(name):
    name = re.sub(r"void (gr4b200::)?(\(anonymous namespace\)|<unnamed>|unnamed>)::", "", name)
    name = re.sub(r"gr4b200::", "", name)
    if "uniform" in name or "distribution" in name:
        return "at::uniform_ (torch RNG fill of the synthetic input)"
    return name.replace("void ", "")


launches = []
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    value = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = value / 1e3 if unit in ("ns", "nsecond") else (value if unit in ("us", "usecond") else value * 1e3)
    launches.append((short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], us))
totals = OrderedDict()
for name, _, _, us in launches:
    n, t = totals.get(name, (0, 0.0))
    totals[name] = (n + 1, t + us)
whole = sum(t for _, t in totals.values())
with open(dst, "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-samples 16777216`\n")
    f.write(f"# (launch list; per-launch times are serialised and cold-cache: compare the SHARES of firKernel / fftRadixKernel with bench.py's CUDA-event split, {fir_share} / {fft_share};\n")
    f.write("#  firFftBlockKernel is the merged-kernel comparison step bench.py also runs)\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
    for name, (n, t) in sorted(totals.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| {name} | {n} | {t / 1e3:.3f} | {100 * t / whole:.1f} % |\n")
    f.write("\n| id | kernel | grid | block | us |\n|---|---|---|---|---|\n")
    for i, (name, grid, block, us) in enumerate(launches):
        f.write(f"| {i} | {name} | {grid} | {block} | {us:.1f} |\n")
