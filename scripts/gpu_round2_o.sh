#!/bin/bash
# ncu --set full of the /8 FIR (default and 128 x 7 x 1), the fused DDC and the full-rate exact FIR as they are now; the
# reports are summarised on the box (they are 40 MB each with the source pages)
mkdir -p gpurun_out
cap() {
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 -s $4 -f -o /tmp/r02o_$1 python scripts/profile_kernels.py $3 > gpurun_out/r02o_$1.log 2>&1; echo "$1: exit $?"
  python scripts/summarize_ncu.py /tmp/r02o_$1.ncu-rep gpurun_out/r02o_$1_ncu.md "$5"
  # per-SASS-instruction samples: the twenty hottest lines with their stall reasons
  ncu -i /tmp/r02o_$1.ncu-rep --page source --csv --print-source sass > /tmp/r02o_$1_source.csv 2>/dev/null
  python - "$1" <<'PY'
import csv, sys
name = sys.argv[1]
rows = list(csv.reader(open(f"/tmp/r02o_{name}_source.csv", errors="replace")))
start = next(i for i, r in enumerate(rows) if "Source" in [c.strip() for c in r])
rows = rows[start:]
hdr = rows[0]
def col(n):
    for i, h in enumerate(hdr):
        if h.strip() == n:
            return i
    return None
src, smp, exe = col("Source"), col("# Samples") or col("Samples"), col("Instructions Executed")
out = open(f"gpurun_out/r02o_{name}_hot_sass.txt", "w")
out.write(" | ".join(hdr) + "\n")
body = [r for r in rows[1:] if len(r) == len(hdr)]
if smp is not None:
    tot = sum(float(r[smp] or 0) for r in body)
    body.sort(key=lambda r: -float(r[smp] or 0))
    out.write(f"total samples {tot}\n")
    for r in body[:40]:
        out.write(f"{r[smp]:>8} {r[exe] if exe is not None else '':>10} {r[src]}\n")
    # the whole listing in address order: samples, executions, instruction (compact)
    import gzip
    with gzip.open(f"gpurun_out/r02o_{name}_sass_profile.txt.gz", "wt") as g:
        for r in [r for r in rows[1:] if len(r) == len(hdr)]:
            g.write(f"{r[smp]:>8} {r[exe] if exe is not None else '':>10} {r[src]}\n")
    # opcode histogram weighted by samples and by executions
    from collections import Counter
    bys, bye = Counter(), Counter()
    for r in body:
        op = r[src].split()[0] if r[src].split() else "?"
        if op.startswith("@"):
            op = r[src].split()[1]
        op = op.split(".")[0]
        bys[op] += float(r[smp] or 0)
        if exe is not None:
            bye[op] += float(r[exe] or 0)
    out.write("\nby opcode: samples, instructions executed\n")
    for op, v in bys.most_common(30):
        out.write(f"{op:>10} {v:>10.0f} {bye[op]:>14.0f}\n")
PY
  rm -f /tmp/r02o_$1.ncu-rep
}
cap ddc firDecim ddc 2 "fused DDC, 128 x 5 single stage with carried halo, shared tap pairs, 2^26 input samples"
#GR4B200_FIR_TAP_MODE=0 cap ddc_mode0 firDecim ddc 2 "fused DDC, same with scalar shared taps"
cap firdecim firDecim firdecim 2 "decimating FIR /8 exact, 128 x 5 x 2 stages, shared tap pairs, 2^26 input samples"
#GR4B200_DECIM8_VARIANT=4 GR4B200_FIR_TAP_MODE=2 cap firdecim_v4_mode2 firDecim firdecim 2 "decimating FIR /8 exact, 128 x 7 x 1 stage, parameter taps"
#cap fir firKernel fir 2 "full-rate exact FIR, parameter taps (uniform registers), 2^26 samples"
ls -la gpurun_out/
