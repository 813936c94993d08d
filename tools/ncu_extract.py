#!/usr/bin/env python
"""Turn `ncu --set full` captures into the numbers bench.py reports next to its live measurements.

    ncu -i gpurun_out/r02_conv.ncu-rep --page raw --csv > /tmp/conv.csv
    python tools/ncu_extract.py /tmp/conv.csv conv_pack1_bf16x3_4x192x640 'conv_igemm' [--sum 'loss_.*kernel']

Writes / updates profiles/ncu_metrics.json: {tag: {dram_bytes, tensor_pipe_active_pct, duration_us, kernel, source, commit}},
`dram_bytes` = dram__bytes_read.sum + dram__bytes_write.sum per launch (summed over the matching launches with --sum: the
loss roofline counts its forward and backward launch together).  The commit stamp lets a reader see at a glance whether the
figure belongs to the kernels at HEAD; bench.py never carries such numbers as literals."""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "%": 1.0, "": 1.0}


def rows(path):
    lines = [ln for ln in open(path, newline="") if not ln.startswith("==")]
    r = list(csv.reader(lines))
    head, units, body = r[0], r[1], r[2:]
    for b in body:
        yield {h: (v, u) for h, v, u in zip(head, b, units)}


def num(cell):
    v, u = cell
    return float(v.replace(",", "")) * UNIT.get(u, 1.0)


def main():
    path, tag, pattern = sys.argv[1:4]
    do_sum = "--sum" in sys.argv
    picked = [r for r in rows(path) if re.search(pattern, r["Kernel Name"][0])]
    if not picked:
        raise SystemExit("no launch matches %r" % pattern)
    if not do_sum:
        picked = picked[-1:]
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    entry = {"dram_bytes": sum(num(r["dram__bytes_read.sum"]) + num(r["dram__bytes_write.sum"]) for r in picked),
             "duration_us": sum(num(r["gpu__time_duration.sum"]) for r in picked),
             "kernel": " + ".join(r["Kernel Name"][0][:80] for r in picked),
             "source": "ncu --set full --clock-control none, %s" % os.path.basename(path), "commit": commit}
    key = "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"
    if key in picked[-1] and picked[-1][key][0] not in ("", "n/a"):
        entry["tensor_pipe_active_pct"] = num(picked[-1][key])
    out = os.path.join(ROOT, "profiles", "ncu_metrics.json")
    d = json.load(open(out)) if os.path.exists(out) else {}
    d[tag] = entry
    json.dump(d, open(out, "w"), indent=1, sort_keys=True)
    print(tag, json.dumps(entry))


if __name__ == "__main__":
    main()
