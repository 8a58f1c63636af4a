#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` log into the launch list kept under profiles/:
one line per launch (id, kernel, grid, block, duration in us) + a per-kernel summary with each kernel's share.

    python tools/launch_list.py gpurun_out/p6_launches.csv "comment for the header" > profiles/r2_launch_list_box_512.csv
"""
import collections
import csv
import io
import sys

src, comment = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = list(csv.DictReader(io.StringIO("".join(ln for ln in open(src) if ln.startswith('"')))))
rows = [r for r in rows if r["Metric Name"] == "gpu__time_duration.sum"]
print(f"# {comment}")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r["Kernel Name"].replace("void ", "").replace("(T2)", "")
    us = float(r["Metric Value"].replace(",", "")) / 1e3
    tot[name][0] += 1
    tot[name][1] += us
all_us = sum(v[1] for v in tot.values())
print("# per kernel: launches, total us, share of all launches' time (serialised, cold-cache per-launch times: shares, not absolutes)")
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"# {name}: {n} launches, {us:.1f} us, {100 * us / all_us:.2f} %")
print("id,kernel,grid,block,duration_us")
for r in rows:
    name = r["Kernel Name"].replace("void ", "").replace("(T2)", "")
    print(f'{r["ID"]},"{name}","{r["Grid Size"]}","{r["Block Size"]}",{float(r["Metric Value"].replace(",", "")) / 1e3:.2f}')
