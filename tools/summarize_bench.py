#!/usr/bin/env python
"""Print one line per bench log in gpurun_out/ (helper for reading GPU passes)."""
import json, sys, glob, os
files = sys.argv[1:] or sorted(glob.glob("gpurun_out/bench*.log"))
for f in files:
    if not os.path.exists(f):
        continue
    for ln in open(f):
        if ln.startswith("=="):
            print(ln.rstrip())
        if ln.startswith("{"):
            d = json.loads(ln)
            r = d.get("roofline") or {}
            e = d.get("e2e") or {}
            c = d.get("clocks") or {}
            print(f"{os.path.basename(f):30s} value {d['value']:9.0f} ms {d['ms_per_step']:.4f} e2e {e.get('value', 0):9.0f} frac {r.get('frac', 0):.3f} "
                  f"k_ms {r.get('kernel_ms_per_launch', 0):.4f} ib_ms {r.get('ib_ms_per_step', 0):.4f} split {(d.get("run") or {}).get("launch_mode")} "
                  f"L {d.get('gpu_launches')} clk {c.get('sm_mhz')} {c.get('reasons')} env/s {e.get('env_steps_per_s')}")
