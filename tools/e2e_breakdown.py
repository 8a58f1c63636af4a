#!/usr/bin/env python
"""Where the end-to-end step (host buffers through the C ABI) spends its time: per-call wall time of
fg_set_markers / fg_step(1) / fg_get_link_wrenches on the default bench workload, next to the device time of the step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gym_fish_b200 as g
import bench

flags = 0
for a in sys.argv[1:]:
    flags |= {"--no-split": g._abi.FLAG_NO_SPLIT, "--no-graphs": g._abi.FLAG_NO_GRAPHS, "--no-flip": g._abi.FLAG_NO_SWEEP_FLIP}[a]
sim, markers = bench.make_sim(g, "cuda", "sphere_256x128x128", 0, 1, 0, flags=flags)
X, U, dV, link, _ = markers
pin = [np.ascontiguousarray(a) for a in (X, U, dV, link)]
sim.step(200)
K = 3000
t = np.zeros(4)
dev = 0.0
pc = time.perf_counter
for i in range(K + 100):
    t0 = pc(); sim.set_markers(*pin)
    t1 = pc(); sim.step(1)
    t2 = pc(); w = sim.get_link_wrenches()
    t3 = pc()
    if i == 100:
        sim.sync(); t_begin = pc()
    if i >= 100:
        t += (t1 - t0, t2 - t1, t3 - t2, t3 - t0)
sim.sync()
wall = (pc() - t_begin) / K
print(f"flags {flags}: per step us: set_markers {t[0]/K*1e6:.1f}  step {t[1]/K*1e6:.1f}  get_wrenches {t[2]/K*1e6:.1f}  total {t[3]/K*1e6:.1f}  "
      f"wall per step incl. the final sync {wall*1e6:.1f} us (fg_step returns when the wrenches are there, the collide runs on)")
