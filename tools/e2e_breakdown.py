#!/usr/bin/env python
"""Where the end-to-end step (host buffers through the C ABI) spends its time, on the default bench workload.

Three loops of K steps each, wall clock per step (incl. the final fg_sync) next to per-call host times and the DEVICE time
of the calls (FgStats.last_step_ms: events around each fg_step on the library's stream, read lazily):
  A  fg_step(1) only                         markers static: index map and band reused, 3 IB kernels per step
  B  fg_step(1) + fg_get_link_wrenches       the same + the read-out
  C  fg_set_markers + fg_step(1) + get       what bench.py's `e2e` times: markers re-sent, 5 IB kernels + H2D per step
A vs fg_step(K) (bench `value`) is the cost of returning to the caller every step; C vs B is the cost of re-sending markers.
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gym_fish_b200 as g
import bench

flags = 0
for a in sys.argv[1:]:
    flags |= {"--no-split": g._abi.FLAG_NO_SPLIT, "--no-graphs": g._abi.FLAG_NO_GRAPHS, "--no-flip": g._abi.FLAG_NO_SWEEP_FLIP,
              "--sync-step": g._abi.FLAG_SYNC_STEP}[a]
sim, markers = bench.make_sim(g, os.environ.get("FG_E2E_LIB", "cuda"), "sphere_256x128x128", 0, 1, 0, flags=flags)   # FG_E2E_LIB: dry runs
X, U, dV, link, _ = markers
pin = [np.ascontiguousarray(a) for a in (X, U, dV, link)]
cells = sim.stats().cells
K = int(os.environ.get("FG_E2E_STEPS", "3000"))
sim.step(min(200, K))
sim.sync()
pc = time.perf_counter
W = min(100, K)        # untimed steps at the start of each loop

t0 = pc(); sim.step(K); sim.sync(); ref = (pc() - t0) / K
print(f"flags {flags}: fg_step({K}) wall per step {ref*1e6:.1f} us = {cells/ref/1e6:.0f} MLUPS (device {sim.stats().last_step_ms/K*1e3:.1f} us)")


def loop(name, send, read):
    t = np.zeros(4)
    dev, ndev = 0.0, 0
    for i in range(K + W):
        if i == W:
            sim.sync(); t_begin = pc()
        a = pc()
        if send:
            sim.set_markers(*pin)
        b = pc(); sim.step(1)
        c = pc()
        if read:
            sim.get_link_wrenches()
        d = pc()
        if i >= W:
            t += (b - a, c - b, d - c, d - a)
            if i % 50 == 0:                       # device time of the latest call that has finished
                dev += sim.stats().last_step_ms; ndev += 1
    sim.sync()
    wall = (pc() - t_begin) / K
    print(f"{name}: wall per step {wall*1e6:.1f} us = {cells/wall/1e6:.0f} MLUPS | host: set_markers {t[0]/K*1e6:.1f}  step {t[1]/K*1e6:.1f}  "
          f"get_wrenches {t[2]/K*1e6:.1f}  sum {t[3]/K*1e6:.1f} | device per call {dev/max(ndev,1)*1e3:.1f} us")


loop("A step only          ", False, False)
loop("B step + read        ", False, True)
loop("C send + step + read ", True, True)
