#!/usr/bin/env python
"""NVLink bytes of the halo push, counted by ncu (VERDICT r1 missing #7): ONE process holds two peered z-slabs of a periodic
512 x 512 box on GPUs 0 and 1 (same-process peers: raw pointers + cudaDeviceEnablePeerAccess — the kernels and the wire are
those of the multi-process run, only the handle exchange differs), stepped by two threads.

    ncu --kernel-name-base demangled -k regex:ZFaceOp --metrics nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,gpu__time_duration.sum \
        --csv --log-file gpurun_out/nvlink_halo.csv python tools/nvlink_halo.py

Expected per ZFaceOp launch of a rank: 5 outgoing populations x 512 x 512 cells x 4 B = 5 242 880 B per internal face.
"""
import os, sys, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gym_fish_b200 as g

nz_local, steps = int(os.environ.get("FG_NVL_NZ", "16")), int(os.environ.get("FG_NVL_STEPS", "4"))
kw = dict(nx=512, ny=512, nz=2 * nz_local, tau=0.6, collision=g.MRT)
parts = [g.Sim(backend="cuda", device=r, n_ranks=2, rank=r, **kw) for r in range(2)]
handles = [s.peer_export() for s in parts]
for r, s in enumerate(parts):
    s.peer_connect(handles[1 - r], handles[1 - r])       # periodic in z: both neighbours are the other rank
rng = np.random.default_rng(3)
for s in parts:
    u = (1e-3 * rng.standard_normal((3,) + s.shape)).astype(np.float32)
    s.set_fields(np.ones(s.shape, np.float32), u)
errs = []


def run(s):
    try:
        s.step(steps)
        s.sync()
    except Exception as e:      # noqa: BLE001
        errs.append(e)


ts = [threading.Thread(target=run, args=(s,)) for s in parts]
[t.start() for t in ts]
[t.join() for t in ts]
if errs:
    raise errs[0]
print("halo bytes per face per step:", 5 * 512 * 512 * 4, "launches", [s.stats().kernel_launches for s in parts])
