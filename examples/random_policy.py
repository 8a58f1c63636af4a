#!/usr/bin/env python
"""A Gym-style loop over the B200 library: one 5-link fish in a 512x256x256 tank (BASELINE.json configs[2]), random actions.

    python examples/random_policy.py                 # CUDA backend (needs a B200)
    python examples/random_policy.py --small         # a 48x40x96 tank, 5 substeps

--lib PATH loads another library that exports include/fishgym.h instead of the CUDA one (the test suite passes the CPU
checker there so that this script is exercised without a GPU).

Shows what a user of the reference's "control through Python interface" (README.md:14) would write: reset(), step(action)
-> obs, reward, terminated, truncated, info.  numpy only; no torch on the simulation path."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from gym_fish_b200.env import EnvConfig, FishEnv, FishSpec  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--small", action="store_true", help="48x40x96 tank, 5 substeps per env step")
ap.add_argument("--lib", default="cuda", help="library exporting the fishgym ABI (default: the CUDA product)")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--task", default="cruise", choices=["cruise", "pose", "path"])
ap.add_argument("--probes", type=int, default=0, help="fluid-velocity probes ahead of the head, appended to the observation")
ap.add_argument("--passes", type=int, default=1, help="direct-forcing passes per substep (EnvConfig.ib_iterations; > 1: multi-direct forcing)")
args = ap.parse_args()

if args.small:
    cfg = EnvConfig(grid=(48, 40, 96), tau=0.8, n_substeps=5, task=args.task, probes=args.probes,
                    fish=(FishSpec(links=((10, 3), (9, 3), (8, 2.5), (7, 2))),), waypoints=((24, 40), (30, 20)))
else:
    cfg = EnvConfig(task=args.task, probes=args.probes, waypoints=((128, 300), (160, 200)))     # 256 x 256 x 512, 20 substeps
cfg.ib_iterations = args.passes
env = FishEnv(cfg, backend=args.lib)

rng = np.random.default_rng(0)
obs, info = env.reset(seed=0)
print(f"markers {info['n_markers']}, obs {obs.shape}, actions {env.action_space.shape}")
t0, ret = time.perf_counter(), 0.0
for it in range(args.steps):
    action = rng.uniform(-1, 1, env.action_space.shape).astype(np.float32)
    obs, reward, terminated, truncated, info = env.step(action)
    ret += reward
    if terminated or truncated:
        obs, info = env.reset()
env.sim.sync()
dt = time.perf_counter() - t0
print(f"{args.steps} env steps in {dt:.2f} s = {args.steps / dt:.1f} env steps/s, return {ret:.4f}, head at {obs[:3]}")
env.close()
