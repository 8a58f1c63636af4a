#!/usr/bin/env python
"""Aggregate env steps/s of N small fish tanks sharing one B200 (VectorFishEnv, one host thread and four CUDA streams
per env) against the same envs stepped one after the other.  Small tanks cannot fill the GPU alone: their immersed-
boundary chain and host body integration are latency, which another env's collide hides."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gym_fish_b200 as g
from gym_fish_b200.env import EnvConfig, FishEnv, FishSpec, VectorFishEnv

def run(n_envs, grid, steps=12, substeps=20):
    spec = FishSpec(links=((20, 5), (18, 5), (16, 4), (14, 3)))
    cfgs = [EnvConfig(grid=grid, n_substeps=substeps, fish=(spec,)) for _ in range(n_envs)]
    out = {}
    vec = VectorFishEnv(cfgs, backend=os.environ.get("FG_BACKEND", "cuda"))
    vec.reset(seed=1)
    acts = np.zeros((n_envs, vec.action_space.shape[0]), np.float32)
    for mode in ("sequential", "threads"):
        for it in range(steps + 2):
            if it == 2:
                t0 = time.perf_counter()
            acts[:] = np.sin(0.3 * it + np.arange(acts.shape[1]))[None, :]
            if mode == "threads":
                vec.step(acts)
            else:
                for e, a in zip(vec.envs, acts):
                    e.step(a)
        dt = time.perf_counter() - t0
        cells = grid[0] * grid[1] * grid[2]
        out[mode] = dict(env_steps_per_s=n_envs * steps / dt, mlups=n_envs * steps * substeps * cells / dt / 1e6)
    vec.close()
    return out

if __name__ == "__main__":
    if os.environ.get("FG_BACKEND"):
        print(run(2, (24, 24, 64), steps=2, substeps=2))
        sys.exit(0)
    for n_envs, grid in ((1, (128, 128, 256)), (4, (128, 128, 256)), (8, (128, 128, 256)), (8, (96, 96, 192)), (16, (64, 64, 128))):
        r = run(n_envs, grid)
        print(json.dumps(dict(n_envs=n_envs, grid_xyz=grid, substeps=20, **{k: {kk: round(vv, 1) for kk, vv in v.items()} for k, v in r.items()})), flush=True)
