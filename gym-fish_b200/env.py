"""Gym-style environment over the C ABI: reset() / step(action) -> obs, reward, terminated, truncated, info.

/root/reference/README.md:1-3,14 promise a Gym-like Python interface to a coupled agent-fluid simulation and ship
nothing else; BASELINE.json:5 fixes the surface.  numpy only — neither gym nor gymnasium is
required (duck-typed API with a minimal Box), and no torch on the sim path.

One env step = `n_substeps` coupled lattice steps inside ONE fg_step call: the action goes down (a few floats), the
library integrates the articulated body on the host every substep, and the observation comes back (a few floats).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _abi
from ._abi import FgFishDesc, Sim


class Box:
    """Minimal stand-in for gym.spaces.Box."""

    def __init__(self, low, high, shape, dtype=np.float32, seed: Optional[int] = None):
        self.low = np.full(shape, low, dtype=dtype)
        self.high = np.full(shape, high, dtype=dtype)
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self._rng = np.random.default_rng(seed)

    def seed(self, seed: Optional[int]):
        self._rng = np.random.default_rng(seed)

    def sample(self) -> np.ndarray:
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


@dataclass
class FishSpec:
    """A 5-link / 4-joint swimmer (BASELINE.json configs[2]: ~2k markers at one marker per unit area)."""
    links: Sequence[Tuple[float, float]] = ((28, 7), (24, 7), (22, 6), (20, 5), (18, 3.5))   # (length, radius)
    root: Optional[Tuple[float, float, float]] = None     # head-link centre; default: tank centre, 1/3 along z
    heading: float = 0.0
    density_ratio: float = 1.0
    joint_gain: float = 0.2
    joint_limit: float = 0.5
    joint_rate_max: float = 0.005                         # radians per substep.  0.01 let a random policy (targets jumping by up to 2 every env step)
                                                          # drive the tail of this 112-cell body at 0.2 - 0.5 lattice units per step and the fluid
                                                          # diverged within 3 env steps (oracle, 128 x 64 x 256 tank); 0.005 held for 40 env steps
    free_root: bool = True

    def desc(self, grid) -> FgFishDesc:
        nx, ny, nz = grid
        d = FgFishDesc()
        d.n_links = len(self.links)
        for k, (length, rad) in enumerate(self.links):
            d.link_len[k], d.link_rad[k] = float(length), float(rad)
        root = self.root if self.root is not None else (nx / 2, ny / 2, nz / 3)
        for i in range(3):
            d.root_pos[i] = float(root[i])
        d.heading, d.density_ratio = float(self.heading), float(self.density_ratio)
        d.joint_gain, d.joint_limit, d.joint_rate_max = float(self.joint_gain), float(self.joint_limit), float(self.joint_rate_max)
        d.free_root = int(self.free_root)
        return d


@dataclass
class EnvConfig:
    grid: Tuple[int, int, int] = (256, 256, 512)          # (nx, ny, nz); z is the swim axis
    tau: float = 0.53
    collision: int = _abi.MRT
    n_substeps: int = 20
    max_episode_steps: int = 500
    fish: Sequence[FishSpec] = field(default_factory=lambda: (FishSpec(),))
    walls: bool = True                                    # tank walls on x, y; periodic along z
    energy_weight: float = 0.01
    device: int = 0
    task: str = "cruise"                                  # "cruise" | "pose" | "path"
    target_heading: float = 0.5                           # pose task: yaw to reach (radians)
    waypoints: Sequence[Tuple[float, float]] = ()         # path task: (x, z) points to visit in order (first fish)
    waypoint_radius: float = 12.0
    probes: int = 0                                       # velocity probes per fish, on a ring ahead of the head
    fluid_check_every: int = 1                            # env steps between fluid divergence checks (fg_check_finite; 0 = never)
    ib_iterations: int = 1                                # direct-forcing passes per substep (FgConfig.ib_iterations; > 1: multi-direct forcing)


class FishEnv:
    """Tasks (all rewards subtract energy_weight * mean(a^2)):
      cruise  swim along the initial heading; reward = progress in body lengths
      pose    turn to `target_heading`; reward = decrease of the yaw error
      path    visit `waypoints` in order; reward = decrease of the distance to the current one (+1 on arrival)
    Observations: per fish (pos3, heading, vel3, yaw rate, joint angles, joint rates) from the host integrator, then
    optionally 3*probes fluid velocities sampled by fg_probe on a ring one head-length ahead of each head."""

    metadata = {"render_modes": []}

    def __init__(self, config: Optional[EnvConfig] = None, backend: str = "cuda"):
        self.cfg = config or EnvConfig()
        nx, ny, nz = self.cfg.grid
        W, P = _abi.BC_WALL, _abi.BC_PERIODIC
        bc = [W, W, W, W, P, P] if self.cfg.walls else [P] * 6
        descs = [f.desc(self.cfg.grid) for f in self.cfg.fish]
        # one marker per unit area: capacity from the spheroid areas (+25 %)
        cap = 0
        for f in self.cfg.fish:
            for length, rad in f.links:
                a, r, p = length / 2, rad, 1.6075
                cap += int(1.25 * 4 * np.pi * (((r * r) ** p + 2 * (r * a) ** p) / 3) ** (1 / p)) + 16
        self.sim = Sim(backend=backend, nx=nx, ny=ny, nz=nz, tau=self.cfg.tau, collision=self.cfg.collision, bc=bc,
                       max_markers=cap, max_links=sum(len(f.links) for f in self.cfg.fish), device=self.cfg.device,
                       ib_iterations=int(self.cfg.ib_iterations))
        for d in descs:
            self.sim.add_fish(d)
        if self.cfg.task not in ("cruise", "pose", "path"):
            raise ValueError(f"unknown task '{self.cfg.task}'")
        self._n_act, self._n_body_obs = self.sim.action_size(), self.sim.obs_size()
        self._n_obs = self._n_body_obs + 3 * self.cfg.probes * len(self.cfg.fish)
        self.action_space = Box(-1.0, 1.0, (self._n_act,))
        self.observation_space = Box(-np.inf, np.inf, (self._n_obs,))
        self._wp = 0
        self._body_len = [sum(length for length, _ in f.links) for f in self.cfg.fish]
        self._obs_stride = [8 + 2 * (len(f.links) - 1) for f in self.cfg.fish]
        self._t = 0
        self._last = None

    # ---- Gym API
    def reset(self, *, seed: Optional[int] = None, options=None):
        if seed is not None:
            self.action_space.seed(seed)
        self.sim.reset(0 if seed is None else int(seed))
        self._t = 0
        self._wp = 0
        obs = self._observe()
        self._last = obs.copy()
        return obs, {"n_markers": self.sim.stats().n_markers}

    def step(self, action):
        a = np.clip(np.asarray(action, dtype=np.float32).reshape(self._n_act), -1.0, 1.0)
        self.sim.set_action(a)
        self.sim.step(self.cfg.n_substeps)
        obs = self._observe()
        reward, terminated = self._reward(obs, a)
        self._last = obs.copy()
        self._t += 1
        truncated = self._t >= self.cfg.max_episode_steps
        st = self.sim.stats()
        # timing of the latest env step whose device work has finished (fg_step returns before its last collide has)
        # divergence: the observation covers the bodies only, and a blown-up fluid around a pinned (or slow) body would go
        # unnoticed there — fg_check_finite reads one population per cell (~0.1 % of an env step of 20 substeps)
        bad_cells = 0
        k = self.cfg.fluid_check_every
        if k > 0 and self._t % k == 0:
            bad_cells = self.sim.check_finite()
        info = {"mlups": st.last_mlups, "step_ms": st.last_step_ms, "fluid_bad_cells": bad_cells,
                "diverged": bool(bad_cells > 0 or not np.isfinite(obs).all())}
        return obs, float(reward), bool(terminated or info["diverged"]), bool(truncated), info

    def close(self):
        self.sim.close()

    # ---- observation: body state (+ fluid velocity probes ahead of each head)
    def _observe(self):
        body = self.sim.get_obs()
        if self.cfg.probes <= 0:
            return body
        pts, off = [], 0
        for f, stride in zip(self.cfg.fish, self._obs_stride):
            o = body[off:off + stride]
            head_len = f.links[0][0]
            ang = np.linspace(-0.6, 0.6, self.cfg.probes) + o[3]
            # nose direction is -(sin h, 0, cos h); probes sit one head length ahead, fanned in the swimming plane
            pts.append(np.stack([o[0] - 1.5 * head_len * np.sin(ang), np.full(self.cfg.probes, o[1]), o[2] - 1.5 * head_len * np.cos(ang)], 1))
            off += stride
        vel = self.sim.probe(np.concatenate(pts).astype(np.float32))[:, 1:]
        return np.concatenate([body, vel.reshape(-1).astype(np.float32)])

    # ---- task
    def _reward(self, obs, a):
        nx, ny, nz = self.cfg.grid
        if self.cfg.task == "pose":
            err_now, err_before = abs(obs[3] - self.cfg.target_heading), abs(self._last[3] - self.cfg.target_heading)
            return (err_before - err_now) - self.cfg.energy_weight * float(np.mean(a * a)), False
        if self.cfg.task == "path" and self.cfg.waypoints:
            wx, wz = self.cfg.waypoints[min(self._wp, len(self.cfg.waypoints) - 1)]
            d_now = float(np.hypot(obs[0] - wx, obs[2] - wz))
            d_before = float(np.hypot(self._last[0] - wx, self._last[2] - wz))
            r = (d_before - d_now) / self._body_len[0]
            done = False
            if d_now < self.cfg.waypoint_radius:
                r += 1.0
                self._wp += 1
                done = self._wp >= len(self.cfg.waypoints)
            return r - self.cfg.energy_weight * float(np.mean(a * a)), done
        reward, off, terminated = 0.0, 0, False
        for f, length, stride in zip(self.cfg.fish, self._body_len, self._obs_stride):
            o, p = obs[off:off + stride], self._last[off:off + stride]
            dx, dz = float(o[0] - p[0]), float(o[2] - p[2])
            dz = (dz + nz / 2) % nz - nz / 2              # z is periodic with or without tank walls on x, y
            # progress along the nose direction of the initial heading: nose = -(sin h, cos h)
            reward += -(np.sin(f.heading) * dx + np.cos(f.heading) * dz) / length
            if self.cfg.walls and not (0.1 * nx < o[0] < 0.9 * nx):
                terminated = True
            off += stride
        reward -= self.cfg.energy_weight * float(np.mean(a * a))
        return reward, terminated


class VectorFishEnv:
    """Several independent FishEnv on one GPU.  Every env owns its handle and CUDA stream; the envs
    are stepped from a thread pool (ctypes releases the GIL inside fg_step), so their kernels and host-side body
    integration overlap on the device.  Small tanks, which cannot fill a B200 alone, gain the most."""

    def __init__(self, configs: Sequence[EnvConfig], backend: str = "cuda"):
        from concurrent.futures import ThreadPoolExecutor
        self.envs = [FishEnv(c, backend=backend) for c in configs]
        self._pool = ThreadPoolExecutor(max_workers=len(self.envs))
        self.num_envs = len(self.envs)
        self.action_space, self.observation_space = self.envs[0].action_space, self.envs[0].observation_space

    def reset(self, *, seed: Optional[int] = None):
        res = [e.reset(seed=None if seed is None else seed + i) for i, e in enumerate(self.envs)]
        return np.stack([r[0] for r in res]), [r[1] for r in res]

    def step(self, actions):
        res = list(self._pool.map(lambda ea: ea[0].step(ea[1]), zip(self.envs, actions)))
        obs = np.stack([r[0] for r in res])
        return (obs, np.array([r[1] for r in res], np.float32), np.array([r[2] for r in res]), np.array([r[3] for r in res]),
                [r[4] for r in res])

    def close(self):
        self._pool.shutdown()
        for e in self.envs:
            e.close()


def save_snapshot(sim: Sim, path: str, vtk: bool = False):
    """Field output: rho, u (and the marker cloud) of one handle as .npz; optionally a legacy-VTK
    structured-points file that ParaView opens directly."""
    rho, u = sim.get_fields()
    data = dict(rho=rho, u=u)
    try:
        X, U, link = sim.get_markers()
        data.update(markers=X, marker_velocity=U, marker_link=link)
    except Exception:      # noqa: BLE001 — handles without markers, or culled multi-rank handles
        pass
    np.savez_compressed(path if path.endswith(".npz") else path + ".npz", **data)
    if vtk:
        nz, ny, nx = rho.shape
        with open((path[:-4] if path.endswith(".npz") else path) + ".vtk", "wb") as f:
            f.write(f"# vtk DataFile Version 3.0\nfishgym snapshot\nBINARY\nDATASET STRUCTURED_POINTS\nDIMENSIONS {nx} {ny} {nz}\n"
                    f"ORIGIN 0 0 0\nSPACING 1 1 1\nPOINT_DATA {nx * ny * nz}\nSCALARS rho float 1\nLOOKUP_TABLE default\n".encode())
            f.write(rho.astype(">f4").tobytes())
            f.write(b"\nVECTORS u float\n")
            f.write(np.ascontiguousarray(np.moveaxis(u, 0, -1)).astype(">f4").tobytes())
