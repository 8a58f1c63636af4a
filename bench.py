#!/usr/bin/env python
"""bench.py — MLUPS of the coupled fluid step (BASELINE.json `metric`) on N B200s, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

A "step" is one lattice update of the whole domain (fused D3Q19 MRT stream-collide + immersed-boundary coupling +
z-face operations).  The default workload is the one BASELINE.json quotes the metric on: 512^3 cells PER GPU (configs[4],
the weak-scaling sweep; a periodic box, z-slabs across the ranks, halos pushed over NVLink by the library itself —
torch.distributed is used only to exchange the 320-byte peer handles, for the barriers and for the max over ranks).

The same JSON line carries sub-records for the other headline configurations (own sims, timed the same way):
  N = 1   ib_overhead   512^3 + 1e5 immersed-boundary markers, static and re-sent every step (target: <= 15 % step time)
          env           512x256x256 tank with one 5-link fish, Gym loop of 20 substeps  -> env steps / s (configs[2])
          sphere        256x128x128 channel with a fixed IB sphere (configs[1], the round-1 default, kept for continuity)
  N > 1   overlap_off   the same slabs with the halo push after the full-slab kernel (configs[4] "overlap on vs off")
          sphere        one 256x128x128 sphere channel per GPU (latency-dominated weak scaling)
          parity_vs_1gpu  a small box split over the N GPUs is bit-identical to the same box on rank 0's GPU alone

Printed JSON line: see the contract in the task statement; `value` is device-timed with inputs resident in HBM, `e2e`
goes through the C ABI with host buffers every step, `roofline` is the stream-collide kernel alone (CUDA events around
each launch, on the library's stream), `cpu_baseline` is the fp64 OpenMP oracle on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BYTES_PER_CELL_UPDATE = 152.0   # 19 populations x 4 B x (read + write), SURVEY.md §8(d)
METRIC = "MLUPS (million lattice-cell updates per second), coupled D3Q19 MRT + IB step"
DEFAULT_WORKLOAD = "box_512"

WORKLOADS = {
    # name: (nx, ny, nz per GPU, sphere diameter, description)
    "sphere_256x128x128": dict(nx=128, ny=128, nz=256, D=24.0, U=0.05, Re=100.0,
                               desc="D3Q19 MRT flow past fixed IB sphere, 256x128x128 (z flow axis) per GPU"),
    "box_256": dict(nx=256, ny=256, nz=256, D=0.0, U=0.02, Re=0.0,
                    desc="D3Q19 MRT periodic box 256^3 per GPU (1.3 GB lattice)"),
    "box_512": dict(nx=512, ny=512, nz=512, D=0.0, U=0.02, Re=0.0,
                    desc="D3Q19 MRT periodic box 512^3 per GPU (weak-scaling sweep, BASELINE.json configs[4])"),
    "box_512_ib": dict(nx=512, ny=512, nz=512, D=22.2, U=0.02, Re=0.0,
                       desc="D3Q19 MRT periodic box 512^3 per GPU with ~1e5 IB markers on 64 spheres (IB overhead, configs[4])"),
    "school_1024x512x512": dict(nx=512, ny=512, nz=1024, D=0.0, U=0.0, Re=0.0, strong=True,
                                desc="D3Q19 MRT tank 1024x512x512 (z swim axis), school of 16 five-link fish (~4e4 markers); "
                                     "z-slabs across the ranks (strong scaling, BASELINE.json configs[3])"),
    "tank_512x256x256": dict(nx=256, ny=256, nz=512, D=0.0, U=0.0, Re=0.0,
                             desc="D3Q19 MRT tank 512x256x256 (z swim axis) with one 5-link fish, Gym substeps"),
}
# FG_BENCH_SHRINK=k divides every axis of every workload by k: DRY RUNS of this script on a machine without a GPU (with
# FG_CUDA_LIB pointing at the CPU emulation of the kernels); such a line says so ("dry_run_shrink") and is not a measurement
SHRINK = int(os.environ.get("FG_BENCH_SHRINK", "1"))
PERIODIC_Z = ("box_512", "box_256", "box_512_ib", "tank_512x256x256", "school_1024x512x512")
FISH_LINKS = [(28, 7), (24, 7), (22, 6), (20, 5), (18, 3.5)]


def sphere_markers(center, radius, n):
    i = np.arange(n)
    z = 1 - (2 * i + 1) / n
    r = np.sqrt(1 - z * z)
    ph = np.pi * (3 - np.sqrt(5)) * i
    X = np.stack([center[0] + radius * r * np.cos(ph), center[1] + radius * r * np.sin(ph), center[2] + radius * z], 1)
    return X.astype(np.float32)


def fish_desc(g, root, scale=1.0):
    """The 5-link swimmer of BASELINE.json configs[2] (~2k markers); `scale` shrinks it with a reduced tank (tests)."""
    d = g.FgFishDesc()
    d.n_links = len(FISH_LINKS)
    for k, (length, rad) in enumerate(FISH_LINKS):
        d.link_len[k], d.link_rad[k] = length * scale, rad * scale
    d.root_pos[0], d.root_pos[1], d.root_pos[2] = root
    d.density_ratio, d.joint_gain, d.joint_limit, d.joint_rate_max, d.free_root = 1.0, 0.2, 0.5, 0.01, 1
    return d


def make_sim(g, backend, wl, rank, world, device, flags=0, nz_override=None, extra=None, shrink=SHRINK):
    """Create one slab of the workload and put it in its initial state.  `shrink` divides every axis (and the bodies) —
    the parity tests run reduced copies of the same set-up through the CPU emulation."""
    w = WORKLOADS[wl]
    nx, ny = w["nx"] // shrink, w["ny"] // shrink
    nzl = nz_override or ((w["nz"] // world if w.get("strong") else w["nz"]) // shrink)
    kw = dict(nx=nx, ny=ny, nz=nzl * world, n_ranks=world, rank=rank, device=device, collision=g.MRT, flags=flags)
    kw.update(extra or {})
    markers = None
    if wl == "sphere_256x128x128":
        nu = w["U"] * w["D"] / w["Re"]
        kw.update(tau=3 * nu + 0.5, bc=[g.BC_PERIODIC, g.BC_PERIODIC, g.BC_WALL, g.BC_WALL, g.BC_INLET, g.BC_OUTLET],
                  inlet_u=[0, 0, w["U"]], max_markers=4096, max_links=2)
        R = w["D"] / 2 / shrink
        n = int(round(4 * np.pi * R * R))
        zc = rank * nzl + min(64.0 / shrink, nzl / 4.0) + 0.37
        X = sphere_markers((nx / 2 + 0.21, ny / 2 + 0.13, zc), R, n)
        markers = (X, np.zeros_like(X), np.full(n, 4 * np.pi * R * R / n, np.float32), np.zeros(n, np.int32),
                   np.array([[nx / 2 + 0.21, ny / 2 + 0.13, zc]]))
    elif wl in ("box_512", "box_256"):
        kw.update(tau=0.6)
    elif wl == "box_512_ib":
        kw.update(tau=0.6, max_markers=110000, max_links=64)
        R = w["D"] / 2 / shrink
        n1 = int(round(4 * np.pi * R * R))
        Xs, links, origins = [], [], []
        for s_ in range(64):
            c = ((64.3 + 128 * (s_ % 4)) / shrink, (64.1 + 128 * ((s_ // 4) % 4)) / shrink, rank * nzl + (64.2 + 128 * (s_ // 16)) / shrink)
            Xs.append(sphere_markers(c, R, n1)); links.append(np.full(n1, s_, np.int32)); origins.append(c)
        if os.environ.get("FG_BENCH_SORT_MARKERS"):      # A/B of the tile spread: list order follows the cells (z, y, x) within each sphere
            Xs = [x[np.lexsort((np.floor(x[:, 0]), np.floor(x[:, 1]), np.floor(x[:, 2])))] for x in Xs]
        X = np.concatenate(Xs)
        markers = (X, np.zeros_like(X), np.full(len(X), 4 * np.pi * R * R / n1, np.float32), np.concatenate(links), np.array(origins))
    elif wl == "school_1024x512x512":
        kw.update(tau=0.6, bc=[g.BC_WALL] * 4 + [g.BC_PERIODIC] * 2, max_markers=65536, max_links=80)
    else:
        kw.update(tau=0.6, bc=[g.BC_WALL] * 4 + [g.BC_PERIODIC] * 2, max_markers=8192, max_links=8)
    sim = g.Sim(backend=backend, **kw)
    # initial state: uniform flow + a deterministic perturbation so that no value is constant
    shape = sim.shape
    rng = np.random.default_rng(1234 + rank)
    rho = np.ones(shape, np.float32)
    u = rng.standard_normal((3,) + shape, dtype=np.float32)
    u *= np.float32(1e-3)
    u[2] += np.float32(w["U"])
    sim.set_fields(rho, u)
    del rho, u
    if markers is not None:
        sim.set_markers(*markers[:4])
        sim.set_link_origins(markers[4])
    elif wl == "school_1024x512x512":
        sim._school = True      # fish are added after the ranks are connected (they cross slab faces)
    elif wl == "tank_512x256x256" and rank == 0:
        sim.add_fish(fish_desc(g, (nx / 2, ny / 2, nzl / 3), 1.0 / shrink))
    return sim, markers


class ClockSampler:
    """SM clock and clock-event reasons sampled DURING the run (B200_PROFILING.md recipe) — through NVML from a thread of
    this process (every ~2 ms; ctypes releases the GIL while fg_step / fg_sync run), so that even a timed region of a few
    milliseconds gets its own samples; `nvidia-smi -lms` (the round-1 sampler) needs ~100 ms before its first line.
    window(t0, t1) marks the timed region; samples outside it (warm-up, the profiled pass: same load) are reported apart."""

    def __init__(self, device, period_s=0.002):
        self.samples = []          # (t, sm_mhz, reasons bitmask, power_w)
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self._nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = device
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[device])
                except (ValueError, IndexError):
                    idx = device
            self._h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._nv = pynvml
        except Exception:       # noqa: BLE001 — no NVML: fall back to one nvidia-smi query at stop()
            self._nv = None
            return
        self._period = period_s
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        nv, h = self._nv, self._h
        while not self._stop.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:   # noqa: BLE001
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.samples.append((time.perf_counter(), mhz, rs, pw))
            except Exception:       # noqa: BLE001
                pass
            self._stop.wait(self._period)

    def sample_now(self):
        """One synchronous sample (called by the timing code while the GPU is under load)."""
        if self._nv is None:
            return
        nv, h = self._nv, self._h
        try:
            self.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                 int(nv.nvmlDeviceGetCurrentClocksEventReasons(h)), nv.nvmlDeviceGetPowerUsage(h) / 1000.0))
        except Exception:           # noqa: BLE001
            pass

    def stop(self, t0=None, t1=None):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "source": "NVML, sampled from a thread of this process"}
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
        if self._nv is None:
            return self._smi_once(out)
        nv = self._nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap,
                 "hw_power_brake": nv.nvmlClocksEventReasonHwPowerBrakeSlowdown}
        inside = [s for s in self.samples if t0 is not None and t0 <= s[0] <= t1]
        # under load = power well above idle or inside the timed region
        use = inside if inside else [s for s in self.samples if s[3] > 300.0] or self.samples
        if use:
            out["sm_mhz"] = statistics.median(s[1] for s in use)
            out["sm_mhz_min"] = min(s[1] for s in use)
            out["power_w_max"] = max(s[3] for s in use)
            seen = 0
            for s in use:
                seen |= s[2]
            out["reasons"] = sorted(k for k, bit in names.items() if seen & bit)
        out["samples"] = len(use)
        out["samples_in_timed_region"] = len(inside)
        out["samples_total"] = len(self.samples)
        out["window"] = "timed region" if inside else "samples under load around the timed region (warm-up / profiled pass)"
        return out

    @staticmethod
    def _smi_once(out):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=20)
            c = [x.strip() for x in r.stdout.splitlines()[0].split(",")]
            out.update(sm_mhz=float(c[0]), sm_max_mhz=float(c[1]), samples=1, source="nvidia-smi, one query after the timed region",
                       reasons=[n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[2:6])
                                if v.lower().startswith("active")])
        except Exception:           # noqa: BLE001
            pass
        return out


class NvlinkCounters:
    """NVLink data bytes sent / received by this rank's GPU (NVML field values NVLINK_THROUGHPUT_DATA_TX / _RX, KiB, summed
    over the links) — read around the timed region at N > 1, so that the line carries the bytes the halo push really put on
    the wire next to the 5 populations x nx x ny x 4 B per face the design says it should."""

    def __init__(self, device):
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = device
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[device])
                except (ValueError, IndexError):
                    idx = device
            self._nv, self._h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.ok = self.read() is not None
        except Exception:           # noqa: BLE001
            self.ok = False

    def read(self):
        try:
            nv = self._nv
            v = nv.nvmlDeviceGetFieldValues(self._h, [(nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, 0xFFFFFFFF),
                                                      (nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, 0xFFFFFFFF)])
            if any(x.nvmlReturn != 0 for x in v):
                return None
            return int(v[0].value.ullVal) * 1024, int(v[1].value.ullVal) * 1024
        except Exception:           # noqa: BLE001
            return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:           # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def base_config(wl, world):
    """What names the workload — the same keys and values in the b200 and the reference arm."""
    w = WORKLOADS[wl]
    nz_local = w["nz"] // world if w.get("strong") else w["nz"]
    cells_local = w["nx"] * w["ny"] * nz_local
    return {"workload": w["desc"], "name": wl, "grid_per_gpu_xyz": [w["nx"], w["ny"], nz_local], "gpus": world,
            "l2": f"populations {19 * 4 * cells_local / 1e6:.0f} MB per GPU (fp32) vs 126 MB L2: inputs larger than L2, no flush between steps"}


def cpu_oracle_mlups(g, wl, seconds=15.0, steps=None, warmup=1, nz_sample=None):
    """The fp64 OpenMP oracle on this box's host cores, bounded sample of the same workload."""
    cores = os.cpu_count() or 1
    w = WORKLOADS[wl]
    nzl = nz_sample or w["nz"]
    sim, _ = make_sim(g, "oracle", wl, 0, 1, 0, nz_override=nzl)
    cells = sim.nx * sim.ny * sim.nz
    for _ in range(warmup):
        sim.step(1)
    t0 = time.perf_counter()
    sim.step(1)
    per = time.perf_counter() - t0
    n = steps if steps is not None else max(2, min(200, int(seconds / max(per, 1e-6))))
    t0 = time.perf_counter()
    sim.step(n)
    dt = time.perf_counter() - t0
    sim.close()
    return cells * n / dt / 1e6, cores, f"{n} steps of {sim.nx}x{sim.ny}x{sim.nz} ({wl}), fp64 two-lattice OpenMP oracle, {cores} threads", dt / n * 1e3


def oracle_sample_nz(wl, per_plane_ms, n_steps, budget_s):
    """Height of the slab of `wl` the oracle can step n_steps times within budget_s (and within ~10 GB of host memory)."""
    w = WORKLOADS[wl]
    nz = int(min(w["nz"], max(16, budget_s * 1e3 / max(n_steps, 1) / max(per_plane_ms, 1e-9))))
    nz = min(nz, max(16, (1 << 25) // (w["nx"] * w["ny"])))      # two fp64 lattices: 304 B per cell
    return max(16, nz - nz % 8)


def run_reference(args, g, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  The reference checkout ships no code
    (README only), so this is the from-scratch fp64 oracle (BASELINE.json north_star (2)); rank 0 alone runs it.
    Warm-up is at least 20 steps: the first steps of a fresh oracle run ~1.5x slower (first-touch page placement of
    its two lattices, OpenMP team start-up), which under-reported this arm in round 1."""
    if rank != 0:
        return
    wl = args.workload
    w = WORKLOADS[wl]
    warm = max(args.warmup, 20)
    # bound the sample so that steps+warmup finish within a few minutes: probe speed on a thin slab first
    _, cores, _, ms = cpu_oracle_mlups(g, wl, steps=2, warmup=1, nz_sample=16)
    nz = oracle_sample_nz(wl, ms / 16, args.steps + warm, args.ref_budget_s)
    sim, _ = make_sim(g, "oracle", wl, 0, 1, 0, nz_override=nz)
    cells = sim.nx * sim.ny * sim.nz
    sim.step(warm)
    t0 = time.perf_counter()
    sim.step(args.steps)
    dt = time.perf_counter() - t0
    v = cells * args.steps / dt / 1e6
    sample = (f"{args.steps} steps (after {warm} warm-up steps) of {sim.nx}x{sim.ny}x{nz} ({wl}; the full workload has nz={w['nz']} per GPU), "
              f"fp64 OpenMP oracle, {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC,
        "value": v, "unit": "MLUPS", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong" if w.get("strong") else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": base_config(wl, world),
        "run": {"sampled_grid_xyz": [sim.nx, sim.ny, nz], "ranks": 1, "warmup_steps_run": warm},
        "cpu_baseline": {"value": v, "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class Ctx:
    """Rank plumbing shared by every measurement of one bench.py process."""

    def __init__(self, g, args, rank, world, local, dist):
        self.g, self.args, self.rank, self.world, self.local, self.dist = g, args, rank, world, local, dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, x):
        if self.dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])

    def sum_over_ranks(self, x):
        if self.dist is None:
            return x
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t[0])


def connect(ctx, sim, wl):
    if ctx.world == 1:
        return
    handles = [None] * ctx.world
    ctx.dist.all_gather_object(handles, sim.peer_export())
    per = wl in PERIODIC_Z
    lo = handles[(ctx.rank - 1) % ctx.world] if (ctx.rank > 0 or per) else None
    hi = handles[(ctx.rank + 1) % ctx.world] if (ctx.rank < ctx.world - 1 or per) else None
    if wl == "school_1024x512x512":
        sim.peer_connect_all(handles)
    else:
        sim.peer_connect(lo, hi)


def add_school(g, sim):
    # 4 x 2 x 2 school; heads placed so that bodies straddle slab faces (z = 512 at N=2) and the periodic wrap
    for zc in (200.0, 456.0, 712.0, 968.0):
        for yc in (128.0, 384.0):
            for xc in (128.0, 384.0):
                sim.add_fish(fish_desc(g, (xc, yc, zc)))


def timed_steps(ctx, sim, steps, warmup, clocks=None, nvlink=None):
    """W untimed warm-up steps, then EXACTLY K steps in one fg_step call bracketed by barrier + sync on both sides; the
    time is the library's CUDA-event pair around the call on its own stream, max over ranks."""
    sim.step(warmup)
    ctx.barrier()
    sim.sync()
    s0 = sim.stats()
    nv0 = nvlink.read() if nvlink is not None and nvlink.ok else None
    ctx.barrier()
    t0 = time.perf_counter()
    sim.step(steps)               # events bracket exactly K steps on the library's stream
    if clocks is not None:
        clocks.sample_now()       # the GPU is busy with the K steps right now
    sim.sync()                    # fg_step returns when the wrenches are there; the timing needs the last collide, too
    t1 = time.perf_counter()
    st = sim.stats()
    ctx.barrier()
    if nv0 is not None:
        time.sleep(0.2)           # the NVML counters are refreshed periodically, not per packet
        nv1 = nvlink.read()
        if nv1 is not None:
            nvlink.delta = ((nv1[0] - nv0[0]) / steps, (nv1[1] - nv0[1]) / steps)
    ms = ctx.max_over_ranks(st.last_step_ms)
    return ms, st, s0, (t0, t1)


def e2e_loop(ctx, sim, wl, markers, k2, cells_total):
    """The same metric through the public C-ABI calls a user makes, with HOST buffers, every step: inputs up, one coupled
    step, result down — wall clock around the loop (fg_sync at the end), max over ranks."""
    g = ctx.g
    if markers is not None:
        X, U, dV, link, _ = markers
        pin = [np.ascontiguousarray(a) for a in (X, U, dV, link)]
        for _ in range(3):
            sim.set_markers(*pin)
            sim.step(1)
            sim.get_link_wrenches()
        ctx.barrier()
        sim.sync()
        t0 = time.perf_counter()
        for _ in range(k2):
            sim.set_markers(*pin)                    # H2D of this step's inputs (host buffers)
            sim.step(1)                              # one coupled step
            wr = sim.get_link_wrenches()             # D2H read of the step's result
        sim.sync()
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        return {"value": cells_total * k2 / dt / 1e6, "unit": "MLUPS", "ms_per_step": dt / k2 * 1e3,
                "h2d_bytes_per_step": int(sum(a.nbytes for a in pin)), "d2h_bytes_per_step": int(wr.nbytes + 4),
                "steps": k2, "call": "fg_set_markers + fg_step(1) + fg_get_link_wrenches per step, host buffers",
                "drag_Fz": float(wr[0, 2])}
    if wl in ("tank_512x256x256", "school_1024x512x512"):
        # the Gym loop: action down, 20 substeps, observation up (env steps/s)
        nsub = 20
        has_fish = sim.action_size() > 0
        act = np.zeros(sim.action_size(), np.float32) if has_fish else None
        nenv = max(3, k2 // nsub)
        ctx.barrier()
        sim.sync()
        t0 = time.perf_counter()
        for i in range(nenv):
            if has_fish:
                act[:] = np.sin(0.3 * i + np.arange(act.size))
                sim.set_action(act)
            sim.step(nsub)
            if has_fish:
                sim.get_obs()
        sim.sync()
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        return {"value": cells_total * nenv * nsub / dt / 1e6, "unit": "MLUPS", "ms_per_step": dt / (nenv * nsub) * 1e3,
                "env_steps_per_s": nenv / dt, "env_steps": nenv,
                "substeps_per_env_step": nsub, "h2d_bytes_per_step": int(4 * sim.action_size()),
                "d2h_bytes_per_step": int(4 * sim.obs_size()),
                "call": "env.step: fg_set_action + fg_step(20) + fg_get_obs"}
    # a pure periodic box has no body: the per-step host input is the set of probe coordinates, the result the fluid
    # state (rho, u) at those points — the velocity-probe observation of the env (fg_probe), 16 points
    rng = np.random.default_rng(99)
    P = (rng.uniform(0.05, 0.95, (16, 3)) * np.array([sim.nx, sim.ny, sim.nz])).astype(np.float32)
    P[:, 2] += ctx.rank * sim.nz
    for _ in range(3):
        sim.step(1)
        out = sim.probe(P)
    ctx.barrier()
    sim.sync()
    t0 = time.perf_counter()
    for _ in range(k2):
        sim.step(1)
        out = sim.probe(P)                           # H2D: 16 points; D2H: (rho, u) at them, after THIS step
    sim.sync()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    return {"value": cells_total * k2 / dt / 1e6, "unit": "MLUPS", "ms_per_step": dt / k2 * 1e3,
            "h2d_bytes_per_step": int(P.nbytes), "d2h_bytes_per_step": int(out.nbytes), "steps": k2,
            "call": "fg_step(1) + fg_probe(16 points) per step, host buffers (a periodic box has no body: the probe coordinates are the "
                    "step's input, the interpolated (rho, u) its result)",
            "probe_u_mean": float(np.asarray(out)[:, 1:].mean())}


def measure(ctx, wl, flags, steps, warmup, storage="f32", clocks=None, want_e2e=True, want_profile=True, keep=False, extra=None, nvlink=None):
    """One workload, timed as the contract says.  Returns a dict (and the live sim when keep=True)."""
    g, args = ctx.g, ctx.args
    w = WORKLOADS[wl]
    backend = "cuda" if storage == "f32" else "cuda_f16"
    bytes_per_update = BYTES_PER_CELL_UPDATE if storage == "f32" else BYTES_PER_CELL_UPDATE / 2
    sim, markers = make_sim(g, backend, wl, ctx.rank, ctx.world, ctx.local, flags=flags, extra=extra)
    connect(ctx, sim, wl)
    if wl == "school_1024x512x512":
        add_school(g, sim)
    cells_local = sim.nx * sim.ny * sim.nz
    cells_total = cells_local * ctx.world

    ms, st, s0, window = timed_steps(ctx, sim, steps, warmup, clocks, nvlink)
    launches = int(ctx.sum_over_ranks(st.kernel_launches - s0.kernel_launches))
    graph_launches = int(st.graph_launches - s0.graph_launches)
    split = int(st.split_substeps - s0.split_substeps)
    res = {"name": wl, "value": cells_total * steps / ms / 1e3, "unit": "MLUPS", "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
           "gpu_launches": launches, "markers_per_gpu": int(st.n_markers), "cells_total": cells_total, "window": window,
           "pct_of_hbm_roofline": None,
           "launch_mode": {"graph_launches": graph_launches, "plane_split_substeps": split,
                           "pair_substeps": int(st.pair_substeps - s0.pair_substeps),
                           "how": ("every substep replayed as one CUDA graph" if graph_launches >= steps else
                                   "kernel-by-kernel launches (substeps with a plane split are not captured)" if graph_launches == 0 else
                                   "mixed: some substeps as graphs, some launched directly")}}
    peak, peak_src = measured_peak()
    res["pct_of_hbm_roofline"] = (res["value"] / ctx.world) * bytes_per_update / 1e3 / peak * 100.0
    if want_profile:
        # second pass of the same K steps with every stream-collide / IB launch bracketed by CUDA events on the library's
        # stream (FG_FLAG_PROFILE; graphs off): the dominant kernel's own duration for the roofline
        sim.set_flags(flags | g._abi.FLAG_PROFILE)
        ctx.barrier()
        sim.step(steps)
        sp = sim.stats()
        sim.set_flags(flags)
        ctx.barrier()
        if sp.collide_launches > 0 and sp.collide_ms > 0:
            achieved = bytes_per_update * sp.collide_cells / (sp.collide_ms * 1e-3) / 1e9
            traffic, traffic_src = None, None
            for f in ("r2_traffic.json", "r1_traffic.json"):
                try:
                    tj = json.load(open(os.path.join(ROOT, "profiles", f)))[wl if storage == "f32" else wl + "_f16"]
                    traffic = tj["dram_bytes_per_launch"]
                    traffic_src = f"STATIC, not measured in this run: profiles/{f} (one `ncu --set full` capture of this workload's kernel, per launch)"
                    break
                except Exception:       # noqa: BLE001
                    continue
            res["roofline"] = {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": bytes_per_update * sp.collide_cells / sp.collide_launches,
                "kernel": "fg::StreamCollide (even + odd step average; at nx % 256 == 0 the two-cell forms StreamCollideEvenVec<MRT, 2> / StreamCollideOddVec2<MRT>)",
                "peak_source": peak_src,
                "frac_of_nominal_8000_gbs": achieved / 8000.0,
                "frac_note": ("frac > 1 is not an accounting error: `peak` is the bandwidth of a torch device-to-device copy (MEASURED_PEAKS.json), and this kernel "
                              "moves its algorithmic bytes (= its DRAM bytes, profiles/r2_traffic.json) faster than that copy does; ncu reports 6.8 TB/s DRAM "
                              "throughput for the even step (profiles/r2_summary.md)") if achieved > peak else None,
                "kernel_ms_per_launch": sp.collide_ms / sp.collide_launches, "launches_timed": int(sp.collide_launches),
                "cells_per_launch": sp.collide_cells / sp.collide_launches, "bytes_per_cell_update": bytes_per_update,
                "ib_ms_per_step": sp.ib_ms / steps,
                "timed_in": "second pass of the same K steps with event brackets (FG_FLAG_PROFILE)",
                "profiled_pass_ms_per_step": sp.last_step_ms / steps}
            k_ms = sp.collide_ms / sp.collide_launches
            if sp.collide_launches == steps and k_ms > res["ms_per_step"]:
                # one bulk launch per step and its bracket is longer than the whole timed step: the bracket of a DIRECTLY launched kernel
                # (the profiled pass runs without graphs) contains that launch's front-end gap, which the graph replay of the timed pass
                # does not have — the kernel's own rate is at least the whole-step rate
                res["roofline"]["timing_note"] = (
                    f"the per-launch bracket ({k_ms:.4f} ms) exceeds the whole timed step ({res['ms_per_step']:.4f} ms) by {(k_ms - res['ms_per_step']) * 1e3:.1f} us: "
                    "event brackets around directly launched kernels (no graphs in the profiled pass) include the launch gap; `achieved` is therefore "
                    f"a lower bound, and the whole step of the timed pass already moves {res['value'] / ctx.world * bytes_per_update / 1e3:.0f} GB/s of algorithmic bytes per GPU")
    if want_e2e:
        res["e2e"] = e2e_loop(ctx, sim, wl, markers, args.e2e_steps or steps, cells_total)
    if keep:
        return res, sim, markers
    sim.close()
    return res


def parity_vs_one_gpu(ctx):
    """N > 1: a small periodic MRT box split into N peered z-slabs (one per GPU, halos by peer stores exactly as in the timed
    run) must leave bit-identical populations to the same box stepped on rank 0's GPU alone."""
    g, world, rank = ctx.g, ctx.world, ctx.rank
    kw = dict(nx=48, ny=40, nz=16 * world, tau=0.7, collision=g.MRT, body_force=[1e-4, 0, 2e-4])
    nz, ny, nx = kw["nz"], kw["ny"], kw["nx"]
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    u = np.zeros((3, nz, ny, nx), np.float32)
    u[0] = 0.02 * np.sin(2 * np.pi * x / nx) * np.cos(2 * np.pi * y / ny)
    u[1] = -0.02 * np.cos(2 * np.pi * x / nx) * np.sin(2 * np.pi * z / nz)
    u[2] = 0.01 * np.cos(2 * np.pi * y / ny)
    rho = (1 + 0.01 * np.cos(2 * np.pi * z / nz)).astype(np.float32)
    h = 16
    s = g.Sim(backend="cuda", n_ranks=world, rank=rank, device=ctx.local, **kw)
    s.set_fields(rho[rank * h:(rank + 1) * h], u[:, rank * h:(rank + 1) * h])
    handles = [None] * world
    ctx.dist.all_gather_object(handles, s.peer_export())
    s.peer_connect(handles[(rank - 1) % world], handles[(rank + 1) % world])
    ctx.barrier()
    s.step(21)
    s.step(12)
    f = s.get_populations()
    s.close()
    import torch
    parts = [torch.empty(f.shape, dtype=torch.float32) for _ in range(world)]
    ctx.dist.all_gather(parts, torch.from_numpy(f))
    ok = None
    if rank == 0:
        whole = g.Sim(backend="cuda", device=ctx.local, **kw)
        whole.set_fields(rho, u)
        whole.step(33)
        ok = bool(np.array_equal(whole.get_populations(), torch.cat(parts, dim=1).numpy()))
        whole.close()
    return ok


def strip(res, keys=("window", "cells_total")):
    return {k: v for k, v in res.items() if k not in keys}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "cpu_leg"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS))
    ap.add_argument("--no-subrecords", action="store_true", help="only the main workload (no IB-overhead / env / sphere / overlap-off sub-records)")
    ap.add_argument("--no-overlap", action="store_true", help="multi-GPU: halo after the full-slab kernel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="launch kernels directly instead of per-substep CUDA graphs")
    ap.add_argument("--no-flip", action="store_true", help="sweep planes upwards in every step (no L2 reuse between steps)")
    ap.add_argument("--pairs", action="store_true", help="fused even+odd wavefront launches (opt-in experiment, measured slower)")
    ap.add_argument("--pair-lag", type=int, default=0, help="--pairs: planes between the even and the odd wavefront (0: automatic)")
    ap.add_argument("--no-xwarp", action="store_true", help="x walls: predicated wall code in every thread instead of only in the row-end warps")
    ap.add_argument("--storage", default="f32", choices=["f32", "f16"],
                    help="f16: the opt-in 16-bit-storage build (fp32 arithmetic, 76 B per cell update; NOT the headline configuration)")
    ap.add_argument("--even-vec", type=int, default=0, choices=[0, 1, 2, 4],
                    help="even steps with 1 / 2 / 4 cells per thread (32- / 64- / 128-bit accesses); 0: the library's default (2 where nx %% 256 == 0)")
    ap.add_argument("--odd-vec", type=int, default=0, choices=[0, 1, 2],
                    help="bulk odd steps with 1 / 2 cells per thread; 0: the library's default (2 where nx %% 256 == 0 and x is periodic)")
    ap.add_argument("--ib-tile-spread", action="store_true", help="A/B: spreading staged per CTA in shared memory (FG_FLAG_IB_TILE_SPREAD)")
    ap.add_argument("--sort-markers", action="store_true", help="A/B: box_512_ib markers ordered by cell within each sphere (neighbours on the surface are neighbours in the list)")
    ap.add_argument("--no-split", action="store_true", help="collide all planes after the IB kernels (no far-plane branch beside them)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer loop (default: --steps)")
    ap.add_argument("--sub-budget-s", type=float, default=480.0,
                    help="wall-clock budget of everything after the main measurement (sub-records, CPU baseline); when it runs out the line is printed with what is there")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="--impl reference: wall-clock budget that sizes the oracle's sample")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1 and args.impl != "reference":
        raise SystemExit("for --gpus N > 1 launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N "
                         "--master-addr 127.0.0.1 --master-port P bench.py --gpus N ...")

    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU legs (reference arm, cpu_baseline) are meant to use every
    # host core ("cores" in the JSON line is what they really get), so the variable is set before the oracle is loaded
    if args.impl in ("reference", "cpu_leg") or world == 1:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    else:
        os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
    import gym_fish_b200 as g
    # the CPU legs (cpu_baseline, --impl reference) time the oracle; the package itself does not know it
    g.register_backend("oracle", os.path.join(ROOT, "oracle", "libfishgym_oracle.so"))

    if args.impl == "reference":
        run_reference(args, g, rank, world)
        return
    if args.impl == "cpu_leg":
        # the `cpu_baseline` object of the b200 arm's line (bounded sample, ~10-30 s), run as a child of that arm
        _, _, _, ms16 = cpu_oracle_mlups(g, args.workload, steps=2, warmup=1, nz_sample=16)
        nzs = oracle_sample_nz(args.workload, ms16 / 16, 40, 20.0)
        v, cores, sample, _ = cpu_oracle_mlups(g, args.workload, seconds=12.0, warmup=20, nz_sample=nzs)
        print(json.dumps({"value": v, "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample + ", 20 warm-up steps, own process"}), flush=True)
        return

    dist = None
    if world > 1:
        import torch.distributed as dist   # plumbing only: handle exchange, barriers, max over ranks
        import datetime
        # a rank that dies inside a sub-record must not hold the others in a barrier for gloo's default 30 minutes
        dist.init_process_group(backend="gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=300))
    ctx = Ctx(g, args, rank, world, local, dist)

    wl = args.workload
    w = WORKLOADS[wl]
    A = g._abi
    flags = ((A.FLAG_NO_OVERLAP if args.no_overlap else 0) | (A.FLAG_NO_GRAPHS if args.no_graphs else 0) |
             (A.FLAG_NO_SPLIT if args.no_split else 0) | (A.FLAG_NO_SWEEP_FLIP if args.no_flip else 0) |
             (A.FLAG_FUSED_PAIRS if args.pairs else 0) | (A.FLAG_NO_XWARP if args.no_xwarp else 0) |
             {0: 0, 1: A.FLAG_EVEN_SCALAR, 2: A.FLAG_EVEN_VEC2, 4: A.FLAG_EVEN_VEC4}[args.even_vec] |
             (A.FLAG_IB_TILE_SPREAD if args.ib_tile_spread else 0) | {0: 0, 1: A.FLAG_ODD_SCALAR, 2: A.FLAG_ODD_VEC2}[args.odd_vec])
    if args.sort_markers:
        os.environ["FG_BENCH_SORT_MARKERS"] = "1"
    bytes_per_update = BYTES_PER_CELL_UPDATE if args.storage == "f32" else BYTES_PER_CELL_UPDATE / 2

    clocks = ClockSampler(local) if rank == 0 else None      # runs through warm-up, the timed region and the profiled pass
    nvlink = NvlinkCounters(local) if rank == 0 and world > 1 else None
    main_res, sim, markers = measure(ctx, wl, flags, args.steps, args.warmup, storage=args.storage, clocks=clocks, keep=True,
                                     extra=dict(pair_lag=args.pair_lag), nvlink=nvlink)
    clk = clocks.stop(*main_res["window"]) if clocks is not None else None

    sub = {}
    want_sub = not args.no_subrecords and wl == DEFAULT_WORKLOAD and args.storage == "f32"
    # The main measurement is done: from here on a watchdog guarantees the line.  Sub-records and the CPU baseline are
    # extras; if they overrun --sub-budget-s (a rank stuck behind a dead peer, a box much slower than expected), rank 0
    # prints the line with what it has and every rank leaves, instead of the driver's own limit ending the run with nothing.
    cfg = base_config(wl, world)
    line = {
        "metric": METRIC,
        "value": main_res["value"], "unit": "MLUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "strong" if w.get("strong") else "weak", "vs_baseline": None,
        "dtype": "f32" if args.storage == "f32" else "f32 arithmetic on f16-stored populations (opt-in build)", "data": "synthetic",
        "config": cfg,
        "run": {"population_storage": args.storage, "markers_per_gpu": main_res["markers_per_gpu"],
                "decomposition": "z-slabs, halos by peer stores over NVLink (CUDA IPC), no NCCL on the data path" if world > 1 else "single GPU",
                "halo_overlap": not args.no_overlap, "launch_mode": main_res["launch_mode"],
                "even_step_cells_per_thread": args.even_vec or "library default (2 where nx % 256 == 0, else 1)",
                "odd_step_cells_per_thread": args.odd_vec or "library default (2 in bulk rows where nx % 256 == 0 and x is periodic, else 1)"},
        "roofline": main_res.get("roofline"), "cpu_baseline": None, "e2e": main_res.get("e2e"), "gpu_launches": main_res["gpu_launches"], "clocks": clk,
        "pct_of_hbm_roofline": main_res["pct_of_hbm_roofline"],
        "sub_records": None,
    }
    if SHRINK != 1:
        line["dry_run_shrink"] = SHRINK
    if nvlink is not None:
        pop_bytes = 4 if args.storage == "f32" else 2
        faces = 2 if (wl in PERIODIC_Z or world > 2) else 1      # rank 0: both faces internal when z is periodic
        exp = faces * 5 * w["nx"] * w["ny"] * pop_bytes
        d = getattr(nvlink, "delta", None)
        line["nvlink"] = {"rank": 0, "tx_bytes_per_step": d[0] if d else None, "rx_bytes_per_step": d[1] if d else None,
                          "expected_halo_bytes_per_step_per_direction": exp,
                          "expected": f"{faces} internal face(s) x 5 outgoing populations x {w['nx']} x {w['ny']} cells x {pop_bytes} B, pushed by peer stores (ZFaceOp); "
                                      "flags and (with bodies) marker / wrench exchange add a few hundred bytes",
                          "source": "NVML NVLINK_THROUGHPUT_DATA_TX / _RX of rank 0's GPU, all links, read around the timed region" if d else
                                    "NVML NVLink throughput counters not available on this box"}
    emit_lock = threading.Lock()
    emitted = []

    def emit(note=None):
        with emit_lock:
            if emitted:
                return
            emitted.append(True)
            if rank == 0:
                line["sub_records"] = dict(sub) or None
                if note:
                    line["watchdog"] = note
                print(json.dumps(line), flush=True)

    def overrun():
        emit(f"sub-records / CPU baseline exceeded --sub-budget-s {args.sub_budget_s:.0f} s: the line carries what was finished by then")
        os._exit(0)

    watchdog = threading.Timer(args.sub_budget_s + (0.0 if rank == 0 else 5.0), overrun)
    watchdog.daemon = True
    watchdog.start()
    if want_sub and world > 1 and not args.no_overlap:
        # configs[4] "halo overlap on vs off": the same slabs, halo push after the full-slab kernel
        sim.set_flags(flags | A.FLAG_NO_OVERLAP)
        ms, st, s0, _ = timed_steps(ctx, sim, args.steps, max(3, args.warmup))
        sim.set_flags(flags)
        sub["overlap_off"] = {"value": main_res["cells_total"] * args.steps / ms / 1e3, "unit": "MLUPS", "ms_per_step": ms / args.steps,
                              "overlap_on_over_off": (ms / args.steps) / main_res["ms_per_step"]}
    sim.close()
    del sim

    # Sub-records never cost the main line: each one runs under `guarded`; after a failure (on any rank: a rank stuck in a barrier
    # behind a dead peer leaves it by the process group's time-out) the remaining ones are skipped and the line is printed.
    broken = []

    def guarded(name, fn):
        if broken:
            sub[name] = {"skipped": f"after the failure of {broken[0]}"}
            return
        try:
            fn()
        except Exception as e:      # noqa: BLE001
            broken.append(name)
            sub[name] = {"error": f"{type(e).__name__}: {e}"[:400]}

    def sub_ib():
        ib = measure(ctx, "box_512_ib", flags, args.steps, args.warmup)
        plain_e2e = main_res["e2e"]["ms_per_step"]
        sub["ib_overhead"] = {
            "workload": WORKLOADS["box_512_ib"]["desc"], "markers": ib["markers_per_gpu"],
            "static_markers": {"value": ib["value"], "ms_per_step": ib["ms_per_step"],
                               "overhead_pct": (ib["ms_per_step"] / main_res["ms_per_step"] - 1.0) * 100.0,
                               "note": "device-timed fg_step(K); marker set unchanged, index map and band reused"},
            "resent_every_step": {"value": ib["e2e"]["value"], "ms_per_step": ib["e2e"]["ms_per_step"],
                                  "overhead_pct": (ib["e2e"]["ms_per_step"] / plain_e2e - 1.0) * 100.0,
                                  "h2d_bytes_per_step": ib["e2e"]["h2d_bytes_per_step"],
                                  "note": "wall clock of fg_set_markers + fg_step(1) + fg_get_link_wrenches per step against the main "
                                          "workload's own host loop; band cleared and re-registered every step"},
            "target_pct": 15.0, "ib_kernels_ms_per_step": (ib.get("roofline") or {}).get("ib_ms_per_step"),
            "launch_mode": ib["launch_mode"]}

    def sub_env():
        tank = measure(ctx, "tank_512x256x256", flags, k_small, max(args.warmup, 20))
        sub["env"] = {"workload": WORKLOADS["tank_512x256x256"]["desc"], "env_steps_per_s": tank["e2e"]["env_steps_per_s"],
                      "substeps_per_env_step": 20, "value": tank["value"], "e2e_value": tank["e2e"]["value"], "markers": tank["markers_per_gpu"],
                      "roofline_frac": (tank.get("roofline") or {}).get("frac"), "launch_mode": tank["launch_mode"]}

    def sub_sphere():
        sph = measure(ctx, "sphere_256x128x128", flags, k_small, max(args.warmup, 20))
        sub["sphere_256x128x128"] = {"workload": WORKLOADS["sphere_256x128x128"]["desc"], "value": sph["value"], "ms_per_step": sph["ms_per_step"],
                                     "steps": k_small, "e2e_value": sph["e2e"]["value"], "roofline_frac": (sph.get("roofline") or {}).get("frac"),
                                     "pct_of_hbm_roofline": sph["pct_of_hbm_roofline"], "launch_mode": sph["launch_mode"],
                                     "scaling": "weak: one channel + sphere per GPU"}

    def sub_mdf():
        # FgConfig.ib_iterations = 3 (multi-direct forcing: two correction passes after the direct-forcing pass) on the IB-overhead
        # workload, static markers; runs last among the sub-records
        passes = 3
        m = measure(ctx, "box_512_ib", flags, args.steps, args.warmup, want_e2e=False, extra=dict(ib_iterations=passes))
        one = (sub.get("ib_overhead") or {}).get("static_markers") or {}
        sub["ib_multi_direct_forcing"] = {
            "workload": WORKLOADS["box_512_ib"]["desc"], "ib_iterations": passes, "markers": m["markers_per_gpu"], "value": m["value"],
            "ms_per_step": m["ms_per_step"], "overhead_pct_vs_no_markers": (m["ms_per_step"] / main_res["ms_per_step"] - 1.0) * 100.0,
            "ms_per_extra_pass": (m["ms_per_step"] - one["ms_per_step"]) / (passes - 1) if one.get("ms_per_step") else None,
            "ib_kernels_ms_per_step": (m.get("roofline") or {}).get("ib_ms_per_step"), "launch_mode": m["launch_mode"],
            "note": "device-timed fg_step(K), marker set unchanged; each extra pass = IbMdfGather + IbMdfSpread over all markers"}

    def sub_school():
        # configs[3]: the school of 16 fish in the 1024x512x512 tank, z-slabs across the ranks (strong scaling), bodies and
        # their marker / wrench exchange across slab faces; Gym loop of 20 substeps per env step
        sch = measure(ctx, "school_1024x512x512", flags, max(args.steps, 100), max(args.warmup, 20))
        sub["school_1024x512x512"] = {"workload": WORKLOADS["school_1024x512x512"]["desc"], "scaling": "strong", "value": sch["value"],
                                      "ms_per_step": sch["ms_per_step"], "e2e_value": sch["e2e"]["value"],
                                      "env_steps_per_s": sch["e2e"]["env_steps_per_s"], "substeps_per_env_step": 20,
                                      "markers_per_gpu": sch["markers_per_gpu"], "roofline_frac": (sch.get("roofline") or {}).get("frac"),
                                      "launch_mode": sch["launch_mode"]}

    def sub_parity():
        sub["parity_vs_1gpu"] = parity_vs_one_gpu(ctx)

    if want_sub:
        k_small = max(args.steps, 200)      # the small workloads step in ~0.1 ms: 20 steps would time 2 ms
        if world == 1:
            guarded("ib_overhead", sub_ib)
            guarded("env", sub_env)
        guarded("sphere_256x128x128", sub_sphere)
        if world == 1:
            guarded("ib_multi_direct_forcing", sub_mdf)
        if world > 1:
            guarded("parity_vs_1gpu", sub_parity)
            guarded("school_1024x512x512", sub_school)

    def leave():
        if dist is not None:
            try:
                if not broken:
                    dist.barrier()
                dist.destroy_process_group()
            except Exception:       # noqa: BLE001 — the line is what counts
                pass

    if rank != 0:
        watchdog.cancel()
        leave()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        # in a fresh process, like the reference arm: timed inside this one (after the GPU legs: pinned buffers, a fragmented
        # heap, helper threads) the same oracle ran 2.6x slower than in the arm on the same box (gpu pass b1: 35 against 91 MLUPS)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "cpu_leg", "--workload", wl], capture_output=True, text=True,
                               timeout=max(60.0, args.sub_budget_s))
            cpu = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
        except Exception as e:   # noqa: BLE001 — the baseline is informative; never let it kill the GPU number
            cpu = {"value": None, "unit": "MLUPS", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    line["cpu_baseline"] = cpu
    watchdog.cancel()
    emit()
    leave()


if __name__ == "__main__":
    main()
