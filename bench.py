#!/usr/bin/env python
"""bench.py — MLUPS of the coupled fluid step (BASELINE.json `metric`) on N B200s, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

A "step" is one lattice update of the whole domain (fused D3Q19 MRT stream-collide + immersed-boundary coupling +
z-face operations).  Workload at N=1 is BASELINE.json configs[1]: MRT flow past a fixed sphere in a 256x128x128
channel (z = flow axis, inlet/outlet on z, walls on y), the sphere an immersed boundary of ~1.8k markers.  At N>1
every rank owns one such 256x128x128 slab with its own sphere (weak scaling, z-slab decomposition, halos pushed
over NVLink by the library itself; torch.distributed is used only to exchange the 320-byte peer handles, for the
barriers and for the max over ranks).

Printed JSON line: see the contract in the task statement; `value` is device-timed with inputs resident in HBM,
`e2e` goes through the C ABI with host buffers (markers up, link wrenches down, every step), `roofline` is the
stream-collide kernel alone (CUDA events around each launch, on the library's stream), `cpu_baseline` is the fp64
OpenMP oracle on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BYTES_PER_CELL_UPDATE = 152.0   # 19 populations x 4 B x (read + write), SURVEY.md §8(d)

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


def sphere_markers(center, radius, n):
    i = np.arange(n)
    z = 1 - (2 * i + 1) / n
    r = np.sqrt(1 - z * z)
    ph = np.pi * (3 - np.sqrt(5)) * i
    X = np.stack([center[0] + radius * r * np.cos(ph), center[1] + radius * r * np.sin(ph), center[2] + radius * z], 1)
    return X.astype(np.float32)


def make_sim(g, backend, wl, rank, world, device, flags=0, nz_override=None, extra=None):
    """Create one slab of the workload and put it in its initial state."""
    w = WORKLOADS[wl]
    nzl = nz_override or (w["nz"] // world if w.get("strong") else w["nz"])
    kw = dict(nx=w["nx"], ny=w["ny"], nz=nzl * world, n_ranks=world, rank=rank, device=device, collision=g.MRT, flags=flags)
    kw.update(extra or {})
    markers = None
    if wl == "sphere_256x128x128":
        nu = w["U"] * w["D"] / w["Re"]
        kw.update(tau=3 * nu + 0.5, bc=[g.BC_PERIODIC, g.BC_PERIODIC, g.BC_WALL, g.BC_WALL, g.BC_INLET, g.BC_OUTLET],
                  inlet_u=[0, 0, w["U"]], max_markers=4096, max_links=2)
        R = w["D"] / 2
        n = int(round(4 * np.pi * R * R))
        zc = rank * nzl + min(64.0, nzl / 4.0) + 0.37
        X = sphere_markers((w["nx"] / 2 + 0.21, w["ny"] / 2 + 0.13, zc), R, n)
        markers = (X, np.zeros_like(X), np.full(n, 4 * np.pi * R * R / n, np.float32), np.zeros(n, np.int32),
                   np.array([[w["nx"] / 2 + 0.21, w["ny"] / 2 + 0.13, zc]]))
    elif wl in ("box_512", "box_256"):
        kw.update(tau=0.6)
    elif wl == "box_512_ib":
        kw.update(tau=0.6, max_markers=110000, max_links=64)
        R = w["D"] / 2
        n1 = int(round(4 * np.pi * R * R))
        Xs, links, origins = [], [], []
        for s_ in range(64):
            c = (64.3 + 128 * (s_ % 4), 64.1 + 128 * ((s_ // 4) % 4), rank * nzl + 64.2 + 128 * (s_ // 16))
            Xs.append(sphere_markers(c, R, n1)); links.append(np.full(n1, s_, np.int32)); origins.append(c)
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
    u = (1e-3 * rng.standard_normal((3,) + shape)).astype(np.float32)
    u[2] += w["U"]
    sim.set_fields(rho, u)
    del rho, u
    if markers is not None:
        sim.set_markers(*markers[:4])
        sim.set_link_origins(markers[4])
    elif wl == "school_1024x512x512":
        sim._school = True      # fish are added after the ranks are connected (they cross slab faces)
    elif wl == "tank_512x256x256" and rank == 0:
        d = g.FgFishDesc()
        d.n_links = 5
        for k, (length, rad) in enumerate([(28, 7), (24, 7), (22, 6), (20, 5), (18, 3.5)]):
            d.link_len[k], d.link_rad[k] = length, rad
        d.root_pos[0], d.root_pos[1], d.root_pos[2] = w["nx"] / 2, w["ny"] / 2, nzl / 3
        d.density_ratio, d.joint_gain, d.joint_limit, d.joint_rate_max, d.free_root = 1.0, 0.2, 0.5, 0.01, 1
        sim.add_fish(d)
    return sim, markers


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_oracle_mlups(g, wl, seconds=15.0, steps=None, warmup=1, nz_sample=None):
    """The fp64 OpenMP oracle on this box's host cores, bounded sample of the same workload."""
    cores = os.cpu_count() or 1
    w = WORKLOADS[wl]
    nzl = nz_sample or w["nz"]
    sim, _ = make_sim(g, "oracle", wl, 0, 1, 0, nz_override=nzl)
    cells = w["nx"] * w["ny"] * nzl
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
    return cells * n / dt / 1e6, cores, f"{n} steps of {w['nx']}x{w['ny']}x{nzl} ({wl}), fp64 two-lattice OpenMP oracle, {cores} threads", dt / n * 1e3


def run_reference(args, g, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  The reference checkout ships no code
    (README only), so this is the from-scratch fp64 oracle (BASELINE.json north_star (2)); rank 0 alone runs it."""
    if rank != 0:
        return
    wl = args.workload
    w = WORKLOADS[wl]
    # bound the sample so that steps+warmup finish within a few minutes: probe speed on a thin slab first
    probe, cores, _, ms = cpu_oracle_mlups(g, wl, steps=2, warmup=1, nz_sample=16)
    budget_s = 150.0
    per_plane_ms = ms / 16
    nz = int(min(w["nz"], max(16, budget_s * 1e3 / max(args.steps + args.warmup, 1) / per_plane_ms)))
    nz = max(16, nz - nz % 8)
    sim, _ = make_sim(g, "oracle", wl, 0, 1, 0, nz_override=nz)
    cells = w["nx"] * w["ny"] * nz
    sim.step(args.warmup)
    t0 = time.perf_counter()
    sim.step(args.steps)
    dt = time.perf_counter() - t0
    v = cells * args.steps / dt / 1e6
    sample = f"{args.steps} steps of {w['nx']}x{w['ny']}x{nz} ({wl}; the full workload has nz={w['nz']}), fp64 OpenMP oracle, {cores} threads"
    line = {
        "impl": "reference", "metric": "MLUPS (million lattice-cell updates per second), coupled D3Q19 MRT + IB step",
        "value": v, "unit": "MLUPS", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "name": wl, "sampled_grid_xyz": [w["nx"], w["ny"], nz], "ranks": 1},
        "cpu_baseline": {"value": v, "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sphere_256x128x128", choices=list(WORKLOADS))
    ap.add_argument("--no-overlap", action="store_true", help="multi-GPU: halo after the full-slab kernel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="launch kernels directly instead of per-substep CUDA graphs")
    ap.add_argument("--no-flip", action="store_true", help="sweep planes upwards in every step (no L2 reuse between steps)")
    ap.add_argument("--pairs", action="store_true", help="fused even+odd wavefront launches (opt-in experiment, measured slower)")
    ap.add_argument("--pair-lag", type=int, default=0, help="planes between the even and odd wavefront / per wavefront chunk (0: automatic)")
    ap.add_argument("--wavefront", action="store_true", help="even + odd step as a launch-level wavefront of plane chunks (opt-in experiment, FG_FLAG_WAVEFRONT)")
    ap.add_argument("--no-xwarp", action="store_true", help="x walls: predicated wall code in every thread instead of only in the row-end warps")
    ap.add_argument("--storage", default="f32", choices=["f32", "f16"],
                    help="f16: the opt-in 16-bit-storage build (fp32 arithmetic, 76 B per cell update; NOT the headline configuration)")
    ap.add_argument("--no-split", action="store_true", help="collide all planes after the IB kernels (no far-plane branch beside them)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer loop (default: --steps)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("for --gpus N > 1 launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N "
                         "--master-addr 127.0.0.1 --master-port P bench.py --gpus N ...")

    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU legs (reference arm, cpu_baseline) are meant to use every
    # host core ("cores" in the JSON line is what they really get), so the variable is set before the oracle is loaded
    if args.impl == "reference" or world == 1:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    else:
        os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
    import gym_fish_b200 as g
    # the CPU legs (cpu_baseline, --impl reference) time the oracle; the package itself does not know it
    g.register_backend("oracle", os.path.join(ROOT, "oracle", "libfishgym_oracle.so"))

    if args.impl == "reference":
        run_reference(args, g, rank, world)
        return

    dist = None
    if world > 1:
        import torch.distributed as dist   # plumbing only: handle exchange, barriers, max over ranks
        dist.init_process_group(backend="gloo", rank=rank, world_size=world)

    def barrier():
        if dist is not None:
            dist.barrier()

    wl = args.workload
    w = WORKLOADS[wl]
    flags = ((g._abi.FLAG_NO_OVERLAP if args.no_overlap else 0) | (g._abi.FLAG_NO_GRAPHS if args.no_graphs else 0) |
             (g._abi.FLAG_NO_SPLIT if args.no_split else 0) | (g._abi.FLAG_NO_SWEEP_FLIP if args.no_flip else 0) |
             (g._abi.FLAG_FUSED_PAIRS if args.pairs else 0) | (g._abi.FLAG_NO_XWARP if args.no_xwarp else 0) |
             (g._abi.FLAG_WAVEFRONT if args.wavefront else 0))
    gpu_backend = "cuda" if args.storage == "f32" else "cuda_f16"
    bytes_per_update = BYTES_PER_CELL_UPDATE if args.storage == "f32" else BYTES_PER_CELL_UPDATE / 2
    sim, markers = make_sim(g, gpu_backend, wl, rank, world, local, flags=flags, extra=dict(pair_lag=args.pair_lag))
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, sim.peer_export())
        per = wl in ("box_512", "box_256", "box_512_ib", "tank_512x256x256", "school_1024x512x512")
        lo = handles[(rank - 1) % world] if (rank > 0 or per) else None
        hi = handles[(rank + 1) % world] if (rank < world - 1 or per) else None
        if wl == "school_1024x512x512":
            sim.peer_connect_all(handles)
        else:
            sim.peer_connect(lo, hi)
    if wl == "school_1024x512x512":
        # 4 x 2 x 2 school; heads placed so that bodies straddle slab faces (z = 512 at N=2) and the periodic wrap
        for iz, zc in enumerate((200.0, 456.0, 712.0, 968.0)):
            for yc in (128.0, 384.0):
                for xc in (128.0, 384.0):
                    d = g.FgFishDesc()
                    d.n_links = 5
                    for k, (length, rad) in enumerate([(28, 7), (24, 7), (22, 6), (20, 5), (18, 3.5)]):
                        d.link_len[k], d.link_rad[k] = length, rad
                    d.root_pos[0], d.root_pos[1], d.root_pos[2] = xc, yc, zc
                    d.density_ratio, d.joint_gain, d.joint_limit, d.joint_rate_max, d.free_root = 1.0, 0.2, 0.5, 0.01, 1
                    sim.add_fish(d)
    nz_local = w["nz"] // world if w.get("strong") else w["nz"]
    cells_local = w["nx"] * w["ny"] * nz_local
    cells_total = cells_local * world

    # ---- device-timed throughput: inputs resident in HBM, K steps in one call, CUDA events inside the library
    clocks = ClockSampler(local)      # runs through warm-up and the timed region (same load), 100 ms period
    sim.step(args.warmup)
    barrier()
    sim.sync()
    launches0 = sim.stats().kernel_launches
    barrier()
    sim.step(args.steps)          # events bracket exactly K steps on the library's stream
    sim.sync()                    # fg_step returns when the wrenches are there; the timing needs the last collide, too
    st = sim.stats()
    barrier()
    clk = clocks.stop()
    ms_local = st.last_step_ms
    launches = st.kernel_launches - launches0
    # second pass of the same K steps with every stream-collide / IB launch bracketed by CUDA events on the library's
    # stream (FG_FLAG_PROFILE; graphs off): the dominant kernel's own duration for the roofline
    sim.set_flags(flags | g._abi.FLAG_PROFILE)
    barrier()
    sim.step(args.steps)
    sp = sim.stats()
    collide_ms, collide_n, ib_ms, collide_cells = sp.collide_ms, sp.collide_launches, sp.ib_ms, sp.collide_cells
    profiled_ms = sp.last_step_ms
    sim.set_flags(flags)
    barrier()
    ms = ms_local
    if dist is not None:
        import torch
        t = torch.tensor([ms_local], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        tl = torch.tensor([float(launches)], dtype=torch.float64)
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
        launches = int(tl[0])
    value = cells_total * args.steps / ms / 1e3     # MLUPS, whole job

    # ---- end to end through the C ABI with host buffers: markers up, link wrenches down, every step
    e2e = None
    k2 = args.e2e_steps or args.steps
    if markers is not None:
        X, U, dV, link, _ = markers
        pin = [np.ascontiguousarray(a) for a in (X, U, dV, link)]
        for _ in range(3):
            sim.set_markers(*pin)
            sim.step(1)
            sim.get_link_wrenches()
        barrier()
        sim.sync()
        t0 = time.perf_counter()
        for _ in range(k2):
            sim.set_markers(*pin)                    # H2D of this step's inputs (host buffers)
            sim.step(1)                              # one coupled step
            wr = sim.get_link_wrenches()             # D2H read of the step's result
        sim.sync()
        dt = time.perf_counter() - t0
        if dist is not None:
            import torch
            t = torch.tensor([dt], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        e2e = {"value": cells_total * k2 / dt / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": int(sum(a.nbytes for a in pin)), "d2h_bytes_per_step": int(wr.nbytes + 4),
               "steps": k2, "call": "fg_set_markers + fg_step(1) + fg_get_link_wrenches per step, host buffers",
               "drag_Fz": float(wr[0, 2])}
    elif wl in ("tank_512x256x256", "school_1024x512x512"):
        # the Gym loop: action down, 20 substeps, observation up (env steps/s)
        nsub = 20
        has_fish = sim.action_size() > 0
        act = np.zeros(sim.action_size(), np.float32) if has_fish else None
        t0 = time.perf_counter()
        nenv = max(3, k2 // nsub)
        for i in range(nenv):
            if has_fish:
                act[:] = np.sin(0.3 * i + np.arange(act.size))
                sim.set_action(act)
            sim.step(nsub)
            if has_fish:
                sim.get_obs()
        sim.sync()
        dt = time.perf_counter() - t0
        e2e = {"value": cells_total * nenv * nsub / dt / 1e6, "unit": "MLUPS", "env_steps_per_s": nenv / dt,
               "substeps_per_env_step": nsub, "h2d_bytes_per_step": int(4 * sim.action_size()),
               "d2h_bytes_per_step": int(4 * sim.obs_size()),
               "call": "env.step: fg_set_action + fg_step(20) + fg_get_obs"}
    else:
        # no per-step host input exists for a pure periodic box: the host-facing call is fg_step(1) + a stats read
        t0 = time.perf_counter()
        for _ in range(k2):
            sim.step(1)
        sim.sync()
        dt = time.perf_counter() - t0
        e2e = {"value": cells_total * k2 / dt / 1e6, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "call": "fg_step(1) per step (no per-step host input or output exists for a pure periodic box; fg_sync at the end)"}

    if rank != 0:
        sim.close()
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    roof = None
    if collide_n > 0 and collide_ms > 0:
        # algorithmic bytes per launch = 152 B x cells the launch updates; averaged over the bulk launches of the timed
        # region (the thin checked launches for wall rows run beside them on another stream and are in neither sum)
        achieved = bytes_per_update * collide_cells / (collide_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))[wl if args.storage == "f32" else wl + "_f16"]
            traffic, traffic_src = tj["dram_bytes_per_launch"], "profiles/r1_traffic.json (ncu --set full capture of this workload)"
        except Exception:
            pass
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": bytes_per_update * collide_cells / collide_n,
                "kernel": "fg::StreamCollide<parity, MRT> (even+odd average)", "peak_source": peak_src,
                "kernel_ms_per_launch": collide_ms / collide_n, "launches_timed": int(collide_n),
                "cells_per_launch": collide_cells / collide_n,
                "bytes_per_cell_update": bytes_per_update, "ib_ms_per_step": ib_ms / args.steps,
                "timed_in": "second pass of the same K steps with event brackets (FG_FLAG_PROFILE)",
                "profiled_pass_ms_per_step": profiled_ms / args.steps}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            v, cores, sample, _ = cpu_oracle_mlups(g, wl, seconds=12.0)
            cpu = {"value": v, "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample}
        except Exception as e:   # the baseline is informative; never let it kill the GPU number
            cpu = {"value": None, "unit": "MLUPS", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    line = {
        "metric": "MLUPS (million lattice-cell updates per second), coupled D3Q19 MRT + IB step",
        "value": value, "unit": "MLUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if w.get("strong") else "weak", "vs_baseline": None,
        "dtype": "f32" if args.storage == "f32" else "f32 arithmetic on f16-stored populations (opt-in build)", "data": "synthetic",
        "config": {"workload": w["desc"], "name": wl, "population_storage": args.storage, "grid_per_gpu_xyz": [w["nx"], w["ny"], nz_local], "ranks": world,
                   "markers_per_gpu": int(st.n_markers), "decomposition": "z-slabs, halos by peer stores over NVLink" if world > 1 else "single GPU",
                   "halo_overlap": not args.no_overlap, "cuda_graphs": not args.no_graphs,
                   "plane_split_substeps": int(st.split_substeps), "fused_pair_substeps": int(st.pair_substeps),
                   "wavefront_pairs": bool(args.wavefront),
                   "l2": f"populations {19 * (4 if args.storage == 'f32' else 2) * cells_local / 1e6:.0f} MB per GPU > 126 MB L2, no flush needed"},
        "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
        "pct_of_hbm_roofline": (value / world) * bytes_per_update / 1e3 / peak * 100.0,
    }
    print(json.dumps(line), flush=True)
    sim.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
