"""z-slab decomposition: the N>1 path on CPU.

 * oracle, host-staged halos, N in-process slabs == unsplit run, bit-identical (fp64 messages)
 * emulated CUDA runtime, host-staged halos and device-peer pushes == unsplit run, bit-identical
 * world_size-2 `gloo` processes exchanging the halo messages with torch.distributed (the plumbing bench.py uses)
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def exchange_in_process(g, parts, periodic):
    A = g._abi
    n = len(parts)
    msgs = {}
    for r, s in enumerate(parts):
        if r < n - 1 or periodic:
            msgs[(r, "up")] = s.halo_pack(A.ZHI)
        if r > 0 or periodic:
            msgs[(r, "down")] = s.halo_pack(A.ZLO)
    for r, s in enumerate(parts):
        if r > 0 or periodic:
            s.halo_unpack(A.ZLO, msgs[((r - 1) % n, "up")])
        if r < n - 1 or periodic:
            s.halo_unpack(A.ZHI, msgs[((r + 1) % n, "down")])


def split_case(g, name):
    P, Wl, IN, OUT = g.BC_PERIODIC, g.BC_WALL, g.BC_INLET, g.BC_OUTLET
    return {
        "periodic_mrt": dict(nx=10, ny=8, nz=12, tau=0.7, collision=g.MRT, body_force=[1e-4, 0, 2e-4]),
        "inlet_outlet_walls": dict(nx=10, ny=8, nz=12, tau=0.8, collision=g.MRT, bc=[Wl, Wl, P, P, IN, OUT], inlet_u=[0, 0.01, 0.04]),
        "zwalls_bgk": dict(nx=10, ny=8, nz=12, tau=0.8, bc=[P, P, Wl, Wl, Wl, Wl], wall_u={g._abi.ZLO: [0.02, 0, 0]}),
    }[name]


@pytest.mark.parametrize("backend_name", ["oracle", "emu"])
@pytest.mark.parametrize("case", ["periodic_mrt", "inlet_outlet_walls", "zwalls_bgk"])
@pytest.mark.parametrize("n_ranks", [2, 3])
def test_host_staged_slabs_equal_unsplit(g, emu, backend_name, case, n_ranks):
    backend = emu if backend_name == "emu" else "oracle"
    kw = split_case(g, case)
    periodic = kw.get("bc", [0] * 6)[4] == g.BC_PERIODIC
    whole = g.Sim(backend=backend, **kw)
    parts = [g.Sim(backend=backend, n_ranks=n_ranks, rank=r, **kw) for r in range(n_ranks)]
    rho, u = util.smooth_fields(whole.shape)
    solid = np.zeros(whole.shape, np.uint8)
    solid[3:5, 2:4, 3:6] = 1          # straddles the slab face at z=4 (3 ranks) and sits next to z=6 (2 ranks)
    whole.set_solid(solid)
    whole.set_fields(rho, u)
    h = whole.nz // n_ranks
    for r, s in enumerate(parts):
        s.set_solid(solid)
        s.set_fields(rho[r * h:(r + 1) * h], u[:, r * h:(r + 1) * h])
    for it in range(7):
        whole.step(1)
        for s in parts:
            s.step(1)
        exchange_in_process(g, parts, periodic)
    f = whole.get_populations()
    fs = np.concatenate([s.get_populations() for s in parts], axis=1)
    keep = solid == 0
    assert np.array_equal(f * keep, fs * keep)
    with pytest.raises(g.FgError):       # stepping again without having exchanged is a state error
        parts[0].step(1)
        parts[0].step(1)


@pytest.mark.parametrize("case", ["periodic_mrt", "inlet_outlet_walls"])
@pytest.mark.parametrize("overlap", [True, False, "beside"])
def test_device_peer_pushes_equal_unsplit(g, emu, case, overlap):
    """The product's multi-GPU path (halo pushes into the neighbour's lattice + step flags), here with both slabs in
    one process on the emulated device; boundary-first ordering (overlap) must not change a bit.  "beside": the interior
    collide on the low-priority branch BESIDE the boundary chain (what slabs of >= 2^20 cells do by default) — under the
    stream-order policies of test_stream_order.py the interior then runs before or after the chain, with equal bits."""
    kw = split_case(g, case)
    kw["nz"] = 16
    periodic = kw.get("bc", [0] * 6)[4] == g.BC_PERIODIC
    whole = g.Sim(backend=emu, **kw)
    flags = 0 if overlap else g._abi.FLAG_NO_OVERLAP
    if overlap == "beside":
        kw = dict(kw, split_min_cells=1)
    parts = [g.Sim(backend=emu, n_ranks=2, rank=r, flags=flags, **kw) for r in range(2)]
    rho, u = util.smooth_fields(whole.shape)
    whole.set_fields(rho, u)
    for r, s in enumerate(parts):
        s.set_fields(rho[8 * r:8 * r + 8], u[:, 8 * r:8 * r + 8])
    h = [s.peer_export() for s in parts]
    parts[0].peer_connect(h[1] if periodic else None, h[1])
    parts[1].peer_connect(h[0], h[0] if periodic else None)
    for it in range(9):
        whole.step(1)
        for s in parts:
            s.step(1)
    assert np.array_equal(whole.get_populations(), np.concatenate([s.get_populations() for s in parts], axis=1))
    assert (parts[0].stats().split_substeps == 9) == (overlap == "beside")
    with pytest.raises(g.FgError):       # a slab that runs ahead of its neighbour is caught, not silently wrong
        parts[0].step(1)
        parts[0].step(1)


@pytest.mark.parametrize("halo_branch", [False, True])
@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("overlap", [True, False])
def test_plane_split_on_peered_slabs(g, emu, overlap, split, halo_branch, monkeypatch):
    """Plane split + z-slabs: on each rank the interior planes away from its body collide before the rank even waits
    for its neighbours' halos (sim.hpp step()).  Each slab holds a moving sphere; populations must equal the unsplit,
    un-decomposed run bit for bit (the emulation adds in a fixed order).
    Halo branch (the default on small slabs; FG_HALO_BRANCH / FG_NO_HALO_BRANCH force it on / off): while no stencil reaches a boundary plane, boundary planes -> halo push -> signal
    run on their own stream beside the IB kernels; the spheres drift towards the slab faces, so the last steps fall back to
    the serial order — with split=False the substeps are captured as graphs (under FG_EMU_GRAPHS, tests/test_stream_order.py),
    and the branch is part of the graph key (removing it from the key fails this test)."""
    monkeypatch.setenv("FG_HALO_BRANCH" if halo_branch else "FG_NO_HALO_BRANCH", "1")      # read when the handle is created
    P = g.BC_PERIODIC
    kw = dict(nx=12, ny=10, nz=64, tau=0.8, collision=g.MRT, max_markers=400, max_links=2, bc=[P] * 6, body_force=[0, 0, 2e-5])
    whole = g.Sim(backend=emu, flags=g._abi.FLAG_NO_SPLIT, **kw)
    flags = 0 if overlap else g._abi.FLAG_NO_OVERLAP
    parts = [g.Sim(backend=emu, n_ranks=2, rank=r, flags=flags, split_min_cells=1 if split else 1 << 30, **kw) for r in range(2)]
    rho, u = util.smooth_fields(whole.shape, amp=0.01)
    whole.set_fields(rho, u)
    for r, s in enumerate(parts):
        s.set_fields(rho[32 * r:32 * r + 32], u[:, 32 * r:32 * r + 32])
    h = [s.peer_export() for s in parts]
    parts[0].peer_connect(h[1], h[1])
    parts[1].peer_connect(h[0], h[0])
    for it in range(11):
        Xa = util.sphere_markers((6.2, 5.1, 8.3 + 1.85 * it), 2.5, 100)          # stays inside slab 0; its stencils reach the top plane at the end
        Xb = util.sphere_markers((5.7, 4.6, 56.1 - 2.0 * it), 2.5, 100)          # stays inside slab 1; reaches its bottom plane at the end
        U = np.zeros((100, 3), np.float32)
        U[:, 2] = 0.01
        one = np.ones(100, np.float32)
        whole.set_markers(np.concatenate([Xa, Xb]), np.concatenate([U, -U]), np.ones(200, np.float32), np.array([0] * 100 + [1] * 100, np.int32))
        parts[0].set_markers(Xa, U, one, np.zeros(100, np.int32))
        parts[1].set_markers(Xb, -U, one, np.zeros(100, np.int32))
        whole.step(1)
        for s in parts:
            s.step(1)
    if split:
        assert parts[0].stats().split_substeps >= 5 and parts[1].stats().split_substeps >= 5 and whole.stats().split_substeps == 0
    else:
        assert parts[0].stats().split_substeps == 0
    assert np.array_equal(whole.get_populations(), np.concatenate([s.get_populations() for s in parts], axis=1))


def test_markers_inside_one_slab_and_across_a_face(g, emu):
    kw = dict(nx=12, ny=12, nz=24, tau=0.8, max_markers=256, max_links=1)
    s = g.Sim(backend=emu, n_ranks=2, rank=0, **kw)
    X = util.sphere_markers((6, 6, 5.5), 2.0, 40)
    s.set_markers(X, np.zeros_like(X), np.ones(40, np.float32))          # stencils inside planes 0..11
    with pytest.raises(g.FgError) as e:
        s.set_markers(X + np.array([0, 0, 5], np.float32), np.zeros_like(X), np.ones(40, np.float32))
    assert e.value.code == g._abi.FG_ENOTSUP


WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import numpy as np, torch, torch.distributed as dist
import gym_fish_b200 as g, util
util.register_oracle(g)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
backend = {backend!r}
kw = dict(nx=10, ny=8, nz=12, tau=0.7, collision=g.MRT, body_force=[1e-4, 0, 2e-4])
s = g.Sim(backend=backend, n_ranks=world, rank=rank, **kw)
h = kw["nz"] // world
rho, u = util.smooth_fields((kw["nz"], kw["ny"], kw["nx"]))
s.set_fields(rho[rank*h:(rank+1)*h], u[:, rank*h:(rank+1)*h])
A = g._abi
up, down = (rank + 1) % world, (rank - 1) % world
for it in range(7):
    s.step(1)
    out_hi, out_lo = torch.from_numpy(s.halo_pack(A.ZHI)), torch.from_numpy(s.halo_pack(A.ZLO))
    in_lo, in_hi = torch.empty_like(out_hi), torch.empty_like(out_lo)
    reqs = [dist.isend(out_hi, up, tag=1), dist.isend(out_lo, down, tag=2), dist.irecv(in_lo, down, tag=1), dist.irecv(in_hi, up, tag=2)]
    for r in reqs: r.wait()
    s.halo_unpack(A.ZLO, in_lo.numpy()); s.halo_unpack(A.ZHI, in_hi.numpy())
f = torch.from_numpy(s.get_populations())
parts = [torch.empty_like(f) for _ in range(world)]
dist.all_gather(parts, f)
if rank == 0:
    np.save({out!r}, torch.cat(parts, dim=1).numpy())
dist.barrier(); dist.destroy_process_group()
"""


@pytest.mark.parametrize("backend_name", ["oracle", "emu"])
def test_gloo_world_size_2(g, emu, backend_name, tmp_path):
    backend = emu if backend_name == "emu" else "oracle"
    out = str(tmp_path / "f.npy")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, tests=os.path.join(ROOT, "tests"), backend=backend, out=out))
    env = dict(os.environ, OMP_NUM_THREADS="2")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                    "--master-port", str(29500 + os.getpid() % 2000), str(script)], check=True, env=env, timeout=300,
                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    kw = dict(nx=10, ny=8, nz=12, tau=0.7, collision=g.MRT, body_force=[1e-4, 0, 2e-4])
    whole = g.Sim(backend=backend, **kw)
    rho, u = util.smooth_fields(whole.shape)
    whole.set_fields(rho, u)
    whole.step(7)
    assert np.array_equal(whole.get_populations(), np.load(out))


def _run_threads(fns):
    import threading
    errs = []

    def wrap(f):
        try:
            f()
        except Exception as e:      # noqa: BLE001
            errs.append(e)
    ts = [threading.Thread(target=wrap, args=(f,)) for f in fns]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errs:
        raise errs[0]


@pytest.mark.parametrize("n_ranks", [2, 3])
def test_bodies_across_slab_faces_equal_unsplit(g, emu, n_ranks):
    """fg_peer_connect_all: a sphere straddling a slab face and one that wraps around the periodic z boundary, marker
    list replicated on every rank.  Partial U* go to the face neighbour, wrenches are all-gathered; fields, marker
    forces and wrenches must match the unsplit run (sums are associated differently: round-off only)."""
    nz = 12 * n_ranks
    kw = dict(nx=16, ny=14, nz=nz, tau=0.8, collision=g.MRT, max_markers=600, max_links=3, body_force=[0, 0, 2e-5])
    whole = g.Sim(backend=emu, **kw)
    parts = [g.Sim(backend=emu, n_ranks=n_ranks, rank=r, **kw) for r in range(n_ranks)]
    X = np.concatenate([util.sphere_markers((8.2, 7.1, 12.3), 3.0, 120),       # straddles the face at z = 12
                        util.sphere_markers((5.0, 6.0, nz - 0.4), 2.5, 90),    # wraps around z = 0 / nz
                        util.sphere_markers((10.0, 8.0, 5.5), 2.0, 60)])       # inside slab 0
    U = np.zeros_like(X)
    U[:120, 2] = 0.01
    link = np.array([0] * 120 + [1] * 90 + [2] * 60, np.int32)
    dV = np.ones(len(X), np.float32)
    origins = [[8.2, 7.1, 12.3], [5.0, 6.0, nz - 0.4], [10.0, 8.0, 5.5]]
    rho, u = util.smooth_fields(whole.shape, amp=0.01)
    whole.set_fields(rho, u)
    h = nz // n_ranks
    handles = [s.peer_export() for s in parts]
    for r, s in enumerate(parts):
        s.set_fields(rho[r * h:(r + 1) * h], u[:, r * h:(r + 1) * h])
        s.peer_connect_all(handles)
    for s in [whole] + parts:
        s.set_markers(X, U, dV, link)
        s.set_link_origins(origins)
    whole.step(9)
    _run_threads([lambda s=s: s.step(9) for s in parts])
    f = whole.get_populations()
    fs = np.concatenate([s.get_populations() for s in parts], axis=1)
    assert np.abs(f - fs).max() < 5e-7
    w = whole.get_link_wrenches()
    for s in parts:
        ws = s.get_link_wrenches()
        assert np.abs(ws - w).max() / np.abs(w).max() < 1e-5
        assert np.array_equal(ws, parts[0].get_link_wrenches())              # bit-identical on every rank
    base_w, owner_w = whole.get_index_map()
    base_p, owner_p = parts[0].get_index_map()
    assert np.array_equal(base_w, base_p)
    # per-marker U* and force (IbPushPartial / IbAddPartial): every marker is held by its owner rank (markers a rank
    # does not hold read as zero there), and face-crossing markers carry the COMPLETED sum on both ranks that hold them
    us_w, fm_w = whole.get_marker_velocities(), whole.get_marker_forces()
    us_p = [s.get_marker_velocities() for s in parts]
    fm_p = [s.get_marker_forces() for s in parts]
    us_own = np.stack([us_p[owner_p[k]][k] for k in range(len(X))])
    fm_own = np.stack([fm_p[owner_p[k]][k] for k in range(len(X))])
    assert util.rel_l2(us_own, us_w) < 1e-5
    assert util.rel_l2(fm_own, fm_w) < 1e-4
    crossing = [k for k in range(len(X)) if sum(bool(np.any(u_[k] != 0)) for u_ in us_p) >= 2]
    assert len(crossing) >= 20                                               # the test does exercise the exchange
    for k in crossing:
        held = [u_[k] for u_ in us_p if np.any(u_[k] != 0)]
        assert all(np.abs(h_ - us_w[k]).max() < 1e-6 for h_ in held)
    assert set(np.unique(owner_p)) <= set(range(n_ranks)) and len(np.unique(owner_p)) >= 2


def test_fish_swims_across_a_slab_face(g, emu):
    kw = dict(nx=20, ny=18, nz=48, tau=0.8, max_markers=4000, max_links=8)
    whole = g.Sim(backend=emu, **kw)
    parts = [g.Sim(backend=emu, n_ranks=2, rank=r, **kw) for r in range(2)]
    handles = [s.peer_export() for s in parts]
    for s in parts:
        s.peer_connect_all(handles)
    for s in [whole] + parts:
        s.add_fish(util.fish_desc(g, root=(10, 9, 17)))        # head in slab 0, tail links reach into slab 1 (face at z = 24)
    for it in range(4):
        act = np.sin(0.5 * it + np.arange(3))
        whole.set_action(act)
        whole.step(6)
        for s in parts:
            s.set_action(act)
        _run_threads([lambda s=s: s.step(6) for s in parts])
        ow = whole.get_obs()
        for s in parts:
            assert np.abs(s.get_obs() - ow).max() < 1e-4
        assert np.array_equal(parts[0].get_obs(), parts[1].get_obs())        # replicated integrators stay bit-identical
    f = whole.get_populations()
    fs = np.concatenate([s.get_populations() for s in parts], axis=1)
    assert np.abs(f - fs).max() < 1e-6
