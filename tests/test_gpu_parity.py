"""GPU parity: the CUDA library against the fp64 oracle through the C ABI (include/fishgym.h).

Tolerances are BASELINE.json:5's: rel-L2 of velocity / density <= 1e-5 (fp32 vs fp64), per-link hydrodynamic force
within 1e-4, marker->grid index maps bit-exact.
"""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

TOL_FIELD = 1e-5
TOL_FORCE = 1e-4


def test_backend_is_cuda(g, cuda):
    s = g.Sim(backend=cuda, nx=16, ny=16, nz=16)
    assert s.backend_name == "cuda-sm100a"
    s.step(2)
    assert s.stats().kernel_launches >= 3
    s.close()


@pytest.mark.parametrize("name", list(util.parity_cases(__import__("gym_fish_b200")).keys()))
def test_case_table(g, cuda, name):
    kw = util.parity_cases(g)[name]
    w = util.run_pair(g, "oracle", cuda, kw)
    assert w["u"] <= TOL_FIELD and w["rho"] <= TOL_FIELD, (name, w)
    assert w["f"] <= 2e-7, (name, w)


@pytest.mark.parametrize("name", ["bgk_periodic", "mrt_inlet_outlet_ywalls"])
def test_solid_obstacles(g, cuda, name):
    kw = util.parity_cases(g)[name]
    w = util.run_pair(g, "oracle", cuda, kw, solid=util.solid_block(kw))
    assert w["u"] <= TOL_FIELD and w["rho"] <= TOL_FIELD, (name, w)


@pytest.mark.parametrize("coll", ["bgk", "mrt"])
def test_taylor_green_64_1000_steps(g, cuda, coll):
    """BASELINE.json configs[0]: 64^3 Taylor-Green, N = 1000 steps (15 % amplitude left)."""
    n = 64
    kw = dict(nx=n, ny=n, nz=n, tau=0.8, collision=g.MRT if coll == "mrt" else g.BGK)
    a, b = g.Sim(backend="oracle", **kw), g.Sim(backend=cuda, **kw)
    rho, u = util.taylor_green(n, "xy")
    for s in (a, b):
        s.set_fields(rho, u)
        s.step(1000)
    ra, ua = a.get_fields(f64=True)
    rb, ub = b.get_fields(f64=True)
    assert util.rel_l2(ub, ua) <= TOL_FIELD
    assert util.rel_l2(rb, ra) <= TOL_FIELD
    # and the oracle itself decays at 2 nu k^2
    nu, k = (0.8 - 0.5) / 3, 2 * np.pi / n
    amp = np.sqrt((ua ** 2).mean()) / np.sqrt((u ** 2).mean())
    assert abs(-np.log(amp) / 1000 / (2 * nu * k * k) - 1) < 5e-3


def _ib_pair(g, cuda, kw, X, U, dV, link, origins, init_u, steps):
    out = []
    for backend in ("oracle", cuda):
        s = g.Sim(backend=backend, **kw)
        s.set_markers(X, U, dV, link)
        s.set_link_origins(origins)
        rho = np.ones(s.shape)
        u = np.zeros((3,) + s.shape)
        u[2] = init_u
        s.set_fields(rho, u)
        s.step(steps)
        out.append(s)
    return out


@pytest.mark.parametrize("flags", ["default", "fused_ib", "no_graphs", "tile_spread"])
def test_immersed_boundary_prescribed_markers(g, cuda, flags):
    P, Wl, IN, OUT = g.BC_PERIODIC, g.BC_WALL, g.BC_INLET, g.BC_OUTLET
    fl = {"default": 0, "fused_ib": g._abi.FLAG_FUSED_IB, "no_graphs": g._abi.FLAG_NO_GRAPHS, "tile_spread": g._abi.FLAG_IB_TILE_SPREAD}[flags]
    kw = dict(nx=20, ny=18, nz=24, tau=0.8, collision=g.MRT, max_markers=4000, max_links=4,
              bc=[P, P, Wl, Wl, IN, OUT], inlet_u=[0, 0, 0.05], flags=fl)
    X = np.concatenate([util.sphere_markers((10.3, 9.1, 8.2), 4.0, 200), util.sphere_markers((1.0, 16.5, 20.0), 3.0, 120)])
    U = np.zeros_like(X)
    U[200:, 0] = 0.01
    link = np.array([0] * 200 + [1] * 120, np.int32)
    a, b = _ib_pair(g, cuda, kw, X, U, np.ones(len(X), np.float32), link, [[10.3, 9.1, 8.2], [1.0, 16.5, 20.0]], 0.05, 11)
    ba, oa = a.get_index_map()
    bb, ob = b.get_index_map()
    assert np.array_equal(ba, bb) and np.array_equal(oa, ob)          # bit-exact
    assert a.stats().band_cells == b.stats().band_cells
    assert util.rel_l2(b.get_marker_forces(), a.get_marker_forces()) <= TOL_FORCE
    wa, wb = a.get_link_wrenches(), b.get_link_wrenches()
    assert np.abs(wb - wa).max() / np.abs(wa).max() <= TOL_FORCE
    ra, ua = a.get_fields(f64=True)
    rb, ub = b.get_fields(f64=True)
    assert util.rel_l2(ub, ua) <= TOL_FIELD and util.rel_l2(rb, ra) <= TOL_FIELD
    assert util.rel_l2(b.get_force_field(), a.get_force_field()) <= TOL_FORCE
    # property: the first sphere lies fully inside the domain, so what its markers exert equals what the grid receives
    # from it; with the second cloud clipped by the wall the total can only be checked against the oracle (above)
    Fm = b.get_marker_forces().astype(np.float64)
    assert np.isfinite(Fm).all()
    a.close(); b.close()


@pytest.mark.parametrize("graphs", [True, False])
def test_plane_split_matches_unsplit_and_oracle(g, cuda, graphs):
    """Far planes collide on a low-priority stream beside the IB kernels (default) vs everything after them
    (FG_FLAG_NO_SPLIT): same fluid up to the order of the spreading atomics; both match the oracle.  The sphere is
    re-sent every step and moves 3 planes per step, so old bands get cleared outside the new near range."""
    P, Wl, IN, OUT = g.BC_PERIODIC, g.BC_WALL, g.BC_INLET, g.BC_OUTLET
    fl = 0 if graphs else g._abi.FLAG_NO_GRAPHS
    kw = dict(nx=40, ny=36, nz=96, tau=0.8, collision=g.MRT, max_markers=1000, max_links=1, bc=[P, P, Wl, Wl, IN, OUT], inlet_u=[0, 0, 0.03],
              split_min_cells=1)
    sims = [g.Sim(backend="oracle", **kw), g.Sim(backend=cuda, flags=fl, **kw), g.Sim(backend=cuda, flags=fl | g._abi.FLAG_NO_SPLIT, **kw)]
    rho, u = util.smooth_fields(sims[0].shape, amp=0.01)
    for s in sims:
        s.set_fields(rho, u)
    for it in range(12):
        zc = 14.3 + 3.0 * it
        X = util.sphere_markers((20.2, 18.1, zc), 6.0, 450)
        U = np.zeros_like(X)
        U[:, 2] = 0.02
        for s in sims:
            s.set_markers(X, U, np.ones(450, np.float32))
            s.set_link_origins([[20.2, 18.1, zc]])
            s.step(1)
    for s in sims:
        s.step(5)            # markers left alone: static reuse of index map and band
    o, a, b = sims
    assert a.stats().split_substeps == 17 and b.stats().split_substeps == 0
    assert np.abs(a.get_populations() - b.get_populations()).max() < 2e-7
    assert util.rel_l2(a.get_fields(f64=True)[1], o.get_fields(f64=True)[1]) <= TOL_FIELD
    wa, wb, wo = a.get_link_wrenches(), b.get_link_wrenches(), o.get_link_wrenches()
    assert np.abs(wa - wo).max() / np.abs(wo).max() <= TOL_FORCE
    assert np.abs(wa - wb).max() / np.abs(wo).max() <= 1e-5
    for s in sims:
        s.close()


def test_plane_split_full_size_sphere_channel(g, cuda):
    """bench default workload (256x128x128 channel, 1.8k-marker sphere): 200 steps with and without the split."""
    import bench
    sims = []
    for fl in (0, g._abi.FLAG_NO_SPLIT):
        s, _ = bench.make_sim(g, cuda, "sphere_256x128x128", 0, 1, 0, flags=fl)
        s.step(200)
        sims.append(s)
    a, b = sims
    assert a.stats().split_substeps == 200 and b.stats().split_substeps == 0
    ua, ub = a.get_fields()[1], b.get_fields()[1]
    assert util.rel_l2(ua, ub) < 1e-6
    wa, wb = a.get_link_wrenches(), b.get_link_wrenches()
    assert np.abs(wa - wb).max() / np.abs(wb).max() < 1e-5
    for s in sims:
        s.close()


@pytest.mark.parametrize("name", ["bgk_periodic", "mrt_force", "bgk_ywall_moving", "mrt_xy_walls", "mrt_all_walls_lid",
                                  "mrt_inlet_outlet_ywalls", "mrt_outlet_inlet_xwalls"])
@pytest.mark.parametrize("lag", [0, 1, 3])
def test_fused_step_pairs_are_bit_identical_to_single_steps(g, cuda, name, lag):
    """StreamCollidePair on the GPU: the odd step chases the even step through the planes inside one launch (tickets +
    per-plane completion counters).  Large enough that thousands of CTAs are in flight; populations must equal the
    one-launch-per-step run bit for bit (any race would show up as a difference)."""
    fl = g._abi.FLAG_FUSED_PAIRS
    kw = dict(util.parity_cases(g)[name], nx=200, ny=48, nz=40, pair_lag=lag)
    a, b = g.Sim(backend=cuda, flags=fl, **kw), g.Sim(backend=cuda, **kw)
    rho, u = util.smooth_fields(a.shape)
    for s in (a, b):
        s.set_fields(rho, u)
    for n in (2, 1, 5, 40):
        a.step(n)
        b.step(n)
        assert np.array_equal(a.get_populations(), b.get_populations()), (name, n)
    assert a.stats().pair_substeps == 46 and b.stats().pair_substeps == 0
    a.close(); b.close()


@pytest.mark.parametrize("name", ["mrt_xy_walls", "mrt_all_walls_lid", "mrt_outlet_inlet_xwalls"])
def test_xwall_warp_uniform_variant_is_bit_identical(g, cuda, name):
    """CHECK_XWARP (default): only the warps at the ends of a row run the x-wall code.  Same arithmetic per cell
    as CHECK_XEDGE (FG_FLAG_NO_XWARP: predicated selects in every thread), on rows wide enough to have interior warps, with a moving
    x wall so that the wall term matters."""
    kw = dict(util.parity_cases(g)[name], nx=100, ny=9, nz=6)
    wu = dict(kw.get("wall_u", {}))
    wu[g._abi.XLO] = [0, 0.02, 0.01]
    kw["wall_u"] = wu
    a, b = g.Sim(backend=cuda, flags=g._abi.FLAG_NO_XWARP, **kw), g.Sim(backend=cuda, **kw)
    rho, u = util.smooth_fields(a.shape)
    for s in (a, b):
        s.set_fields(rho, u)
        s.step(7)
    assert np.array_equal(a.get_populations(), b.get_populations())
    o = g.Sim(backend="oracle", **kw)
    o.set_fields(rho, u)
    o.step(7)
    assert util.rel_l2(a.get_fields(f64=True)[1], o.get_fields(f64=True)[1]) <= TOL_FIELD


def test_step_returns_when_wrenches_are_there_and_later_calls_are_ordered(g, cuda):
    """fg_step waits for the link wrenches (event after the IB kernels), not for the collide of its last substep: a loop
    that feeds markers step by step overlaps with the fluid kernel.  Same results as FG_FLAG_SYNC_STEP; the timing of a
    call is reported once its device work has finished."""
    import bench
    sims = []
    for fl in (0, g._abi.FLAG_SYNC_STEP):
        s, markers = bench.make_sim(g, cuda, "sphere_256x128x128", 0, 1, 0, flags=fl)
        X, U, dV, link, _ = markers
        w = []
        for it in range(30):
            s.set_markers(X, U, dV, link)
            s.step(1)
            w.append(s.get_link_wrenches().copy())
        s.sync()
        st = s.stats()
        assert st.steps == 30 and 0.05 < st.last_step_ms < 5.0, st.last_step_ms
        sims.append((s, np.stack(w)))
    (a, wa), (b, wb) = sims
    assert np.abs(wa - wb).max() / np.abs(wb).max() < 1e-5
    assert util.rel_l2(a.get_fields()[1], b.get_fields()[1]) < 1e-6
    a.close(); b.close()


def test_spread_force_equals_marker_force_on_gpu(g, cuda):
    kw = dict(nx=48, ny=48, nz=48, tau=0.8, collision=g.MRT, max_markers=2000, max_links=1)
    s = g.Sim(backend=cuda, **kw)
    X = util.sphere_markers((24.2, 23.7, 24.4), 9.0, 1000)
    dV = np.full(1000, 4 * np.pi * 81 / 1000, np.float32)
    s.set_markers(X, np.zeros_like(X), dV, np.zeros(1000, np.int32))
    s.set_link_origins([[24.2, 23.7, 24.4]])
    u = np.zeros((3,) + s.shape)
    u[2] = 0.05
    s.set_fields(np.ones(s.shape), u)
    for n in (1, 1, 4):
        s.step(n)
        Fm = (s.get_marker_forces().astype(np.float64) * dV[:, None]).sum(0)
        Fg = s.get_force_field().astype(np.float64).sum(axis=(1, 2, 3))
        assert np.allclose(Fg, Fm, rtol=2e-5, atol=1e-7)           # sum of the delta weights is 1 per marker
        assert np.allclose(s.get_link_wrenches()[0, :3], -Fm, rtol=1e-5, atol=1e-7)
    s.close()


def test_swimming_fish_loop(g, cuda):
    kw = dict(nx=20, ny=18, nz=40, tau=0.8, max_markers=4000, max_links=8)
    a, b = g.Sim(backend="oracle", **kw), g.Sim(backend=cuda, **kw)
    for s in (a, b):
        s.add_fish(util.fish_desc(g))
    assert a.stats().n_markers == b.stats().n_markers
    for it in range(12):
        act = np.sin(0.4 * it + np.arange(3))
        for s in (a, b):
            s.set_action(act)
            s.step(10)
        wa, wb = a.get_link_wrenches(), b.get_link_wrenches()
        assert np.abs(wb - wa).max() / np.abs(wa).max() <= TOL_FORCE
        assert np.abs(a.get_obs() - b.get_obs()).max() <= 1e-4
    a.close(); b.close()


def test_mass_and_momentum_conservation_full_size(g, cuda):
    """Size-independent property at BASELINE.json configs[1] size (256x128x128, z = flow axis): a periodic box
    conserves mass and momentum to fp32 round-off."""
    kw = dict(nx=128, ny=128, nz=256, tau=0.6, collision=g.MRT)
    s = g.Sim(backend=cuda, **kw)
    rho, u = util.smooth_fields(s.shape, amp=0.03)
    s.set_fields(rho, u)
    r0, u0 = s.get_fields(f64=True)
    m0, p0 = r0.sum(), (r0 * u0).sum(axis=(1, 2, 3))
    s.step(101)
    r1, u1 = s.get_fields(f64=True)
    m1, p1 = r1.sum(), (r1 * u1).sum(axis=(1, 2, 3))
    assert abs(m1 - m0) / m0 < 1e-9
    assert np.abs(p1 - p0).max() / r0.size < 1e-9
    assert np.isfinite(u1).all()
    s.close()


def test_two_slabs_on_one_gpu_equal_unsplit(g, cuda):
    """z-slab halos with device peers: two handles in one process push halos into each other's lattice; the
    populations must be bit-identical with the unsplit run (stream-collide is per-cell deterministic)."""
    P = g.BC_PERIODIC
    kw = dict(nx=16, ny=12, nz=16, tau=0.7, collision=g.MRT)
    whole = g.Sim(backend=cuda, **kw)
    parts = [g.Sim(backend=cuda, n_ranks=2, rank=r, **kw) for r in range(2)]
    rho, u = util.smooth_fields(whole.shape)
    whole.set_fields(rho, u)
    for r, s in enumerate(parts):
        s.set_fields(rho[8 * r:8 * r + 8], u[:, 8 * r:8 * r + 8])
    h = [s.peer_export() for s in parts]
    parts[0].peer_connect(h[1], h[1])
    parts[1].peer_connect(h[0], h[0])
    for it in range(7):
        whole.step(1)
        for s in parts:
            s.step(1)
    f = whole.get_populations()
    fs = np.concatenate([s.get_populations() for s in parts], axis=1)
    assert np.array_equal(f, fs)
    for s in parts + [whole]:
        s.close()


@pytest.mark.parametrize("branch", ["default", "off", "forced_first"])
def test_halo_branch_on_two_peered_slabs_with_bodies(g, cuda, branch, monkeypatch):
    """Peered slabs with a body inside each: while no stencil reaches a boundary plane, boundary planes -> halo push -> signal run
    on their own high-priority stream beside the IB kernels (sim.hpp step(), the default on slabs below 8 M cells, which then
    run the two-cell kernels).  128-wide rows (NARROW two-cell kernels), y walls (thin rows on the branch's second stream),
    spheres that drift to the slab faces so that the last steps fall back to the serial chain; with and without the plane
    split.  Against the unsplit handle: populations within the order of the spreading atomics, wrenches within 1e-6 relative."""
    if branch == "off":
        monkeypatch.setenv("FG_NO_HALO_BRANCH", "1")
    elif branch == "forced_first":
        monkeypatch.setenv("FG_HALO_BRANCH", "1")
        monkeypatch.setenv("FG_HALO_FIRST", "1")
    P, Wl = g.BC_PERIODIC, g.BC_WALL
    for split_min in (1, 1 << 30):
        kw = dict(nx=128, ny=12, nz=64, tau=0.8, collision=g.MRT, max_markers=400, max_links=2, bc=[P, P, Wl, Wl, P, P], body_force=[0, 0, 2e-5])
        whole = g.Sim(backend=cuda, flags=g._abi.FLAG_NO_SPLIT, **kw)
        parts = [g.Sim(backend=cuda, n_ranks=2, rank=r, split_min_cells=split_min, **kw) for r in range(2)]
        rho, u = util.smooth_fields(whole.shape, amp=0.01)
        whole.set_fields(rho, u)
        for r, s in enumerate(parts):
            s.set_fields(rho[32 * r:32 * r + 32], u[:, 32 * r:32 * r + 32])
        h = [s.peer_export() for s in parts]
        parts[0].peer_connect(h[1], h[1])
        parts[1].peer_connect(h[0], h[0])
        for it in range(11):
            Xa = util.sphere_markers((60.2, 6.1, 8.3 + 1.85 * it), 2.5, 100)         # reaches slab 0's top plane at the end
            Xb = util.sphere_markers((70.7, 5.6, 56.1 - 2.0 * it), 2.5, 100)         # reaches slab 1's bottom plane at the end
            U = np.zeros((100, 3), np.float32)
            U[:, 2] = 0.01
            one = np.ones(100, np.float32)
            whole.set_markers(np.concatenate([Xa, Xb]), np.concatenate([U, -U]), np.ones(200, np.float32), np.array([0] * 100 + [1] * 100, np.int32))
            parts[0].set_markers(Xa, U, one, np.zeros(100, np.int32))
            parts[1].set_markers(Xb, -U, one, np.zeros(100, np.int32))
            whole.step(1)
            for s in parts:
                s.step(1)
        f = whole.get_populations()
        fs = np.concatenate([s.get_populations() for s in parts], axis=1)
        assert np.abs(f - fs).max() < 5e-7, (branch, split_min)
        w = whole.get_link_wrenches()
        ws = np.stack([parts[0].get_link_wrenches()[0], parts[1].get_link_wrenches()[0]])
        assert np.abs(w[:2] - ws).max() <= 1e-6 * np.abs(w).max() + 1e-12
        for s in parts + [whole]:
            s.close()


def test_velocity_probes_match_oracle_on_gpu(g, cuda):
    P, Wl = g.BC_PERIODIC, g.BC_WALL
    kw = dict(nx=14, ny=12, nz=10, tau=0.8, collision=g.MRT, bc=[P, P, Wl, Wl, P, P], body_force=[1e-4, 0, 1e-4])
    a, b = g.Sim(backend="oracle", **kw), g.Sim(backend=cuda, **kw)
    rho, u = util.smooth_fields(a.shape)
    rng = np.random.default_rng(5)
    X = (rng.uniform(0, 1, (40, 3)) * [14, 12, 10]).astype(np.float32)
    X[0] = [0.2, 0.1, 9.9]
    for s in (a, b):
        s.set_fields(rho, u)
    for n in (0, 1, 1, 4):
        for s in (a, b):
            s.step(n)
        assert np.abs(a.probe(X) - b.probe(X)).max() < 2e-6
    a.close(); b.close()


def test_vector_env_on_one_gpu(g, cuda):
    """Independent handles stepped from threads share the GPU (own streams) and reproduce the single-env trajectories."""
    from gym_fish_b200.env import EnvConfig, FishEnv, FishSpec, VectorFishEnv
    fish = FishSpec(links=((8, 2.5), (7, 2.5), (6, 2), (5, 1.5)), root=(12, 10, 14), joint_rate_max=0.02, joint_limit=0.6)
    cfgs = [EnvConfig(grid=(24, 20, 40), tau=t, collision=g.BGK, n_substeps=4, max_episode_steps=3, fish=(fish,), probes=3) for t in (0.8, 0.7, 0.9)]
    vec = VectorFishEnv(cfgs, backend=cuda)
    vec.reset(seed=2)
    acts = np.array([[0.3, -0.2, 0.1], [-1.0, 1.0, 0.0], [0.5, 0.5, -0.5]], np.float32)
    vo, vr, *_ = vec.step(acts)
    vo2, vr2, *_ = vec.step(acts)
    for i, c in enumerate(cfgs):
        e = FishEnv(c, backend=cuda)
        e.reset(seed=2 + i)
        e.step(acts[i])
        o, r, *_ = e.step(acts[i])
        assert np.abs(o - vo2[i]).max() < 1e-5 and abs(r - vr2[i]) < 1e-5     # float atomics order may differ
        e.close()
    vec.close()


def test_slab_that_runs_ahead_reports_fg_epeer(g, cuda, monkeypatch):
    """Ranks out of step: a slab stepped twice without its neighbour waits for halos that never come.  The neighbour-wait
    kernel gives up after FG_PEER_TIMEOUT_MS (20 s by default), raises a pinned host word, and the library reports
    FG_EPEER — from fg_step if the kernel has already given up, from fg_sync at the latest — instead of hanging."""
    monkeypatch.setenv("FG_PEER_TIMEOUT_MS", "250")
    kw = dict(nx=16, ny=12, nz=16, tau=0.7, collision=g.MRT)
    parts = [g.Sim(backend=cuda, n_ranks=2, rank=r, **kw) for r in range(2)]
    h = [s.peer_export() for s in parts]
    parts[0].peer_connect(h[1], h[1])
    parts[1].peer_connect(h[0], h[0])
    for it in range(2):
        for s in parts:
            s.step(1)
    parts[0].step(1)              # fine: rank 1's halos of the previous step are there
    parts[0].sync()
    with pytest.raises(g.FgError) as ei:
        parts[0].step(1)          # needs halos rank 1 has not pushed
        parts[0].sync()
    assert ei.value.code == g._abi.FG_EPEER, ei.value
    for s in parts:
        s.close()


@pytest.mark.parametrize("stop_at", [4, 5])
def test_checkpoint_resume_on_gpu(g, cuda, stop_at, tmp_path):
    """SURVEY.md §8 f3 on the CUDA path: fg_get_populations / fg_set_populations as a checkpoint, stopped after an even or
    an odd number of steps (natural vs swapped AA storage), resumed in a FRESH handle from a file; plus the field snapshot
    (npz + VTK).  The interface carries unshifted fp32 populations (one ulp of f = 3e-8), so the continuation agrees to
    that round-off; the round trip itself is exact."""
    from gym_fish_b200.env import save_snapshot
    P, Wl = g.BC_PERIODIC, g.BC_WALL
    kw = dict(nx=40, ny=24, nz=20, tau=0.7, collision=g.MRT, bc=[P, P, Wl, Wl, P, P], body_force=[1e-4, 0, 1e-4], max_markers=400, max_links=1)
    a = g.Sim(backend=cuda, **kw)
    rho, u = util.smooth_fields(a.shape)
    X = util.sphere_markers((20.2, 12.1, 10.3), 4.0, 200)
    a.set_fields(rho, u)
    a.set_markers(X, np.zeros_like(X), np.ones(200, np.float32))
    a.step(stop_at)
    np.save(tmp_path / "ckpt.npy", a.get_populations())
    save_snapshot(a, str(tmp_path / "snap"), vtk=True)
    a.step(6)
    b = g.Sim(backend=cuda, **kw)
    snap = np.load(tmp_path / "ckpt.npy")
    b.set_populations(snap)
    assert np.array_equal(b.get_populations(), snap)            # exact round trip through the device
    b.set_markers(X, np.zeros_like(X), np.ones(200, np.float32))
    b.step(6)
    assert np.abs(a.get_populations() - b.get_populations()).max() <= 1e-7
    z = np.load(tmp_path / "snap.npz")
    assert z["rho"].shape == a.shape and z["u"].shape == (3,) + a.shape and len(z["markers"]) == 200
    assert (tmp_path / "snap.vtk").stat().st_size > 4 * 4 * a.shape[0] * a.shape[1] * a.shape[2]
    a.close(); b.close()


def test_fluid_divergence_guard_on_gpu(g, cuda):
    """fg_check_finite: 0 bad cells on a healthy run; a NaN planted in one population of one cell is seen after the next
    collision at the latest, and an over-driven flow (tau close to 1/2, Mach ~ 1) is flagged once it has blown up —
    while the body observation alone would not notice (env.py info["diverged"])."""
    kw = dict(nx=32, ny=24, nz=16, tau=0.8, collision=g.BGK)
    s = g.Sim(backend=cuda, **kw)
    rho, u = util.smooth_fields(s.shape)
    s.set_fields(rho, u)
    s.step(5)
    assert s.check_finite() == 0
    f = s.get_populations()
    f[7, 3, 4, 5] = np.nan
    s.set_populations(f)
    s.step(1)
    n1 = s.check_finite()
    s.step(3)
    assert n1 >= 1 and s.check_finite() > n1                     # the NaN spreads with the streaming
    s.close()
    t = g.Sim(backend=cuda, nx=32, ny=24, nz=16, tau=0.5005, collision=g.BGK)
    rho, u = util.smooth_fields(t.shape, amp=0.6)
    t.set_fields(rho, u)
    t.step(400)
    assert t.check_finite() > 0
    t.close()


@pytest.mark.parametrize("name", ["mrt_force", "mrt_all_walls_lid", "mrt_inlet_outlet_ywalls", "mrt_xwalls_moving"])
@pytest.mark.parametrize("vec", [4, 2, "odd2", "even2+odd2"])
def test_vectorised_even_step_is_bit_identical_on_gpu(g, cuda, name, vec):
    """FG_FLAG_EVEN_VEC4 / _VEC2: 128- / 64-bit accesses, 4 / 2 cells per thread in the even step (the A/B of
    BASELINE.json:5 (a)); same arithmetic per cell => same bits as the scalar kernel, with an immersed sphere as well."""
    kw = dict(util.parity_cases(g)[name], nx=40, ny=12, nz=10, max_markers=300, max_links=1)
    A = g._abi
    # odd2: FG_FLAG_ODD_VEC2, two cells per thread in the bulk odd step as well (between x walls: its XWALL form)
    flag = {4: A.FLAG_EVEN_VEC4, 2: A.FLAG_EVEN_VEC2, "odd2": A.FLAG_ODD_VEC2, "even2+odd2": A.FLAG_EVEN_VEC2 | A.FLAG_ODD_VEC2}[vec]
    a, b = g.Sim(backend=cuda, flags=A.FLAG_EVEN_SCALAR | A.FLAG_ODD_SCALAR, **kw), g.Sim(backend=cuda, flags=flag, **kw)
    rho, u = util.smooth_fields(a.shape)
    X = util.sphere_markers((20.2, 6.1, 5.3), 2.5, 80)
    for s in (a, b):
        s.set_fields(rho, u)
        s.step(7)
    assert np.array_equal(a.get_populations(), b.get_populations())
    for s in (a, b):
        s.set_markers(X, np.zeros_like(X), np.ones(80, np.float32))
        s.step(6)
    assert np.abs(a.get_populations() - b.get_populations()).max() < 2e-7      # spreading atomics are unordered
    a.close(); b.close()


def test_default_two_cell_kernels_equal_the_scalar_ones_on_256_wide_rows(g, cuda):
    """nx % 256 == 0: the library picks the two-cell even AND odd kernels by itself (sim.hpp even_vec_width / odd_vec2); the run
    with FG_FLAG_EVEN_SCALAR | FG_FLAG_ODD_SCALAR must leave the same bits — periodic box (every row bulk) and a channel with
    y walls + inlet / outlet (bulk rows between checked wall rows)."""
    A = g._abi
    for name in ("mrt_force", "mrt_inlet_outlet_ywalls", "bgk_periodic", "mrt_xy_walls", "mrt_xwalls_moving", "mrt_all_walls_lid"):
        kw = dict(util.parity_cases(g)[name], nx=256, ny=10, nz=12)
        a, b = g.Sim(backend=cuda, **kw), g.Sim(backend=cuda, flags=A.FLAG_EVEN_SCALAR | A.FLAG_ODD_SCALAR, **kw)
        rho, u = util.smooth_fields(a.shape)
        for s in (a, b):
            s.set_fields(rho, u)
            s.step(9)
        assert np.array_equal(a.get_populations(), b.get_populations()), name
        a.close(); b.close()


@pytest.mark.parametrize("nx", [128, 64])
def test_two_cell_kernels_on_narrow_rows_take_several_rows_per_cta(g, cuda, nx):
    """nx = 128 / 64: a CTA of the two-cell kernels covers 2 / 4 consecutive rows (the NARROW instantiations; default on the
    256x128x128 channel of configs[1]); bit-identical to the scalar kernels, odd row counts, walls, and with an immersed
    sphere so that the plane split launches ranges with a hole."""
    A = g._abi
    for name in ("mrt_force", "mrt_inlet_outlet_ywalls", "mrt_xy_walls", "mrt_xwalls_moving", "bgk_periodic"):
        kw = dict(util.parity_cases(g)[name], nx=nx, ny=23, nz=20, max_markers=300, max_links=1)
        a, b = g.Sim(backend=cuda, **kw), g.Sim(backend=cuda, flags=A.FLAG_EVEN_SCALAR | A.FLAG_ODD_SCALAR, **kw)
        rho, u = util.smooth_fields(a.shape)
        for s in (a, b):
            s.set_fields(rho, u)
            s.step(9)
        assert np.array_equal(a.get_populations(), b.get_populations()), name
        X = util.sphere_markers((nx / 2 + 0.2, 11.1, 9.3), 3.0, 100)
        for s in (a, b):
            s.set_markers(X, np.zeros_like(X), np.ones(100, np.float32))
            s.step(6)
        assert np.abs(a.get_populations() - b.get_populations()).max() < 2e-7, name      # spreading atomics are unordered
        a.close(); b.close()
