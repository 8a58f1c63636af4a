"""The CUDA kernel bodies (gym-fish_b200/csrc/*.cuh), compiled by g++ and run in CPU loops by the test-only host
emulation (tests/emu), against the fp64 oracle.  This is how AA-pattern indexing, bounce-back, z-face plane ops,
the IB kernels and the body integrator are checked without a GPU; the same case table runs on the B200 in
test_gpu_parity.py through the real library."""
import numpy as np
import pytest

import util

TOL_FIELD = 1e-5
TOL_FORCE = 1e-4


@pytest.mark.parametrize("name", list(util.parity_cases(__import__("gym_fish_b200")).keys()))
def test_case_table(g, emu, name):
    w = util.run_pair(g, "oracle", emu, util.parity_cases(g)[name])
    assert w["u"] <= TOL_FIELD and w["rho"] <= TOL_FIELD and w["f"] <= 2e-7, (name, w)


@pytest.mark.parametrize("name", ["bgk_periodic", "mrt_inlet_outlet_ywalls", "mrt_all_walls_lid"])
def test_solid_obstacles(g, emu, name):
    kw = util.parity_cases(g)[name]
    w = util.run_pair(g, "oracle", emu, kw, solid=util.solid_block(kw))
    assert w["u"] <= TOL_FIELD and w["rho"] <= TOL_FIELD, (name, w)


def test_taylor_green_fp32_drift_stays_below_tolerance(g, emu):
    """60 steps of a 16^3 vortex (amplitude e^-1.85 left; later the field has vanished and a relative norm is
    meaningless): shifted fp32 populations stay well inside the 1e-5 budget."""
    n = 16
    kw = dict(nx=n, ny=n, nz=n, tau=0.8)
    a, b = g.Sim(backend="oracle", **kw), g.Sim(backend=emu, **kw)
    rho, u = util.taylor_green(n, "xz")
    for s in (a, b):
        s.set_fields(rho, u)
        s.step(60)
    assert util.rel_l2(b.get_fields(f64=True)[1], a.get_fields(f64=True)[1]) < 2e-6
    assert util.rel_l2(b.get_fields(f64=True)[0], a.get_fields(f64=True)[0]) < 1e-7


def test_set_get_populations_roundtrip_both_parities(g, emu):
    s = g.Sim(backend=emu, nx=6, ny=5, nz=4, tau=0.9)
    rng = np.random.default_rng(3)
    f = (g._abi.W[:, None, None, None] * (1 + 0.01 * rng.standard_normal((19,) + s.shape))).astype(np.float32)
    s.set_populations(f)
    assert np.abs(s.get_populations() - f).max() < 6e-8
    o = g.Sim(backend="oracle", nx=6, ny=5, nz=4, tau=0.9)
    o.set_populations(f)
    for n in (1, 1, 1):          # read-out after an even step gathers across cells (parity 1), after an odd step is local
        s.step(n)
        o.step(n)
        assert np.abs(s.get_populations() - o.get_populations()).max() < 1e-7
        assert s.stats().parity == s.stats().steps % 2


@pytest.mark.parametrize("fused", [False, True])
def test_immersed_boundary_prescribed_markers(g, emu, fused):
    """fused=True: the IB pipeline as one phased (cooperative) kernel; False (default): one launch per phase."""
    P, Wl, IN, OUT = g.BC_PERIODIC, g.BC_WALL, g.BC_INLET, g.BC_OUTLET
    kw = dict(nx=20, ny=18, nz=24, tau=0.8, collision=g.MRT, max_markers=4000, max_links=4,
              bc=[P, P, Wl, Wl, IN, OUT], inlet_u=[0, 0, 0.05], flags=g._abi.FLAG_FUSED_IB if fused else 0)
    # the second cloud wraps across periodic x and pokes through the y wall (nodes outside are dropped)
    X = np.concatenate([util.sphere_markers((10.3, 9.1, 8.2), 4.0, 200), util.sphere_markers((1.0, 16.5, 20.0), 3.0, 120)])
    U = np.zeros_like(X)
    U[200:, 0] = 0.01
    link = np.array([0] * 200 + [1] * 120, np.int32)
    sims = []
    for backend in ("oracle", emu):
        s = g.Sim(backend=backend, **kw)
        s.set_markers(X, U, np.ones(len(X), np.float32), link)
        s.set_link_origins([[10.3, 9.1, 8.2], [1.0, 16.5, 20.0]])
        u = np.zeros((3,) + s.shape)
        u[2] = 0.05
        s.set_fields(np.ones(s.shape), u)
        sims.append(s)
    a, b = sims
    for n in (1, 1, 5):           # both parities
        a.step(n)
        b.step(n)
        ba, oa = a.get_index_map()
        bb, ob = b.get_index_map()
        assert np.array_equal(ba, bb) and np.array_equal(oa, ob)      # bit-exact
        assert a.stats().band_cells == b.stats().band_cells
        assert util.rel_l2(b.get_marker_forces(), a.get_marker_forces()) <= TOL_FORCE
        assert util.rel_l2(b.get_marker_velocities(), a.get_marker_velocities()) <= TOL_FORCE
        wa, wb = a.get_link_wrenches(), b.get_link_wrenches()
        assert np.abs(wb - wa).max() / np.abs(wa).max() <= TOL_FORCE
        assert util.rel_l2(b.get_force_field(), a.get_force_field()) <= TOL_FORCE
        assert util.rel_l2(b.get_fields(f64=True)[1], a.get_fields(f64=True)[1]) <= TOL_FIELD


@pytest.mark.parametrize("walls", ["periodic", "channel", "tank"])
def test_plane_split_far_collide_beside_ib_kernels(g, emu, walls):
    """Planes away from the bodies collide on a parallel branch that does not wait for the IB kernels (sim.hpp step()).
    The emulation runs that branch FIRST, the default order without the split runs it LAST: identical populations
    prove that the far planes neither read what the IB kernels write nor write what they read.  The sphere moves
    3 planes per step so that the old band (cleared by IbClearBand) regularly lies outside the new near range."""
    P, Wl, IN, OUT = g.BC_PERIODIC, g.BC_WALL, g.BC_INLET, g.BC_OUTLET
    bc = {"periodic": [P] * 6, "channel": [P, P, Wl, Wl, IN, OUT], "tank": [Wl] * 6}[walls]
    kw = dict(nx=14, ny=12, nz=64, tau=0.8, collision=g.MRT, max_markers=300, max_links=1, bc=bc, inlet_u=[0, 0, 0.03],
              split_min_cells=1)
    sims = [g.Sim(backend="oracle", **kw), g.Sim(backend=emu, **kw), g.Sim(backend=emu, flags=g._abi.FLAG_NO_SPLIT, **kw)]
    rho, u = util.smooth_fields(sims[0].shape, amp=0.01)
    for s in sims:
        s.set_fields(rho, u)
    for it in range(9):
        zc = 12.3 + 3.0 * it
        X = util.sphere_markers((7.2, 6.1, zc), 3.0, 150)
        U = np.zeros_like(X)
        U[:, 2] = 0.02
        for s in sims:
            s.set_markers(X, U, np.ones(150, np.float32))
            s.set_link_origins([[7.2, 6.1, zc]])
            s.step(1)
    for s in sims:       # ... and a few steps with the markers left alone (static reuse of the index map and band)
        s.step(3)
    o, a, b = sims
    assert a.stats().split_substeps == 12 and b.stats().split_substeps == 0
    assert np.array_equal(a.get_populations(), b.get_populations())
    assert np.array_equal(a.get_link_wrenches(), b.get_link_wrenches())
    assert util.rel_l2(a.get_fields(f64=True)[1], o.get_fields(f64=True)[1]) <= TOL_FIELD
    wa, wo = a.get_link_wrenches(), o.get_link_wrenches()
    assert np.abs(wa - wo).max() / np.abs(wo).max() <= TOL_FORCE


@pytest.mark.parametrize("name", ["bgk_periodic", "mrt_force", "bgk_ywall_moving", "mrt_xy_walls", "mrt_all_walls_lid", "bgk_inlet_outlet",
                                  "mrt_inlet_outlet_ywalls", "mrt_outlet_inlet_xwalls"])
def test_fused_step_pairs_are_bit_identical_to_single_steps(g, emu, name):
    """StreamCollidePair (opt-in, FG_FLAG_FUSED_PAIRS): even step + following odd step of the same planes in one launch.  Per cell it is the same
    arithmetic in the same order, so populations must not differ by a bit from stepping one launch per step, for any
    number of substeps per call (pairs only form inside one fg_step call) and every boundary kind."""
    kw = dict(util.parity_cases(g)[name], nz=12)
    fl = g._abi.FLAG_FUSED_PAIRS
    a, b = g.Sim(backend=emu, flags=fl, **kw), g.Sim(backend=emu, **kw)
    rho, u = util.smooth_fields(a.shape)
    for s in (a, b):
        s.set_fields(rho, u)
    for n in (2, 1, 5, 4):
        a.step(n)
        b.step(n)
        assert np.array_equal(a.get_populations(), b.get_populations()), (name, n)
    assert a.stats().pair_substeps == 10 and b.stats().pair_substeps == 0     # 2 + 0 + (1 single, then 4) + 4
    o = g.Sim(backend="oracle", **kw)
    o.set_fields(rho, u)
    o.step(12)
    assert util.rel_l2(a.get_fields(f64=True)[1], o.get_fields(f64=True)[1]) <= TOL_FIELD


@pytest.mark.parametrize("name", ["mrt_xy_walls", "mrt_all_walls_lid", "mrt_outlet_inlet_xwalls"])
def test_xwall_warp_uniform_variant_is_bit_identical(g, emu, name):
    """CHECK_XWARP (default): only the warps at the ends of a row run the x-wall code.  Same arithmetic per cell
    as CHECK_XEDGE (FG_FLAG_NO_XWARP: predicated selects in every thread), on rows wide enough to have interior warps, with a moving
    x wall so that the wall term matters."""
    kw = dict(util.parity_cases(g)[name], nx=100, ny=9, nz=6)
    wu = dict(kw.get("wall_u", {}))
    wu[g._abi.XLO] = [0, 0.02, 0.01]
    kw["wall_u"] = wu
    a, b = g.Sim(backend=emu, flags=g._abi.FLAG_NO_XWARP, **kw), g.Sim(backend=emu, **kw)
    rho, u = util.smooth_fields(a.shape)
    for s in (a, b):
        s.set_fields(rho, u)
        s.step(7)
    assert np.array_equal(a.get_populations(), b.get_populations())
    o = g.Sim(backend="oracle", **kw)
    o.set_fields(rho, u)
    o.step(7)
    assert util.rel_l2(a.get_fields(f64=True)[1], o.get_fields(f64=True)[1]) <= TOL_FIELD


def test_no_markers_and_marker_removal(g, emu):
    kw = dict(nx=10, ny=10, nz=10, tau=0.8, max_markers=64, max_links=1)
    a, b = g.Sim(backend="oracle", **kw), g.Sim(backend=emu, **kw)
    X = util.sphere_markers((5, 5, 5), 2.0, 30)
    for s in (a, b):
        s.step(2)                                           # empty marker set
        s.set_markers(X, np.full_like(X, 0.01), np.ones(30, np.float32))
        s.step(3)
        s.set_markers(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0))   # remove them again: force must vanish
        s.step(2)
    assert util.rel_l2(b.get_fields(f64=True)[1], a.get_fields(f64=True)[1]) <= TOL_FIELD


@pytest.mark.parametrize("free", [0, 1])
def test_swimming_fish_matches_oracle_integrator(g, emu, free):
    """Two independently written host integrators (oracle/oracle_body.hpp, csrc/body.hpp) + coupled fluid."""
    kw = dict(nx=20, ny=18, nz=40, tau=0.8, max_markers=4000, max_links=8)
    a, b = g.Sim(backend="oracle", **kw), g.Sim(backend=emu, **kw)
    for s in (a, b):
        s.add_fish(util.fish_desc(g, free=free))
    assert a.stats().n_markers == b.stats().n_markers and a.obs_size() == b.obs_size() == 14 and a.action_size() == 3
    Xa, Ua, la = a.get_markers()
    Xb, Ub, lb = b.get_markers()
    assert np.array_equal(Xa, Xb) and np.array_equal(la, lb)
    for it in range(10):
        act = np.sin(0.4 * it + np.arange(3))
        for s in (a, b):
            s.set_action(act)
            s.step(10)
        wa, wb = a.get_link_wrenches(), b.get_link_wrenches()
        assert np.abs(wb - wa).max() / np.abs(wa).max() <= TOL_FORCE
        assert np.abs(a.get_obs() - b.get_obs()).max() <= 1e-4
    o = a.get_obs()
    assert np.isfinite(o).all() and abs(o[0] - 10) < 5      # free swimming stays bounded (added-mass stabilisation)
    if free:
        assert abs(o[4]) + abs(o[6]) > 0                    # and the body did pick up momentum from the fluid
    a.reset(0); b.reset(0)
    assert np.array_equal(a.get_obs(), b.get_obs())


def test_slab_larger_than_32bit_cell_index_is_rejected_before_allocating(g, emu):
    """Maximum sizes: cell indices are 32-bit inside the kernels; a slab beyond 2^31 cells must be refused up front."""
    with pytest.raises(g.FgError) as e:
        g.Sim(backend=emu, nx=2048, ny=2048, nz=512)
    assert e.value.code == g._abi.FG_EINVAL and "32-bit" in str(e.value)
    with pytest.raises(g.FgError):
        g.Sim(backend=emu, nx=8, ny=8, nz=1)          # a slab needs two planes


@pytest.mark.parametrize("stop_at", [4, 5])
def test_checkpoint_resume(g, emu, stop_at):
    """fg_get_populations / fg_set_populations as a checkpoint: stopping after an even or an odd
    number of steps (natural vs swapped AA storage) and resuming in a fresh handle continues the same trajectory.
    The interface carries unshifted fp32 populations, i.e. it rounds the internal shifted ones to one ulp of f
    (3e-8), so the continuation agrees to that round-off, not bit for bit."""
    P, Wl = g.BC_PERIODIC, g.BC_WALL
    kw = dict(nx=12, ny=9, nz=7, tau=0.7, collision=g.MRT, bc=[P, P, Wl, Wl, P, P], body_force=[1e-4, 0, 1e-4])
    a = g.Sim(backend=emu, **kw)
    rho, u = util.smooth_fields(a.shape)
    a.set_fields(rho, u)
    a.step(stop_at)
    snap = a.get_populations()
    a.step(6)
    b = g.Sim(backend=emu, **kw)
    b.set_populations(snap)
    assert np.array_equal(b.get_populations(), snap)            # the round trip itself is exact
    b.step(6)
    assert np.abs(a.get_populations() - b.get_populations()).max() <= 6e-8


def test_velocity_probes_match_oracle_at_both_parities(g, emu):
    """fg_probe: (rho, u) at arbitrary points with the 4-point delta — the observation side of the IB interpolation."""
    P, Wl = g.BC_PERIODIC, g.BC_WALL
    kw = dict(nx=14, ny=12, nz=10, tau=0.8, collision=g.MRT, bc=[P, P, Wl, Wl, P, P], body_force=[1e-4, 0, 1e-4])
    a, b = g.Sim(backend="oracle", **kw), g.Sim(backend=emu, **kw)
    rho, u = util.smooth_fields(a.shape)
    rng = np.random.default_rng(5)
    X = (rng.uniform(0, 1, (40, 3)) * [14, 12, 10]).astype(np.float32)
    X[0] = [0.2, 0.1, 9.9]            # wraps in x and z, clipped by the y wall
    X[1] = [7.0, 6.0, 5.0]            # exactly on a node
    for s in (a, b):
        s.set_fields(rho, u)
    for n in (0, 1, 1, 4):
        for s in (a, b):
            s.step(n)
        pa, pb = a.probe(X), b.probe(X)
        assert np.abs(pa - pb).max() < 2e-6
    # a probe on a node far from walls sees the cell values smoothed by the kernel: weights sum to one
    one = a.probe(np.array([[7.3, 6.4, 5.2]], np.float32))[0]
    assert abs(one[0] - 1.0) < 0.02


@pytest.mark.parametrize("name", ["bgk_periodic", "mrt_force", "mrt_xy_walls", "mrt_all_walls_lid", "mrt_inlet_outlet_ywalls", "mrt_xwalls_moving",
                                  "mrt_outlet_inlet_xwalls"])
@pytest.mark.parametrize("vec", [4, 2, "odd2", "even2+odd2"])
def test_vectorised_even_step_is_bit_identical(g, emu, name, vec):
    """FG_FLAG_EVEN_VEC4 / _VEC2 (lbm_core.cuh StreamCollideEvenVec): V cells per thread with 16- / 8-byte accesses in the
    even step — the same arithmetic per cell, so populations must equal the scalar kernel's bit for bit."""
    kw = dict(util.parity_cases(g)[name])
    A = g._abi
    # odd2: FG_FLAG_ODD_VEC2 (StreamCollideOddVec2), two cells per thread in the bulk odd step; between x walls its XWALL form
    flag = {4: A.FLAG_EVEN_VEC4, 2: A.FLAG_EVEN_VEC2, "odd2": A.FLAG_ODD_VEC2, "even2+odd2": A.FLAG_EVEN_VEC2 | A.FLAG_ODD_VEC2}[vec]
    a, b = g.Sim(backend=emu, flags=A.FLAG_EVEN_SCALAR | A.FLAG_ODD_SCALAR, **kw), g.Sim(backend=emu, flags=flag, **kw)
    rho, u = util.smooth_fields(a.shape)
    for s in (a, b):
        s.set_fields(rho, u)
        s.step(7)
    assert np.array_equal(a.get_populations(), b.get_populations())


@pytest.mark.parametrize("graphs", [False, True])
def test_static_body_node_cache_matches_recomputation(g, emu, graphs, monkeypatch):
    """Static bodies (marker set not re-sent): from the second such step on, stencil weights and band slots come from a
    per-node cache instead of being recomputed (ib_core.cuh node_weight / IbInterpSpread::cta).  Same values, same order of
    additions in the emulation => bit-identical fluid and wrenches; re-sending the markers invalidates the cache."""
    if graphs:
        monkeypatch.setenv("FG_EMU_GRAPHS", "1")
    kw = dict(nx=20, ny=18, nz=24, tau=0.8, collision=g.MRT, max_markers=600, max_links=2, body_force=[0, 0, 2e-5])
    a = g.Sim(backend=emu, **kw)                       # 320 markers < 8192: never cached
    monkeypatch.setenv("FG_IB_CACHE_MIN", "1")
    b = g.Sim(backend=emu, **kw)                       # cached from the third step of each static stretch
    X = np.concatenate([util.sphere_markers((10.3, 9.1, 8.2), 4.0, 200), util.sphere_markers((1.0, 16.5, 20.0), 3.0, 120)])   # wraps in x
    link = np.array([0] * 200 + [1] * 120, np.int32)
    rho, u = util.smooth_fields(a.shape)
    for s in (a, b):
        s.set_fields(rho, u)
    for shift in (0.0, 0.6):
        Xs = X.copy()
        Xs[:, 2] += shift
        for s in (a, b):
            s.set_markers(Xs, np.zeros_like(Xs), np.ones(len(Xs), np.float32), link)
            s.set_link_origins([[10.3, 9.1, 8.2 + shift], [1.0, 16.5, 20.0 + shift]])
            s.step(1)
            s.step(5)
        assert np.array_equal(a.get_populations(), b.get_populations())
        assert np.array_equal(a.get_link_wrenches(), b.get_link_wrenches())
        assert np.array_equal(a.get_marker_forces(), b.get_marker_forces())


def test_default_even_step_on_wide_rows_is_the_two_cell_kernel_and_bit_identical(g, emu):
    """nx a multiple of 256: the even step defaults to two cells per thread (sim.hpp even_vec_width); FG_FLAG_EVEN_SCALAR
    is the one-cell kernel.  Same bits, with y walls and a body force; against the oracle as well."""
    P, Wl = g.BC_PERIODIC, g.BC_WALL
    kw = dict(nx=256, ny=5, nz=4, tau=0.7, collision=g.MRT, bc=[P, P, Wl, Wl, P, P], body_force=[1e-4, 0, 2e-4])
    a, b, o = g.Sim(backend=emu, **kw), g.Sim(backend=emu, flags=g._abi.FLAG_EVEN_SCALAR, **kw), g.Sim(backend="oracle", **kw)
    rho, u = util.smooth_fields(a.shape)
    for s in (a, b, o):
        s.set_fields(rho, u)
        s.step(6)
    assert np.array_equal(a.get_populations(), b.get_populations())
    assert util.rel_l2(a.get_fields(f64=True)[1], o.get_fields(f64=True)[1]) <= TOL_FIELD


def test_default_two_cell_kernels_equal_the_scalar_ones_on_256_wide_rows(g, emu):
    """nx % 256 == 0: the library picks the two-cell even and odd kernels by itself; bit-identical to the scalar kernels
    (FG_FLAG_EVEN_SCALAR | FG_FLAG_ODD_SCALAR) and within tolerance of the oracle."""
    A = g._abi
    # the wall cases run StreamCollideOddVec2<MRT, XWALL>: wall selects in the two row-end warps, the plain pair code between
    for name in ("mrt_force", "mrt_inlet_outlet_ywalls", "mrt_xy_walls", "mrt_xwalls_moving", "mrt_all_walls_lid"):
        kw = dict(util.parity_cases(g)[name], nx=256, ny=6, nz=8)
        a, b, o = g.Sim(backend=emu, **kw), g.Sim(backend=emu, flags=A.FLAG_EVEN_SCALAR | A.FLAG_ODD_SCALAR, **kw), g.Sim(backend="oracle", **kw)
        rho, u = util.smooth_fields(a.shape)
        for s in (a, b, o):
            s.set_fields(rho, u)
            s.step(7)
        assert np.array_equal(a.get_populations(), b.get_populations()), name
        assert util.rel_l2(a.get_fields(f64=True)[1], o.get_fields(f64=True)[1]) <= TOL_FIELD


@pytest.mark.parametrize("nx", [128, 64])
def test_two_cell_kernels_on_narrow_rows_take_several_rows_per_cta(g, emu, nx):
    """nx = 128 / 64: a CTA of the two-cell kernels covers 2 / 4 consecutive rows of the launch (sim.hpp vec2_grid, the NARROW
    instantiations) — the 256x128x128 channel of BASELINE.json configs[1] runs them by default.  Row counts that do not fill
    the last CTA (ny = 7: 7 periodic rows, or 5 bulk rows between y walls), x walls, a plane hole from the split."""
    A = g._abi
    for name in ("mrt_force", "mrt_inlet_outlet_ywalls", "mrt_xy_walls", "mrt_xwalls_moving", "bgk_periodic"):
        kw = dict(util.parity_cases(g)[name], nx=nx, ny=7, nz=6)
        a, b, o = g.Sim(backend=emu, **kw), g.Sim(backend=emu, flags=A.FLAG_EVEN_SCALAR | A.FLAG_ODD_SCALAR, **kw), g.Sim(backend="oracle", **kw)
        rho, u = util.smooth_fields(a.shape)
        for s in (a, b, o):
            s.set_fields(rho, u)
            s.step(7)
        assert np.array_equal(a.get_populations(), b.get_populations()), name
        assert util.rel_l2(a.get_fields(f64=True)[1], o.get_fields(f64=True)[1]) <= TOL_FIELD
