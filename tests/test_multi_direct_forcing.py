"""Multi-direct forcing (SURVEY.md §8 a6 "optional multi-direct-forcing iterations", A7 (4) "optional n_iter > 1"):
FgConfig.ib_iterations = n runs n direct-forcing passes per substep.  With the Guo half-force the collide works with
u = u* + F/(2 rho0), so after the first pass a marker sees U*_k + E_k/2 (E_k = the spread force interpolated back)
instead of its target U_d,k; every further pass adds dF_k = 2 rho0 (U_d,k - U*_k) - E_k.

CPU: the oracle's loop against an independent numpy evaluation of the residual, the product's kernels (IbMdfGather /
IbMdfSpread, emulated) against the oracle.  GPU: the CUDA library against the oracle through the C ABI."""
import numpy as np
import pytest

import util

TOL_FIELD = 1e-5
TOL_FORCE = 1e-4


def _peskin(r):
    r = np.abs(np.asarray(r, dtype=np.float64))
    out = np.zeros_like(r)
    a, b = r < 1, (r >= 1) & (r < 2)
    out[a] = (3 - 2 * r[a] + np.sqrt(1 + 4 * r[a] - 4 * r[a] ** 2)) / 8
    out[b] = (5 - 2 * r[b] - np.sqrt(np.maximum(0, -7 + 12 * r[b] - 4 * r[b] ** 2))) / 8
    return out


def _interp_periodic(F, X):
    """sum_x F(x) delta_h(x - X_k) on a fully periodic lattice, F [3][nz][ny][nx] (numpy, written for this test)."""
    nz, ny, nx = F.shape[1:]
    out = np.zeros((len(X), 3))
    o = np.arange(4)
    for k, (x, y, z) in enumerate(np.asarray(X, dtype=np.float64)):
        i0, j0, k0 = int(np.floor(x)) - 1, int(np.floor(y)) - 1, int(np.floor(z)) - 1
        w = _peskin(z - (k0 + o))[:, None, None] * _peskin(y - (j0 + o))[None, :, None] * _peskin(x - (i0 + o))[None, None, :]
        blk = F[:, (k0 + o)[:, None, None] % nz, (j0 + o)[None, :, None] % ny, (i0 + o)[None, None, :] % nx]
        out[k] = (blk * w).sum(axis=(1, 2, 3))
    return out


def _sphere_case(R=4.0):
    n = int(4 * np.pi * R * R)
    X = util.sphere_markers((10.3, 9.6, 11.2), R, n)
    U = np.zeros((n, 3), np.float32)
    U[:, 2] = 0.004
    dV = np.full(n, 4 * np.pi * R * R / n, np.float32)
    return X, U, dV, np.zeros(n, np.int32)


def _make(g, backend, iters, **over):
    kw = dict(nx=20, ny=20, nz=24, tau=0.7, collision=g.MRT, max_markers=400, max_links=2, ib_iterations=iters)
    kw.update(over)
    s = g.Sim(backend=backend, **kw)
    rho, u = util.smooth_fields(s.shape, amp=0.03)
    s.set_fields(rho, u)
    return s


def _residual(s, X, U):
    """no-slip residual U_d - (U* + E/2) at the markers, from the read-outs of the last step"""
    E = _interp_periodic(s.get_force_field().astype(np.float64), X)
    return U - (s.get_marker_velocities().astype(np.float64) + 0.5 * E)


def test_oracle_residual_contracts_with_every_pass(g):
    X, U, dV, link = _sphere_case()
    rms, wrench = [], []
    for iters in (1, 2, 3, 5, 8):
        s = _make(g, "oracle", iters)
        s.set_markers(X, U, dV, link)
        s.step(1)
        rms.append(float(np.sqrt((_residual(s, X, U) ** 2).mean())))
        wrench.append(s.get_link_wrenches()[0].copy())
        s.close()
    assert all(b < 0.75 * a for a, b in zip(rms, rms[1:])), rms      # measured factor 0.64 per pass on this sphere
    assert rms[-1] < 0.06 * rms[0], rms
    # the force on the body grows towards its converged value: successive differences shrink
    d = [np.linalg.norm(b - a) for a, b in zip(wrench, wrench[1:])]
    assert d[-1] < d[0], d


def test_oracle_pass_identity(g):
    """After pass m the marker force is F^(m) = F^(m-1) + 2 (U_d - U*) - interp(spread(F^(m-1))): checked for m = 2
    against numpy from the one-pass run's own read-outs (same fluid state: one step from the same fields)."""
    X, U, dV, link = _sphere_case()
    a, b = _make(g, "oracle", 1), _make(g, "oracle", 2)
    for s in (a, b):
        s.set_markers(X, U, dV, link)
        s.step(1)
    F1 = a.get_marker_forces().astype(np.float64)
    E1 = _interp_periodic(a.get_force_field().astype(np.float64), X)
    expect = F1 + 2.0 * (U - a.get_marker_velocities().astype(np.float64)) - E1
    assert util.rel_l2(b.get_marker_forces(), expect) < 2e-6          # read-outs are fp32
    assert np.array_equal(a.get_marker_velocities(), b.get_marker_velocities())   # U* is the unforced velocity either way
    # 0 and 1 mean the same single pass
    c = _make(g, "oracle", 0)
    c.set_markers(X, U, dV, link)
    c.step(1)
    assert np.array_equal(c.get_marker_forces(), a.get_marker_forces())


@pytest.mark.parametrize("passes", [1, 3, 6])
def test_spread_force_equals_accumulated_marker_force(g, emu, passes):
    """SURVEY.md §4b property "sum of the spread force == sum of the marker forces": spreading is linear, so after any number of
    passes the force field is the spread of the ACCUMULATED marker forces (read-outs of both backends), and the link wrench is
    minus their sum."""
    X, U, dV, link = _sphere_case()
    for backend, tol in (("oracle", 2e-6), (emu, 2e-5)):        # the read-outs are fp32
        s = _make(g, backend, passes)
        s.set_markers(X, U, dV, link)
        s.set_link_origins([[10.3, 9.6, 11.2]])
        s.step(2)
        Fm = (s.get_marker_forces().astype(np.float64) * dV[:, None]).sum(0)
        Fg = s.get_force_field().astype(np.float64).sum(axis=(1, 2, 3))
        assert np.abs(Fg - Fm).max() <= tol * np.abs(Fm).max(), (backend, Fg, Fm)
        assert np.abs(s.get_link_wrenches()[0][:3] + Fm).max() <= tol * np.abs(Fm).max()
        s.close()


@pytest.mark.parametrize("passes", [1, 2, 5])
def test_oracle_fluid_momentum_changes_by_minus_the_wrench(g, passes):
    """Newton's third law for the coupling, to round-off and for any number of passes: per step the fluid's momentum changes by the
    body force on its cells minus the hydrodynamic force ON the bodies (the link wrenches, fp64 read-outs)."""
    gf = np.array([2e-5, -1e-5, 3e-5])
    X, U, dV, link = _sphere_case()
    s = _make(g, "oracle", passes, body_force=list(gf))
    s.set_markers(X, U, dV, link)

    def momentum():
        r, v = s.get_fields(f64=True)
        return (r * v).sum(axis=(1, 2, 3))
    p0 = momentum()
    for _ in range(4):
        s.step(1)
        p1 = momentum()
        w = s.get_link_wrenches().sum(axis=0)[:3]
        assert np.abs((p1 - p0) + w - gf * np.prod(s.shape)).max() < 1e-12
        p0 = p1
    s.close()


def _compare(a, b, steps=(1, 1, 4)):
    for n in steps:
        a.step(n)
        b.step(n)
        ba, oa = a.get_index_map()
        bb, ob = b.get_index_map()
        assert np.array_equal(ba, bb) and np.array_equal(oa, ob)
        assert a.stats().band_cells == b.stats().band_cells
        assert util.rel_l2(b.get_marker_forces(), a.get_marker_forces()) <= TOL_FORCE
        assert util.rel_l2(b.get_marker_velocities(), a.get_marker_velocities()) <= TOL_FORCE
        wa, wb = a.get_link_wrenches(), b.get_link_wrenches()
        assert np.abs(wb - wa).max() / np.abs(wa).max() <= TOL_FORCE
        assert util.rel_l2(b.get_force_field(), a.get_force_field()) <= TOL_FORCE
        assert util.rel_l2(b.get_fields(f64=True)[1], a.get_fields(f64=True)[1]) <= TOL_FIELD


def _walled_case(g):
    P, Wl, IN, OUT = g.BC_PERIODIC, g.BC_WALL, g.BC_INLET, g.BC_OUTLET
    kw = dict(nx=20, ny=18, nz=24, tau=0.8, collision=g.MRT, max_markers=4000, max_links=4, bc=[P, P, Wl, Wl, IN, OUT], inlet_u=[0, 0, 0.05])
    # the second cloud wraps across periodic x and pokes through the y wall (nodes outside are dropped)
    X = np.concatenate([util.sphere_markers((10.3, 9.1, 8.2), 4.0, 200), util.sphere_markers((1.0, 16.5, 20.0), 3.0, 120)])
    U = np.zeros_like(X)
    U[200:, 0] = 0.01
    link = np.array([0] * 200 + [1] * 120, np.int32)
    return kw, X, U, link


def _run_walled(g, test_backend, iters, flags=0):
    kw, X, U, link = _walled_case(g)
    sims = []
    for backend in ("oracle", test_backend):
        s = g.Sim(backend=backend, ib_iterations=iters, flags=flags if backend != "oracle" else 0, **kw)
        s.set_markers(X, U, np.full(len(X), 0.9, np.float32), link)
        s.set_link_origins([[10.3, 9.1, 8.2], [1.0, 16.5, 20.0]])
        u = np.zeros((3,) + s.shape)
        u[2] = 0.05
        s.set_fields(np.ones(s.shape), u)
        sims.append(s)
    _compare(*sims)
    for s in sims:
        s.close()


@pytest.mark.parametrize("iters", [2, 4])
@pytest.mark.parametrize("flags", ["default", "no_graphs", "fused_flag", "no_split"])
def test_emulated_kernels_match_oracle(g, emu, iters, flags):
    """FG_FLAG_FUSED_IB with ib_iterations > 1 falls back to one launch per phase (the cooperative kernel has no such loop)."""
    A = g._abi
    _run_walled(g, emu, iters, {"default": 0, "no_graphs": A.FLAG_NO_GRAPHS, "fused_flag": A.FLAG_FUSED_IB, "no_split": A.FLAG_NO_SPLIT}[flags])


@pytest.mark.parametrize("graphs", [False, True])
def test_emulated_static_body_cache_with_iterations(g, emu, graphs, monkeypatch):
    """The correction passes read stencil weights and band slots from the per-node cache of static bodies once the first
    pass of a step has filled it: bit-identical to recomputing them."""
    if graphs:
        monkeypatch.setenv("FG_EMU_GRAPHS", "1")
    kw = dict(nx=20, ny=18, nz=24, tau=0.8, collision=g.MRT, max_markers=600, max_links=2, body_force=[0, 0, 2e-5], ib_iterations=3)
    a = g.Sim(backend=emu, **kw)
    monkeypatch.setenv("FG_IB_CACHE_MIN", "1")
    b = g.Sim(backend=emu, **kw)
    o = g.Sim(backend="oracle", **kw)
    X = np.concatenate([util.sphere_markers((10.3, 9.1, 8.2), 4.0, 200), util.sphere_markers((1.0, 16.5, 20.0), 3.0, 120)])
    link = np.array([0] * 200 + [1] * 120, np.int32)
    rho, u = util.smooth_fields(a.shape)
    for s in (a, b, o):
        s.set_fields(rho, u)
    for shift in (0.0, 0.6):
        Xs = X.copy()
        Xs[:, 2] += shift
        for s in (a, b, o):
            s.set_markers(Xs, np.zeros_like(Xs), np.ones(len(Xs), np.float32), link)
            s.set_link_origins([[10.3, 9.1, 8.2 + shift], [1.0, 16.5, 20.0 + shift]])
            s.step(1)
            s.step(5)
        assert np.array_equal(a.get_populations(), b.get_populations())
        assert np.array_equal(a.get_link_wrenches(), b.get_link_wrenches())
        assert np.array_equal(a.get_marker_forces(), b.get_marker_forces())
        assert util.rel_l2(b.get_marker_forces(), o.get_marker_forces()) <= TOL_FORCE
        assert util.rel_l2(b.get_fields(f64=True)[1], o.get_fields(f64=True)[1]) <= TOL_FIELD


@pytest.mark.parametrize("free", [0, 1])
def test_emulated_swimming_fish_with_iterations(g, emu, free):
    """The Gym loop (body integration on the host between substeps) with three passes per substep."""
    P, Wl = g.BC_PERIODIC, g.BC_WALL
    kw = dict(nx=20, ny=18, nz=40, tau=0.8, collision=g.MRT, max_markers=3000, max_links=4, bc=[Wl, Wl, Wl, Wl, P, P], ib_iterations=3)
    a, b = g.Sim(backend="oracle", **kw), g.Sim(backend=emu, **kw)
    d = util.fish_desc(g, root=(10, 9, 12), free=free)
    d.joint_rate_max = 0.006      # tail-tip speed ~0.05: the passes enforce no-slip almost fully, so the fluid really moves that fast
    for s in (a, b):
        s.add_fish(d)
    rng = np.random.default_rng(0)
    for _ in range(3):
        act = rng.uniform(-1, 1, a.action_size()).astype(np.float32)
        for s in (a, b):
            s.set_action(act)
            s.step(6)
        assert np.abs(b.get_obs() - a.get_obs()).max() <= 2e-4
        wa, wb = a.get_link_wrenches(), b.get_link_wrenches()
        assert np.abs(wb - wa).max() / np.abs(wa).max() <= 5e-4
        assert util.rel_l2(b.get_fields(f64=True)[1], a.get_fields(f64=True)[1]) <= 5e-5


@pytest.mark.parametrize("passes", [1, 3, 5])
def test_passive_fish_is_advected_by_a_uniform_stream(g, emu, passes):
    """A neutrally buoyant, unactuated, free fish released at rest into a uniform stream: it must pick up the stream's velocity and
    the wrench must vanish (Galilean invariance of the coupling).  With 3 passes and the virtual mass of the one-pass scheme
    scaled by the number of passes this blew up within 40 steps — the forcing's period-2 memory (eigenvalue 2 (1 - l)^n - 1) against
    the explicit body update, body.hpp set_forcing_passes; with 2^n - 1 the body settles at 0.987 of the stream speed after 600 steps
    for any number of passes (the one-pass scheme: the same 0.987).  Product's host integrator against the oracle's on the way."""
    kw = dict(nx=24, ny=20, nz=48, tau=0.8, collision=g.MRT, max_markers=3000, max_links=4, ib_iterations=passes)
    a, b = g.Sim(backend="oracle", **kw), g.Sim(backend=emu, **kw)
    u = np.zeros((3,) + a.shape)
    u[0], u[2] = 0.01, 0.02
    for s in (a, b):
        s.set_fields(np.ones(a.shape), u)
        s.add_fish(util.fish_desc(g, root=(12, 10, 20), free=1, heading=0.0))
        s.set_action(np.zeros(s.action_size(), np.float32))
    peak = 0.0
    for _ in range(6):
        for s in (a, b):
            s.step(100)
        oa, ob = a.get_obs(), b.get_obs()
        assert np.abs(oa - ob).max() < 2e-4
        peak = max(peak, float(np.abs(a.get_link_wrenches()).max()))
    assert abs(oa[4] / 0.01 - 1) < 0.03 and abs(oa[6] / 0.02 - 1) < 0.03 and abs(oa[7]) < 1e-4, oa      # velocity x, z; yaw rate
    assert np.abs(a.get_link_wrenches()[:, :3].sum(0)).max() < 0.02 and peak < 5.0                         # start-up force was 24 (3 passes)
    for s in (a, b):
        s.close()


def test_host_staged_slabs_with_iterations_equal_oracle_slabs(g, emu):
    """z-slabs whose halos travel through the host (fg_halo_pack / unpack), a body inside one slab: each rank runs the
    passes over its own cells, as the oracle's ranks do."""
    kw = dict(nx=16, ny=14, nz=24, tau=0.8, collision=g.MRT, max_markers=300, max_links=1, ib_iterations=3)
    X = util.sphere_markers((8.3, 7.1, 6.2), 3.0, 110)       # stencils stay inside slab 0 (planes 0 .. 11)
    U = np.zeros_like(X)
    dV = np.ones(len(X), np.float32)
    A = g._abi
    one = g.Sim(backend="oracle", **kw)
    rho, u = util.smooth_fields(one.shape)
    one.set_fields(rho, u)
    one.set_markers(X, U, dV, np.zeros(len(X), np.int32))
    ranks = [g.Sim(backend=emu, n_ranks=2, rank=r, **kw) for r in range(2)]
    for r, s in enumerate(ranks):
        s.set_fields(rho[12 * r:12 * r + 12], u[:, 12 * r:12 * r + 12])
    ranks[0].set_markers(X, U, dV, np.zeros(len(X), np.int32))      # without fg_peer_connect_all a rank takes the markers inside its slab
    for _ in range(4):
        one.step(1)
        for s in ranks:
            s.step(1)
        msgs = [(s.halo_pack(A.ZLO), s.halo_pack(A.ZHI)) for s in ranks]
        ranks[0].halo_unpack(A.ZHI, msgs[1][0])
        ranks[0].halo_unpack(A.ZLO, msgs[1][1])
        ranks[1].halo_unpack(A.ZLO, msgs[0][1])
        ranks[1].halo_unpack(A.ZHI, msgs[0][0])
    got = np.concatenate([s.get_fields(f64=True)[1] for s in ranks], axis=1)
    assert util.rel_l2(got, one.get_fields(f64=True)[1]) <= TOL_FIELD
    assert util.rel_l2(ranks[0].get_marker_forces(), one.get_marker_forces()) <= TOL_FORCE


@pytest.mark.parametrize("n_ranks", [2, 3])
@pytest.mark.parametrize("passes", [2, 3])
def test_bodies_across_slab_faces_with_passes_equal_unsplit(g, emu, n_ranks, passes):
    """fg_peer_connect_all + ib_iterations > 1: every pass completes the gathered force E_k of a face-crossing marker with the
    neighbour's part (the U* exchange kernels, one exchange epoch per pass) before the correction is spread.  A sphere across a
    face, one across the periodic seam, one inside a slab; peered emulation ranks on threads against the unsplit run."""
    import test_slabs
    nz = 12 * n_ranks
    kw = dict(nx=16, ny=14, nz=nz, tau=0.8, collision=g.MRT, max_markers=600, max_links=3, body_force=[0, 0, 2e-5], ib_iterations=passes)
    whole = g.Sim(backend=emu, **kw)
    one_pass = g.Sim(backend=emu, **dict(kw, ib_iterations=1))
    parts = [g.Sim(backend=emu, n_ranks=n_ranks, rank=r, **kw) for r in range(n_ranks)]
    X = np.concatenate([util.sphere_markers((8.2, 7.1, 12.3), 3.0, 120), util.sphere_markers((5.0, 6.0, nz - 0.4), 2.5, 90),
                        util.sphere_markers((10.0, 8.0, 5.5), 2.0, 60)])
    U = np.zeros_like(X)
    U[:120, 2] = 0.01
    link = np.array([0] * 120 + [1] * 90 + [2] * 60, np.int32)
    dV = np.full(len(X), 0.9, np.float32)
    origins = [[8.2, 7.1, 12.3], [5.0, 6.0, nz - 0.4], [10.0, 8.0, 5.5]]
    rho, u = util.smooth_fields(whole.shape, amp=0.01)
    h = nz // n_ranks
    handles = [s.peer_export() for s in parts]
    for r, s in enumerate(parts):
        s.set_fields(rho[r * h:(r + 1) * h], u[:, r * h:(r + 1) * h])
        s.peer_connect_all(handles)
    for s in (whole, one_pass):
        s.set_fields(rho, u)
    for s in [whole, one_pass] + parts:
        s.set_markers(X, U, dV, link)
        s.set_link_origins(origins)
    for n in (1, 4, 4):                  # both parities, both buffer sets of the exchange, static reuse of the band
        whole.step(n)
        one_pass.step(n)
        test_slabs._run_threads([lambda s=s: s.step(n) for s in parts])
        f = whole.get_populations()
        fs = np.concatenate([s.get_populations() for s in parts], axis=1)
        assert np.abs(f - fs).max() < 5e-7
        w = whole.get_link_wrenches()
        for s in parts:
            ws = s.get_link_wrenches()
            assert np.abs(ws - w).max() / np.abs(w).max() < 1e-5
            assert np.array_equal(ws, parts[0].get_link_wrenches())          # bit-identical on every rank
        _, owner = parts[0].get_index_map()
        fm_p = [s.get_marker_forces() for s in parts]
        fm_own = np.stack([fm_p[owner[k]][k] for k in range(len(X))])
        assert util.rel_l2(fm_own, whole.get_marker_forces()) < 1e-4
        # a face-crossing marker carries the same completed force on both ranks that hold it
        crossing = [k for k in range(len(X)) if sum(bool(np.any(f_[k] != 0)) for f_ in fm_p) >= 2]
        assert len(crossing) >= 20
        for k in crossing:
            held = [f_[k] for f_ in fm_p if np.any(f_[k] != 0)]
            assert all(np.abs(h_ - held[0]).max() < 2e-6 for h_ in held)
    assert np.abs(whole.get_link_wrenches() - one_pass.get_link_wrenches()).max() / np.abs(w).max() > 1e-2     # the passes do act
    # the marker set re-sent (band rebuilt) between steps
    X2 = X.copy()
    X2[:, 2] += 0.7
    for s in [whole] + parts:
        s.set_markers(X2, U, dV, link)
    whole.step(3)
    test_slabs._run_threads([lambda s=s: s.step(3) for s in parts])
    fs = np.concatenate([s.get_populations() for s in parts], axis=1)
    assert np.abs(whole.get_populations() - fs).max() < 5e-7


@pytest.mark.parametrize("passes", [2, 3])
def test_random_bodies_across_slab_faces_with_passes(g, emu, passes):
    """The randomised generator of test_random_cases.py for bodies across faces (2 - 4 peered ranks, drifting clouds that
    straddle faces, wrap around z or leave the box; walls, inlet / outlet, overlap and plane split switched at random)."""
    import test_random_cases as T
    bad, ran = [], 0
    for seed in range(30):
        worst, kw, n_ranks = T.run_bodies_across_slabs_case(g, emu, seed, passes)
        if worst is None:
            continue
        ran += 1
        if worst["f"] > 1e-6 or worst["wrench"] > 1e-4:
            bad.append((seed, worst, n_ranks, kw))
    assert not bad, bad[:3]
    assert ran >= 24


def test_fish_swims_across_a_slab_face_with_passes(g, emu):
    kw = dict(nx=20, ny=18, nz=48, tau=0.8, max_markers=4000, max_links=8, ib_iterations=2)
    import test_slabs
    whole = g.Sim(backend=emu, **kw)
    parts = [g.Sim(backend=emu, n_ranks=2, rank=r, **kw) for r in range(2)]
    handles = [s.peer_export() for s in parts]
    for s in parts:
        s.peer_connect_all(handles)
    d = util.fish_desc(g, root=(10, 9, 17))        # head in slab 0, tail links reach into slab 1 (face at z = 24)
    d.joint_rate_max = 0.006
    for s in [whole] + parts:
        s.add_fish(d)
    for it in range(4):
        act = np.sin(0.5 * it + np.arange(3))
        whole.set_action(act)
        whole.step(6)
        for s in parts:
            s.set_action(act)
        test_slabs._run_threads([lambda s=s: s.step(6) for s in parts])
        ow = whole.get_obs()
        for s in parts:
            assert np.abs(s.get_obs() - ow).max() < 1e-4
        assert np.array_equal(parts[0].get_obs(), parts[1].get_obs())        # replicated integrators stay bit-identical
    fs = np.concatenate([s.get_populations() for s in parts], axis=1)
    assert np.abs(whole.get_populations() - fs).max() < 1e-6


def test_error_paths(g, emu):
    for backend in ("oracle", emu):
        for bad in (-1, 17):
            with pytest.raises(g.FgError) as e:
                g.Sim(backend=backend, nx=8, ny=8, nz=8, max_markers=10, ib_iterations=bad)
            assert e.value.code == g._abi.FG_EINVAL and "ib_iterations" in str(e.value)


@pytest.mark.parametrize("passes", [2, 4])
def test_random_moving_marker_clouds_with_passes(g, emu, passes):
    """The randomised generator of test_random_cases.py (drifting, replaced, removed marker clouds; walls, obstacles, plane
    split, fused pairs, 1-3 substeps per call) with several passes per substep, emulated kernels against the oracle."""
    import test_random_cases as T
    limits = dict(T.LIMITS, probe=5e-6)
    bad, ran = [], 0
    for seed in range(40):
        worst, kw, _, _ = T.run_moving_markers_case(g, emu, seed, passes)
        if worst is None:
            continue
        ran += 1
        if any(worst[k] > limits[k] for k in worst):
            bad.append((seed, worst, kw))
    assert not bad, bad[:3]
    assert ran >= 30


def _env(g, backend, passes):
    from gym_fish_b200.env import EnvConfig, FishEnv, FishSpec
    fish = FishSpec(links=((8, 2.5), (7, 2.5), (6, 2), (5, 1.5)), root=(12, 10, 12), joint_rate_max=0.006, joint_limit=0.6)
    return FishEnv(EnvConfig(grid=(24, 20, 40), tau=0.8, n_substeps=5, max_episode_steps=4, fish=(fish,), ib_iterations=passes), backend=backend)


def test_env_option_reaches_the_library(g, emu):
    """EnvConfig.ib_iterations: the Gym loop with two passes differs from one pass and matches the oracle's two-pass loop."""
    one, a, b = _env(g, "oracle", 1), _env(g, "oracle", 2), _env(g, emu, 2)
    for e in (one, a, b):
        e.reset(seed=1)
    for t in range(3):
        act = np.sin(t + np.arange(3)).astype(np.float32)
        o1, *_ = one.step(act)
        oa, ra, _, _, info = a.step(act)
        ob, rb, *_ = b.step(act)
        assert not info["diverged"]
        assert np.abs(oa - ob).max() < 1e-4 and abs(ra - rb) < 1e-5
    assert np.abs(oa - o1).max() > 1e-4       # the stronger coupling moves the swimmer differently
    for e in (one, a, b):
        e.close()


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("iters", [2, 4])
def test_gpu_matches_oracle(g, cuda, iters):
    _run_walled(g, cuda, iters)


@pytest.mark.gpu
def test_gpu_static_cache_and_many_markers(g, cuda, monkeypatch):
    """8 spheres, 12 000 markers on 96 x 64 x 64 (above the 8 192-marker threshold of the static-body cache): the first
    step fills the cache, the following ones read it in every pass; then the set is re-sent, shifted."""
    kw = dict(nx=96, ny=64, nz=64, tau=0.7, collision=g.MRT, max_markers=12000, max_links=8, ib_iterations=3, body_force=[0, 0, 1e-5])
    a, b = g.Sim(backend="oracle", **kw), g.Sim(backend=cuda, **kw)
    cs = [(16 + 21 * (i % 4), 18 + 28 * (i // 4), 20 + 6 * i) for i in range(8)]
    X = np.concatenate([util.sphere_markers(c, 9.0, 1500) for c in cs])
    link = np.repeat(np.arange(8, dtype=np.int32), 1500)
    dV = np.full(len(X), 4 * np.pi * 81 / 1500, np.float32)
    rho, u = util.smooth_fields(a.shape)
    for s in (a, b):
        s.set_fields(rho, u)
    for shift in (0.0, 0.4):
        Xs = X.copy()
        Xs[:, 0] += shift
        for s in (a, b):
            s.set_markers(Xs, np.zeros_like(Xs), dV, link)
            s.set_link_origins(np.array(cs, dtype=np.float64) + [shift, 0, 0])
        _compare(a, b, steps=(1, 1, 3))
