"""Host-side logic of the step orchestration, through the C ABI on the emulated device: when the plane split applies,
the staged (deferred) marker upload, call timing.  No GPU needed; the same code drives the CUDA device policy."""
import numpy as np
import pytest

import util


def _sim(g, emu, **kw):
    base = dict(nx=12, ny=10, nz=64, tau=0.8, collision=g.MRT, max_markers=600, max_links=2, split_min_cells=1)
    base.update(kw)
    return g.Sim(backend=emu, **base)


def _sphere(z, r=2.5, n=80, x=6.2, y=5.1):
    X = util.sphere_markers((x, y, z), r, n)
    return X, np.zeros_like(X), np.ones(n, np.float32)


@pytest.mark.parametrize("z, expect_split", [(20.0, True),      # stencils inside planes ~16..24: most planes are far
                                               (58.0, True),      # near the top: only planes below are far
                                               (63.2, False)])    # wraps around the periodic z boundary: everywhere
def test_plane_split_applies_only_when_far_planes_exist(g, emu, z, expect_split):
    s = _sim(g, emu)
    s.set_markers(*_sphere(z))
    s.step(2)
    assert (s.stats().split_substeps == 2) == expect_split


def test_plane_split_needs_a_quarter_of_the_planes_and_enough_cells(g, emu):
    s = _sim(g, emu)
    X = np.concatenate([_sphere(8.0)[0], _sphere(30.0)[0], _sphere(52.0)[0]])      # bodies spread over the whole height
    s.set_markers(X, np.zeros_like(X), np.ones(len(X), np.float32))
    s.step(1)
    assert s.stats().split_substeps == 0
    t = _sim(g, emu, split_min_cells=0)          # default threshold: 2^20 far cells — this lattice has 7 680 cells in all
    t.set_markers(*_sphere(20.0))
    t.step(1)
    assert t.stats().split_substeps == 0
    u = _sim(g, emu, flags=g._abi.FLAG_NO_SPLIT)
    u.set_markers(*_sphere(20.0))
    u.step(1)
    assert u.stats().split_substeps == 0


def test_staged_marker_upload_is_flushed_by_whoever_needs_the_markers(g, emu):
    """fg_set_markers only stages the message; the copy is queued by the next step (after the far-plane launch), by
    fg_get_markers or by fg_set_link_origins.  Whatever the order of calls, the device sees the LAST marker set."""
    a, b = _sim(g, emu), _sim(g, emu)
    X1, U1, dV1 = _sphere(20.0)
    X2, U2, dV2 = _sphere(33.0, n=90)
    rho, u = util.smooth_fields(a.shape, amp=0.01)
    for s in (a, b):
        s.set_fields(rho, u)
    # a: two sets in a row (the first is superseded before it was ever uploaded), read-back, origins, then step
    a.set_markers(X1, U1, dV1)
    a.set_markers(X2, U2, dV2)
    Xr, Ur, lr = a.get_markers()
    assert np.array_equal(Xr, X2) and len(lr) == 90
    a.set_link_origins([[6.2, 5.1, 33.0]])
    a.step(3)
    # b: the plain order
    b.set_markers(X2, U2, dV2)
    b.set_link_origins([[6.2, 5.1, 33.0]])
    b.step(3)
    assert np.array_equal(a.get_populations(), b.get_populations())
    assert np.array_equal(a.get_link_wrenches(), b.get_link_wrenches())
    assert a.stats().n_markers == 90


def test_call_timing_is_reported_once_the_call_has_finished(g, emu):
    s = _sim(g, emu)
    assert s.stats().last_step_ms == 0 and s.stats().last_mlups == 0
    s.step(4)
    s.sync()
    st = s.stats()
    assert st.steps == 4 and st.last_step_ms > 0
    assert abs(st.last_mlups - st.cells * 4 / st.last_step_ms / 1e3) < 1e-6 * st.last_mlups
    s2 = _sim(g, emu, flags=g._abi.FLAG_SYNC_STEP)
    s2.step(1)
    assert s2.stats().last_step_ms > 0


@pytest.mark.parametrize("backend", ["oracle", "emu"])
@pytest.mark.parametrize("how", ["reset", "set_fields", "set_populations"])
def test_reset_clears_an_owed_halo_exchange(g, emu, backend, how):
    """Host-staged z-slabs: fg_step(1) leaves a halo exchange owed.  fg_reset / fg_set_fields / fg_set_populations rebuild
    the ghost planes, so the handle must be steppable again afterwards — same behaviour on the oracle and the product."""
    lib = emu if backend == "emu" else "oracle"
    kw = dict(nx=8, ny=6, nz=8, tau=0.8)
    parts = [g.Sim(backend=lib, n_ranks=2, rank=r, **kw) for r in range(2)]
    for s in parts:
        s.step(1)
        with pytest.raises(g.FgError):           # the exchange is owed: stepping on is refused
            s.step(1)
        if how == "reset":
            s.reset()
        elif how == "set_fields":
            s.set_fields(np.ones(s.shape, np.float32), np.zeros((3,) + s.shape, np.float32))
        else:
            s.set_populations(np.broadcast_to(g._abi.W[:, None, None, None], (19,) + s.shape).astype(np.float32))
        s.step(1)                                # accepted again
    # and the protocol still works from here: exchange, step
    msgs = [(s.halo_pack(g._abi.ZLO), s.halo_pack(g._abi.ZHI)) for s in parts]
    parts[0].halo_unpack(g._abi.ZHI, msgs[1][0]); parts[0].halo_unpack(g._abi.ZLO, msgs[1][1])
    parts[1].halo_unpack(g._abi.ZLO, msgs[0][1]); parts[1].halo_unpack(g._abi.ZHI, msgs[0][0])
    for s in parts:
        s.step(1)


def test_peer_connect_all_after_set_markers_is_refused(g, emu):
    """The marker message on the device has another layout once bodies may cross slab faces (a global-id column): a marker
    set sent BEFORE fg_peer_connect_all would be read with the wrong layout, so the call is refused."""
    kw = dict(nx=16, ny=14, nz=24, tau=0.8, max_markers=600, max_links=2)
    parts = [g.Sim(backend=emu, n_ranks=2, rank=r, **kw) for r in range(2)]
    handles = [s.peer_export() for s in parts]
    X = util.sphere_markers((8.2, 7.1, 5.5), 2.0, 60)         # inside slab 0: accepted without the exchange
    parts[0].set_markers(X, np.zeros_like(X), np.ones(60, np.float32))
    with pytest.raises(g.FgError) as e:
        parts[0].peer_connect_all(handles)
    assert e.value.code == g._abi.FG_ESTATE
    parts[0].set_markers(X[:0], X[:0], np.ones(0, np.float32))   # withdrawn: connecting works, then the set is sent again
    parts[0].peer_connect_all(handles)
    parts[1].peer_connect_all(handles)
    for s in parts:
        s.set_markers(X, np.zeros_like(X), np.ones(60, np.float32))
        s.set_link_origins([[8.2, 7.1, 5.5]])


def test_fluid_divergence_guard_counts_the_same_cells_on_oracle_and_product(g, emu):
    """fg_check_finite (ABI v7): cells whose rest population is not finite or out of range — 0 on a healthy run, the same
    count on the oracle and on the product's kernel body once a NaN has been planted and spreads with the streaming."""
    counts = []
    for be in ("oracle", emu):
        s = g.Sim(backend=be, nx=32, ny=24, nz=16, tau=0.8, collision=g.BGK)
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
        counts.append((n1, s.check_finite()))
    assert counts[0] == counts[1] and counts[0][0] >= 1 and counts[0][1] > counts[0][0]


def test_env_reports_a_diverged_fluid(g, emu):
    """env.py: info["diverged"] looks at the FLUID (fg_check_finite), not only at the body observation."""
    from gym_fish_b200.env import EnvConfig, FishEnv, FishSpec
    fish = FishSpec(links=((8, 2.5), (7, 2.5)), root=(12, 10, 14), free_root=False)
    env = FishEnv(EnvConfig(grid=(24, 20, 32), tau=0.8, collision=g.BGK, n_substeps=2, fish=(fish,)), backend=emu)
    env.reset(seed=0)
    _, _, term, _, info = env.step(np.zeros(1, np.float32))
    assert not term and info["fluid_bad_cells"] == 0 and not info["diverged"]
    f = env.sim.get_populations()
    f[:, 20:24] = np.nan                      # the fluid far from the (pinned) body blows up: the observation stays finite
    env.sim.set_populations(f)
    obs, _, term, _, info = env.step(np.zeros(1, np.float32))
    assert np.isfinite(obs).all() and info["fluid_bad_cells"] > 0 and info["diverged"] and term
    env.close()
