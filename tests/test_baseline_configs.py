"""Oracle parity at the sizes BASELINE.json `configs` names — the workloads bench.py times, compared with the fp64 oracle
through the C ABI on the same seeded inputs ("next round" item 1).

  configs[1]  256x128x128 MRT channel, fixed immersed sphere        bench.make_sim("sphere_256x128x128"), 100 steps
  configs[2]  512x256x256 tank, one 5-link fish, Gym substeps       bench.make_sim("tank_512x256x256"), 20 substeps
  configs[4]  10^5 immersed-boundary markers (64 spheres)           256^3 periodic box, static and re-sent every step

Tolerances are BASELINE.json:5's: rel-L2 of velocity / density <= 1e-5, link wrenches within 1e-4, marker->grid index
maps (and the band size they imply) bit-exact.  The oracle needs seconds per case on the GPU box's host cores; the
tank falls back to half the size on every axis when the host has too little memory for the fp64 two-lattice oracle.

The same cases run through the CPU emulation of the kernel bodies at a reduced height in `-m "not gpu"` (the size-
independent part: the workload set-up and the tolerances), so a failure on the GPU is a failure of the GPU path.
"""
import os

import numpy as np
import pytest

import bench
import util

TOL_FIELD = 1e-5
TOL_FORCE = 1e-4


def _host_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2 ** 30
    except Exception:       # noqa: BLE001
        return 16.0


def _fields_match(a, b):
    ra, ua = a.get_fields(f64=True)
    rb, ub = b.get_fields(f64=True)
    assert np.isfinite(ub).all() and np.isfinite(rb).all()
    eu, er = util.rel_l2(ub, ua), util.rel_l2(rb, ra)
    assert eu <= TOL_FIELD and er <= TOL_FIELD, (eu, er)
    return eu, er


def _sphere_channel(g, backend, shrink=1):
    ref, m = bench.make_sim(g, "oracle", "sphere_256x128x128", 0, 1, 0, shrink=shrink)
    dev, _ = bench.make_sim(g, backend, "sphere_256x128x128", 0, 1, 0, shrink=shrink)
    for s in (ref, dev):
        s.step(100)
    ba, oa = ref.get_index_map()
    bb, ob = dev.get_index_map()
    assert np.array_equal(ba, bb) and np.array_equal(oa, ob)                       # bit-exact
    assert ref.stats().band_cells == dev.stats().band_cells > 0
    wa, wb = ref.get_link_wrenches(), dev.get_link_wrenches()
    assert np.abs(wb - wa).max() / np.abs(wa).max() <= TOL_FORCE
    assert util.rel_l2(dev.get_marker_forces(), ref.get_marker_forces()) <= TOL_FORCE
    _fields_match(ref, dev)
    ref.close(); dev.close()


@pytest.mark.gpu
def test_config1_sphere_channel_256x128x128_matches_oracle(g, cuda):
    """BASELINE.json configs[1] at full size, exactly the lattice / markers / initial state bench.py times."""
    util.register_oracle(g)
    _sphere_channel(g, cuda)
    # the timed path itself: plane split on, i.e. every substep of the run above took the far-plane branch
    dev, _ = bench.make_sim(g, cuda, "sphere_256x128x128", 0, 1, 0)
    dev.step(4)
    assert dev.stats().split_substeps == 4
    dev.close()


def test_config1_sphere_channel_reduced_height_emulated(g, emu):
    util.register_oracle(g)
    _sphere_channel(g, emu, shrink=4)    # 32 x 32 x 64, sphere of diameter 6


def _tank(g, backend, shrink):
    ref, _ = bench.make_sim(g, "oracle", "tank_512x256x256", 0, 1, 0, shrink=shrink)
    dev, _ = bench.make_sim(g, backend, "tank_512x256x256", 0, 1, 0, shrink=shrink)
    assert ref.stats().n_markers == dev.stats().n_markers > 0
    for it in range(2):                      # two env steps of 20 / 10 substeps: the body moves, the band is rebuilt every substep
        act = np.sin(0.3 * it + np.arange(ref.action_size())).astype(np.float32)
        for s in (ref, dev):
            s.set_action(act)
            s.step(20 if it == 0 else 10)
        wa, wb = ref.get_link_wrenches(), dev.get_link_wrenches()
        assert np.abs(wb - wa).max() / np.abs(wa).max() <= TOL_FORCE
        assert np.abs(ref.get_obs() - dev.get_obs()).max() <= 1e-4
    ba, _ = ref.get_index_map()
    bb, _ = dev.get_index_map()
    assert np.array_equal(ba, bb)
    assert ref.stats().band_cells == dev.stats().band_cells
    _fields_match(ref, dev)
    ref.close(); dev.close()


@pytest.mark.gpu
def test_config2_tank_with_swimming_fish_matches_oracle(g, cuda):
    """BASELINE.json configs[2]: 512x256x256 tank (walls on x and y, z periodic), one 5-link fish, Gym-style substeps.
    Full size needs ~11 GB of host memory for the oracle's two fp64 lattices; with less the case halves every axis."""
    util.register_oracle(g)
    _tank(g, cuda, shrink=1 if _host_gb() > 24 else 2)


def test_config2_tank_reduced_emulated(g, emu):
    util.register_oracle(g)
    _tank(g, emu, shrink=4)


def _many_spheres(n_side, spacing, R, z_shift=0.0):
    n1 = int(round(4 * np.pi * R * R))
    Xs, links, origins = [], [], []
    for s_ in range(n_side ** 3):
        c = (spacing * 0.5 + 0.3 + spacing * (s_ % n_side), spacing * 0.5 + 0.1 + spacing * ((s_ // n_side) % n_side),
             spacing * 0.5 + 0.2 + spacing * (s_ // n_side ** 2) + z_shift)
        Xs.append(util.sphere_markers(c, R, n1)); links.append(np.full(n1, s_, np.int32)); origins.append(c)
    X = np.concatenate(Xs)
    return X, np.concatenate(links), np.array(origins), np.full(len(X), 4 * np.pi * R * R / n1, np.float32)


def _marker_cloud(g, backend, n_box, n_side, R, min_markers, flags=0):
    """Static cloud for 5 steps (index map and band reused), then re-sent and displaced every step for 5 more."""
    kw = dict(nx=n_box, ny=n_box, nz=n_box, tau=0.6, collision=g.MRT, max_markers=110000, max_links=64)
    ref, dev = g.Sim(backend="oracle", **kw), g.Sim(backend=backend, flags=flags, **kw)
    X, link, origins, dV = _many_spheres(n_side, n_box / n_side, R)
    assert len(X) >= min_markers
    rng = np.random.default_rng(7)
    rho = np.ones(ref.shape, np.float32)
    u = (1e-3 * rng.standard_normal((3,) + ref.shape)).astype(np.float32)
    u[2] += 0.02
    U = np.zeros_like(X)

    def check(tag):
        ba, oa = ref.get_index_map()
        bb, ob = dev.get_index_map()
        assert np.array_equal(ba, bb) and np.array_equal(oa, ob), tag
        assert ref.stats().band_cells == dev.stats().band_cells > 0, tag
        fa, fb = ref.get_marker_forces(), dev.get_marker_forces()
        assert util.rel_l2(fb, fa) <= TOL_FORCE, tag
        wa, wb = ref.get_link_wrenches(), dev.get_link_wrenches()
        assert np.abs(wb - wa).max() / np.abs(wa).max() <= TOL_FORCE, tag
        assert util.rel_l2(dev.get_marker_velocities(), ref.get_marker_velocities()) <= TOL_FORCE, tag

    for s in (ref, dev):
        s.set_fields(rho, u)
        s.set_markers(X, U, dV, link)
        s.set_link_origins(origins)
        s.step(5)
    check("static")
    for it in range(5):
        # the whole cloud drifts 0.7 cells per step along z and carries a velocity: old band cells are cleared and new
        # ones registered every step
        Xm, _, om, _ = _many_spheres(n_side, n_box / n_side, R, z_shift=0.7 * (it + 1))
        Um = np.zeros_like(Xm)
        Um[:, 2] = 0.01
        for s in (ref, dev):
            s.set_markers(Xm, Um, dV, link)
            s.set_link_origins(om)
            s.step(1)
        if it in (0, 4):
            check(f"moving {it}")
    check("moving end")
    _fields_match(ref, dev)
    assert util.rel_l2(dev.get_force_field(), ref.get_force_field()) <= TOL_FORCE
    ref.close(); dev.close()


@pytest.mark.gpu
def test_config4_1e5_markers_on_256_cubed_match_oracle(g, cuda):
    """10^5 markers (64 spheres of 1 548 markers) — the marker count of the IB-overhead target in BASELINE.json:5 — on a
    256^3 box: first oracle parity for the CTA-aggregated band registration (IbIndexMarkT<true>), the shuffle-weight
    interpolate / spread launch and IbClearBand of a moving cloud at this scale."""
    util.register_oracle(g)
    _marker_cloud(g, cuda, 256, 4, 11.1, 99000)


@pytest.mark.gpu
def test_config4_tile_spread_variant_matches_oracle(g, cuda):
    """The A/B variant of the spreading (FG_FLAG_IB_TILE_SPREAD: per-CTA shared-memory table, one global reduction per touched
    band cell) against the oracle on 27 spheres / 4 x 10^4 markers, static and moving."""
    util.register_oracle(g)
    _marker_cloud(g, cuda, 192, 3, 11.1, 41000, flags=g._abi.FLAG_IB_TILE_SPREAD)


def test_config4_marker_cloud_reduced_emulated(g, emu):
    util.register_oracle(g)
    _marker_cloud(g, emu, 48, 2, 7.0, 4000)
