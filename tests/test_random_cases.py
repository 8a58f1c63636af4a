"""Randomised parity: the kernel bodies (emulated on the CPU; the real library on the GPU) against the fp64 oracle over
seeded random configurations — ragged sizes down to one cell, every valid combination of periodic / wall / inlet / outlet
faces, moving walls, BGK / MRT, body force, random obstacle fields, random marker clouds (inside, on the faces and
outside the box), the scheduling flags.  The fixed case table (util.parity_cases) covers what was thought of; this covers
what was not.  Its first run found one: a zero-gradient outlet above an obstacle cell copied a stale value
(lbm_core.cuh ZFaceOp), kept below as a named regression case.
"""
import os

import numpy as np
import pytest

import util


def random_case(g, rng):
    P, Wl, IN, OUT = g.BC_PERIODIC, g.BC_WALL, g.BC_INLET, g.BC_OUTLET
    A = g._abi
    nx, ny, nz = int(rng.integers(1, 14)), int(rng.integers(1, 12)), int(rng.integers(2, 12))
    if rng.random() < 0.15:
        nx = int(rng.integers(120, 140))          # rows wider than one CTA: interior warps and both edge warps
    bc = [P] * 6
    if rng.random() < 0.5:
        bc[0] = bc[1] = Wl
    if rng.random() < 0.5:
        bc[2] = bc[3] = Wl
    r = rng.random()
    if r < 0.25:
        bc[4] = bc[5] = Wl
    elif r < 0.45:
        bc[4], bc[5] = IN, OUT
    elif r < 0.6:
        bc[4], bc[5] = OUT, IN
    elif r < 0.7:
        bc[4], bc[5] = Wl, OUT
    elif r < 0.8:
        bc[4], bc[5] = IN, Wl
    kw = dict(nx=nx, ny=ny, nz=nz, tau=float(rng.uniform(0.55, 1.5)), collision=int(rng.integers(0, 2)), bc=bc)
    if rng.random() < 0.6:
        kw["body_force"] = [float(x) for x in rng.uniform(-3e-4, 3e-4, 3)]
    wu = {f: [float(x) for x in rng.uniform(-0.05, 0.05, 3)] for f in range(6) if bc[f] == Wl and rng.random() < 0.5}
    if wu:
        kw["wall_u"] = wu
    if IN in bc:
        kw["inlet_u"] = [float(x) for x in rng.uniform(-0.03, 0.03, 3)]
        if rng.random() < 0.5:
            kw["inlet_rho"] = float(rng.uniform(0.98, 1.02))
    solid = (rng.random((nz, ny, nx)) < rng.uniform(0.02, 0.3)).astype(np.uint8) if rng.random() < 0.4 else None
    markers = None
    if rng.random() < 0.5:
        n = int(rng.integers(1, 16))
        span = (-0.5, 1.5) if rng.random() < 0.3 else (0.0, 1.0)      # sometimes outside the box: wrapped or dropped nodes
        X = (rng.uniform(*span, (n, 3)) * [nx, ny, nz]).astype(np.float32)
        nl = int(rng.integers(1, 4))
        markers = (X, rng.uniform(-0.03, 0.03, (n, 3)).astype(np.float32), rng.uniform(0.2, 1.0, n).astype(np.float32),
                   np.sort(rng.integers(0, nl, n)).astype(np.int32), rng.uniform(0, 10, (nl, 3)))
        kw.update(max_markers=32, max_links=4)
    flags = 0
    for flag, p in ((A.FLAG_NO_SPLIT, 0.3), (A.FLAG_NO_XWARP, 0.2), (A.FLAG_NO_SWEEP_FLIP, 0.2), (A.FLAG_NO_GRAPHS, 0.2)):
        if rng.random() < p:
            flags |= flag
    kw["flags"] = flags
    if rng.random() < 0.5:
        kw["split_min_cells"] = 1
    return kw, solid, markers


def widen(g, rng, kw, solid, markers):
    """The same random configuration on rows where the two-cell kernels run: nx = 64 / 128 / 256 (their defaults: NARROW rows
    and whole CTAs) or another even width with the forms forced by flag; few rows and planes so that a case stays cheap."""
    A = g._abi
    nx = int(rng.choice([64, 128, 256, 2 * int(rng.integers(2, 40))]))
    ny, nz = min(kw["ny"], 5), min(kw["nz"], 5)
    old = (kw["nx"], kw["ny"], kw["nz"])
    kw = dict(kw, nx=nx, ny=ny, nz=max(nz, 2))
    if nx not in (64, 128, 256):
        kw["flags"] |= int(rng.choice([A.FLAG_EVEN_VEC2 | A.FLAG_ODD_VEC2, A.FLAG_EVEN_VEC2, A.FLAG_ODD_VEC2, A.FLAG_EVEN_VEC4 if nx % 4 == 0 else A.FLAG_ODD_VEC2]))
    if solid is not None:
        solid = (rng.random((kw["nz"], ny, nx)) < 0.05).astype(np.uint8) if rng.random() < 0.3 else None   # obstacles: mostly off (they force the checked kernels)
    if markers is not None:
        X = (markers[0] / np.array(old, np.float32) * np.array([nx, ny, kw["nz"]], np.float32)).astype(np.float32)
        markers = (X,) + tuple(markers[1:])
    return kw, solid, markers


def run_case(g, backend, seed, wide=False, solid_force=False):
    """Worst absolute differences to the oracle over 5 checkpoints (9 steps, both parities), or None if the random
    configuration is not a stable simulation (the comparison of two diverging runs means nothing)."""
    rng = np.random.default_rng(seed)
    kw, solid, markers = random_case(g, rng)
    if wide:
        kw, solid, markers = widen(g, np.random.default_rng(10 ** 6 + seed), kw, solid, markers)
    a, b = g.Sim(backend="oracle", **kw), g.Sim(backend=backend, **kw)
    rho, u = util.smooth_fields(a.shape)
    rho = (rho + 0.003 * rng.standard_normal(a.shape)).astype(np.float32)
    u = (u + 0.003 * rng.standard_normal((3,) + a.shape)).astype(np.float32)
    for s in (a, b):
        if solid is not None:
            s.set_solid(solid)
        s.set_fields(rho, u)
        if markers is not None:
            s.set_markers(*markers[:4])
            s.set_link_origins(markers[4])
    keep = 1 if solid is None else (solid == 0)
    worst = {}
    for n in (1, 1, 1, 2, 4):
        a.step(n)
        b.step(n)
        ra, ua = a.get_fields(f64=True)
        if not (np.isfinite(ra).all() and 0.7 < (ra * keep + (1 - keep)).min() and (ra * keep).max() < 1.3):
            worst = None
            break
        rb, ub = b.get_fields(f64=True)
        e = dict(u=np.abs((ua - ub) * keep).max(), rho=np.abs((ra - rb) * keep).max(),
                 f=np.abs((a.get_populations() - b.get_populations()) * keep).max())
        if markers is not None:
            (ba, oa), (bb, ob) = a.get_index_map(), b.get_index_map()
            wa, wb = a.get_link_wrenches(), b.get_link_wrenches()
            e.update(index_map=0.0 if np.array_equal(ba, bb) and np.array_equal(oa, ob) else 1.0,
                     band=abs(a.stats().band_cells - b.stats().band_cells),
                     Fm=np.abs(a.get_marker_forces() - b.get_marker_forces()).max(),
                     wrench=np.abs(wa - wb).max() / max(np.abs(wa).max(), 1e-3))
        if solid_force and solid is not None:
            # fg_get_solid_force: thousands of link terms of either sign, each known to the populations' 1e-6: compared on the
            # scale of the largest component or of one link's momentum (2 w_i ~ 0.1), whichever is larger
            o = [kw["nx"] / 2, kw["ny"] / 2, kw["nz"] / 2]
            fa, fb = a.get_solid_force(o), b.get_solid_force(o)
            e["solid_force"] = np.abs(fa - fb).max() / max(np.abs(fa).max(), 0.1)
        for k, v in e.items():
            worst[k] = max(worst.get(k, 0.0), float(v))
    a.close()
    b.close()
    return worst, kw


# absolute, on fields of size ~1 (rho), ~0.02 (u), ~0.05 (f): BASELINE.json:5's 1e-5 relative L2 with room to spare
LIMITS = dict(u=3e-6, rho=3e-6, f=1e-6, index_map=0.5, band=0.5, Fm=2e-5, wrench=1e-4, solid_force=1e-5)


def check(g, backend, seeds, wide=False, solid_force=False):
    bad, ran = [], 0
    for seed in seeds:
        worst, kw = run_case(g, backend, seed, wide, solid_force)
        if worst is None:
            continue
        ran += 1
        if any(worst[k] > LIMITS[k] for k in worst):
            bad.append((seed, worst, kw))
    assert not bad, bad[:3]
    assert ran >= 0.9 * len(seeds)        # the generator produces stable simulations almost always


def test_random_configurations_emulated_kernels_vs_oracle(g, emu):
    check(g, emu, range(400), solid_force=True)      # + the momentum-exchange read-out wherever the case has obstacles (measured worst: 1e-6)


def test_random_configurations_on_wide_rows_emulated_kernels_vs_oracle(g, emu):
    """The two-cell kernels (whole CTAs, NARROW rows, x walls, forced by flag on other even widths) under the random generator."""
    check(g, emu, range(150), wide=True, solid_force=True)


@pytest.mark.parametrize("seed", [4, 36, 132, 17, 53, 268])
def test_outlet_next_to_obstacles_regression(g, emu, seed):
    """Seeds that failed before the ZFaceOp fix: a zero-gradient outlet (or inlet / periodic wrap) whose boundary plane or
    the plane inside it holds obstacle cells, with and without markers whose stencils touch those cells."""
    worst, kw = run_case(g, emu, seed)
    assert worst is not None and all(worst[k] <= LIMITS[k] for k in worst), (worst, kw)


@pytest.mark.gpu
def test_random_configurations_cuda_vs_oracle(g, cuda):
    check(g, cuda, range(1000, 1150))
    check(g, cuda, range(2000, 2150), wide=True)


def run_slab_case(g, emu, seed):
    """A random configuration cut into 2 or 3 z-slabs (host-staged halo messages, or pushes into the neighbour's lattice
    with boundary-first ordering on or off) against the same configuration in one piece: populations bit-identical."""
    import test_slabs
    rng = np.random.default_rng(seed)
    kw, solid, _ = random_case(g, rng)
    n_ranks, h = int(rng.integers(2, 4)), int(rng.integers(2, 6))
    kw["nz"] = h * n_ranks
    if solid is not None:
        solid = (rng.random((kw["nz"], kw["ny"], kw["nx"])) < 0.15).astype(np.uint8)
    peer = rng.random() < 0.5
    kw.pop("max_markers", None)
    kw.pop("max_links", None)
    if peer and rng.random() < 0.5:
        kw["flags"] |= g._abi.FLAG_NO_OVERLAP
    periodic = kw["bc"][4] == g.BC_PERIODIC
    whole = g.Sim(backend=emu, **kw)
    parts = [g.Sim(backend=emu, n_ranks=n_ranks, rank=r, **kw) for r in range(n_ranks)]
    rho, u = util.smooth_fields(whole.shape)
    rho = (rho + 0.003 * rng.standard_normal(whole.shape)).astype(np.float32)
    u = (u + 0.003 * rng.standard_normal((3,) + whole.shape)).astype(np.float32)
    for s in [whole] + parts:
        if solid is not None:
            s.set_solid(solid)
    whole.set_fields(rho, u)
    for r, s in enumerate(parts):
        s.set_fields(rho[r * h:(r + 1) * h], u[:, r * h:(r + 1) * h])
    if peer:
        hs = [s.peer_export() for s in parts]
        for r, s in enumerate(parts):
            s.peer_connect(hs[(r - 1) % n_ranks] if (r > 0 or periodic) else None, hs[(r + 1) % n_ranks] if (r < n_ranks - 1 or periodic) else None)
    for _ in range(7):
        whole.step(1)
        for s in parts:
            s.step(1)
        if not peer:
            test_slabs.exchange_in_process(g, parts, periodic)
    keep = 1 if solid is None else (solid == 0)
    f, fs = whole.get_populations(), np.concatenate([s.get_populations() for s in parts], axis=1)
    for s in [whole] + parts:
        s.close()
    return np.array_equal(f * keep, fs * keep), kw


def test_random_configurations_in_slabs_equal_unsplit(g, emu):
    bad = [(seed, kw) for seed in range(120) for ok, kw in [run_slab_case(g, emu, seed)] if not ok]
    assert not bad, bad[:3]


def stable(rho, keep=1):
    return bool(np.isfinite(rho).all() and 0.7 < (rho * keep + (1 - keep)).min() and (rho * keep).max() < 1.3)


def run_moving_markers_case(g, backend, seed, passes=1):
    """A marker cloud that drifts (up to 1.2 planes per step), is sometimes replaced by another one, left alone (static
    reuse of index map and band) or removed, steps of 1-3 substeps, obstacles, probes, the plane split and the fused
    step pairs switched on at random: everything the host logic decides per substep, against the oracle."""
    A = g._abi
    rng = np.random.default_rng(seed)
    kw, solid, _ = random_case(g, rng)
    kw["nz"], kw["ny"] = int(rng.integers(4, 40)), max(kw["ny"], 4)
    kw["nx"] = max(kw["nx"], 4)
    nx, ny, nz = kw["nx"], kw["ny"], kw["nz"]
    if solid is not None:
        solid = (rng.random((nz, ny, nx)) < 0.05).astype(np.uint8)
    kw.update(max_markers=64, max_links=4)
    if passes > 1:
        kw["ib_iterations"] = passes        # multi-direct forcing (tests/test_multi_direct_forcing.py); drawn outside this generator's stream
    if rng.random() < 0.3:
        kw["flags"] |= A.FLAG_FUSED_PAIRS
    a, b = g.Sim(backend="oracle", **kw), g.Sim(backend=backend, **kw)
    rho, u = util.smooth_fields(a.shape)
    rho = (rho + 0.003 * rng.standard_normal(a.shape)).astype(np.float32)
    u = (u + 0.003 * rng.standard_normal((3,) + a.shape)).astype(np.float32)
    for s in (a, b):
        if solid is not None:
            s.set_solid(solid)
        s.set_fields(rho, u)
    nl = int(rng.integers(1, 4))

    def cloud(c):
        n = int(rng.integers(1, 24))
        return ((c + rng.uniform(-2.5, 2.5, (n, 3))).astype(np.float32), np.sort(rng.integers(0, nl, n)).astype(np.int32),
                rng.uniform(0.2, 1.0, n).astype(np.float32))

    c = rng.uniform(0.2, 0.8, 3) * [nx, ny, nz]
    X, link, dV = cloud(c)
    V = rng.uniform(-0.4, 0.4, 3) * [1, 1, 3]
    origins = rng.uniform(0, 10, (nl, 3))
    keep = 1 if solid is None else (solid == 0)
    worst, have = {}, False
    for _ in range(10):
        r = rng.random()
        if r < 0.55 or not have:
            if rng.random() < 0.15:
                X, link, dV = cloud(c)
            X, c = (X + V).astype(np.float32), c + V
            U = np.tile(V.astype(np.float32) * 0.05, (len(X), 1))
            for s in (a, b):
                s.set_markers(X, U, dV, link)
                s.set_link_origins(origins)
            have = True
        elif r < 0.65:
            for s in (a, b):
                s.set_markers(X[:0], U[:0], dV[:0], link[:0])
            have = False
        k = int(rng.integers(1, 4))
        a.step(k)
        b.step(k)
        ra, ua = a.get_fields(f64=True)
        if not stable(ra, keep):
            worst = None
            break
        rb, ub = b.get_fields(f64=True)
        e = dict(u=np.abs((ua - ub) * keep).max(), rho=np.abs((ra - rb) * keep).max(),
                 f=np.abs((a.get_populations() - b.get_populations()) * keep).max())
        if have:
            (ba, oa), (bb, ob) = a.get_index_map(), b.get_index_map()
            wa, wb = a.get_link_wrenches(), b.get_link_wrenches()
            pts = (rng.uniform(0, 1, (5, 3)) * [nx, ny, nz]).astype(np.float32)
            e.update(index_map=0.0 if np.array_equal(ba, bb) and np.array_equal(oa, ob) else 1.0,
                     band=abs(a.stats().band_cells - b.stats().band_cells),
                     Fm=np.abs(a.get_marker_forces() - b.get_marker_forces()).max(),
                     wrench=np.abs(wa - wb).max() / max(np.abs(wa).max(), 1e-3), probe=np.abs(a.probe(pts) - b.probe(pts)).max())
        for key, v in e.items():
            worst[key] = max(worst.get(key, 0.0), float(v))
    st = b.stats()
    a.close()
    b.close()
    return worst, kw, st.split_substeps > 0, st.pair_substeps > 0


def check_moving(g, backend, seeds):
    limits = dict(LIMITS, probe=5e-6)
    bad, ran, splits, pairs = [], 0, 0, 0
    for seed in seeds:
        worst, kw, split, pair = run_moving_markers_case(g, backend, seed)
        if worst is None:
            continue
        ran, splits, pairs = ran + 1, splits + split, pairs + pair
        if any(worst[k] > limits[k] for k in worst):
            bad.append((seed, worst, kw))
    assert not bad, bad[:3]
    assert ran >= 0.75 * len(seeds) and splits >= 0.1 * len(seeds) and pairs >= 2      # the interesting paths were taken


def test_random_moving_markers_emulated_kernels_vs_oracle(g, emu):
    check_moving(g, emu, range(120))


def run_fish_case(g, backend, seed, passes=1):
    """One or two random articulated fish (1-5 links, pinned or free, random servo gains / limits / densities), random
    actions, 1-7 substeps per call, now and then a reset: the product's host integrator (csrc/body.hpp) + IB + fluid
    against the oracle's independently written one (oracle/oracle_body.hpp)."""
    P, Wl = g.BC_PERIODIC, g.BC_WALL
    rng = np.random.default_rng(seed)
    nx, ny, nz = int(rng.integers(16, 26)), int(rng.integers(14, 22)), int(rng.integers(30, 56))
    kw = dict(nx=nx, ny=ny, nz=nz, tau=float(rng.uniform(0.6, 1.0)), collision=int(rng.integers(0, 2)),
              bc=[Wl] * 4 + [P] * 2 if rng.random() < 0.5 else [P] * 6, max_markers=6000, max_links=16)
    if rng.random() < 0.5:
        kw["split_min_cells"] = 1
    if passes > 1:
        kw["ib_iterations"] = passes        # multi-direct forcing; not drawn from this generator's stream
    a, b = g.Sim(backend="oracle", **kw), g.Sim(backend=backend, **kw)
    nf = int(rng.integers(1, 3))
    for f in range(nf):
        links = tuple((float(rng.uniform(4, 8)), float(rng.uniform(1.2, 2.8))) for _ in range(int(rng.integers(1, 6))))
        d = util.fish_desc(g, root=(nx / 2 + rng.uniform(-2, 2), ny / 2 + rng.uniform(-1, 1), nz * (0.3 + 0.35 * f) + rng.uniform(-2, 2)),
                           links=links, free=int(rng.random() < 0.6), heading=float(rng.uniform(-0.5, 0.5)))
        d.density_ratio, d.joint_gain = float(rng.uniform(0.8, 1.5)), float(rng.uniform(0.05, 0.4))
        d.joint_limit, d.joint_rate_max = float(rng.uniform(0.2, 0.8)), float(rng.uniform(0.005, 0.03))
        if rng.random() < 0.3:
            d.markers_per_link = int(rng.integers(8, 60))
        for s in (a, b):
            s.add_fish(d)
    assert a.stats().n_markers == b.stats().n_markers and a.obs_size() == b.obs_size()
    worst = dict(obs=0.0, wrench=0.0, u=0.0)
    for it in range(6):
        act = rng.uniform(-1.3, 1.3, a.action_size()).astype(np.float32)      # beyond [-1, 1]: clipped by the library
        k = int(rng.integers(1, 8))
        for s in (a, b):
            s.set_action(act)
            s.step(k)
        oa, ob = a.get_obs(), b.get_obs()
        ra, ua = a.get_fields(f64=True)
        if not (np.isfinite(oa).all() and np.isfinite(ra).all() and 0.85 < ra.min() and ra.max() < 1.15):
            worst = None            # fish too large for their tank: the coupled run itself diverges
            break
        wa, wb = a.get_link_wrenches(), b.get_link_wrenches()
        worst["obs"] = max(worst["obs"], float(np.abs(oa - ob).max()))
        worst["wrench"] = max(worst["wrench"], float(np.abs(wa - wb).max() / max(np.abs(wa).max(), 1e-6)))
        if it % 3 == 2:
            if rng.random() < 0.3:
                for s in (a, b):
                    s.reset(0)
            else:
                worst["u"] = max(worst["u"], util.rel_l2(b.get_fields(f64=True)[1], ua))
    a.close()
    b.close()
    return worst, kw


def check_fish(g, backend, seeds):
    bad, ran = [], 0
    for seed in seeds:
        worst, kw = run_fish_case(g, backend, seed)
        if worst is None:
            continue
        ran += 1
        if worst["obs"] > 1e-4 or worst["wrench"] > 1e-4 or worst["u"] > 1e-5:
            bad.append((seed, worst, kw))
    assert not bad, bad[:3]
    assert ran >= 0.5 * len(seeds)


def test_random_fish_emulated_kernels_vs_oracle(g, emu):
    check_fish(g, emu, range(24))


@pytest.mark.gpu
def test_random_moving_markers_and_fish_cuda_vs_oracle(g, cuda):
    check_moving(g, cuda, range(2000, 2080))
    check_fish(g, cuda, range(2000, 2012))


def run_bodies_across_slabs_case(g, emu, seed, passes=1):
    """fg_peer_connect_all on 2-4 slabs (ranks stepped from threads): drifting marker clouds that straddle faces, wrap
    around a periodic z axis or leave a non-periodic box, against the unsplit run.  Its first run found that a marker
    whose stencil lies entirely outside the box was culled on every rank and dropped out of its link's wrench."""
    import test_slabs
    P, Wl, IN, OUT = g.BC_PERIODIC, g.BC_WALL, g.BC_INLET, g.BC_OUTLET
    A = g._abi
    rng = np.random.default_rng(seed)
    n_ranks, h = int(rng.integers(2, 5)), int(rng.integers(4, 10))
    nx, ny, nz = int(rng.integers(6, 16)), int(rng.integers(6, 14)), h * n_ranks
    zlo, zhi = [(P, P), (P, P), (Wl, Wl), (IN, OUT)][int(rng.integers(0, 4))]
    bc = [P, P, P, P, zlo, zhi]
    if rng.random() < 0.4:
        bc[0] = bc[1] = Wl
    if rng.random() < 0.4:
        bc[2] = bc[3] = Wl
    kw = dict(nx=nx, ny=ny, nz=nz, tau=float(rng.uniform(0.6, 1.2)), collision=int(rng.integers(0, 2)), bc=bc, max_markers=600,
              max_links=4, body_force=[0, 0, float(rng.uniform(0, 3e-5))], inlet_u=[0, 0, 0.02], flags=0)
    for flag in (A.FLAG_NO_OVERLAP, A.FLAG_NO_SPLIT):
        if rng.random() < 0.3:
            kw["flags"] |= flag
    if rng.random() < 0.6:
        kw["split_min_cells"] = 1
    if passes > 1:
        kw["ib_iterations"] = passes        # multi-direct forcing across faces; set outside this generator's random stream
    whole = g.Sim(backend=emu, **kw)
    parts = [g.Sim(backend=emu, n_ranks=n_ranks, rank=r, **kw) for r in range(n_ranks)]
    rho, u = util.smooth_fields(whole.shape, amp=0.01)
    whole.set_fields(rho, u)
    handles = [s.peer_export() for s in parts]
    for r, s in enumerate(parts):
        s.set_fields(rho[r * h:(r + 1) * h], u[:, r * h:(r + 1) * h])
        s.peer_connect_all(handles)
    clouds = []
    for _ in range(int(rng.integers(1, 4))):
        c = np.array([rng.uniform(3, nx - 3), rng.uniform(3, ny - 3), rng.uniform(0, nz)])
        clouds.append([util.sphere_markers(c, float(rng.uniform(1.0, 3.0)), int(rng.integers(1, 80))), c, rng.uniform(-0.6, 0.6, 3) * [0.3, 0.3, 2.0]])
    worst = dict(f=0.0, wrench=0.0)
    for it in range(6):
        if it == 0 or rng.random() < 0.7:
            for cl in clouds:
                cl[0], cl[1] = (cl[0] + cl[2]).astype(np.float32), cl[1] + cl[2]
            X = np.concatenate([cl[0] for cl in clouds])
            link = np.concatenate([np.full(len(cl[0]), i, np.int32) for i, cl in enumerate(clouds)])
            U = np.concatenate([np.tile((cl[2] * 0.02).astype(np.float32), (len(cl[0]), 1)) for cl in clouds])
            for s in [whole] + parts:
                s.set_markers(X, U, np.ones(len(X), np.float32), link)
                s.set_link_origins([list(cl[1]) for cl in clouds])
        k = int(rng.integers(1, 4))
        whole.step(k)
        test_slabs._run_threads([lambda s=s: s.step(k) for s in parts])
        if not stable(whole.get_fields(f64=True)[0]):
            worst = None
            break
        f, fs = whole.get_populations(), np.concatenate([s.get_populations() for s in parts], axis=1)
        worst["f"] = max(worst["f"], float(np.abs(f - fs).max()))
        w = whole.get_link_wrenches()
        for s in parts:
            ws = s.get_link_wrenches()
            worst["wrench"] = max(worst["wrench"], float(np.abs(ws - w).max() / max(np.abs(w).max(), 1e-3)))
            assert np.array_equal(ws, parts[0].get_link_wrenches())          # replicated integrators need identical bits
        assert np.array_equal(whole.get_index_map()[0], parts[0].get_index_map()[0])
    for s in [whole] + parts:
        s.close()
    return worst, kw, n_ranks


def test_random_bodies_across_slab_faces_equal_unsplit(g, emu):
    bad, ran = [], 0
    for seed in range(100):
        worst, kw, n_ranks = run_bodies_across_slabs_case(g, emu, seed)
        if worst is None:
            continue
        ran += 1
        if worst["f"] > 1e-6 or worst["wrench"] > 1e-4:
            bad.append((seed, worst, n_ranks, kw))
    assert not bad, bad[:3]
    assert ran >= 80


def run_api_sequence_case(g, backend, seed, solid_force=False):
    """A random SEQUENCE of ABI calls, valid and invalid (too many markers, link ids beyond max_links, obstacles changed
    after an even or an odd step, read-outs before the first step, resets, zero-substep steps), issued to the oracle and to
    the product's host code: every call must have the same outcome (ok, or the same error code) on both, and every read-out
    must agree.  First run: a crash in the oracle (state touched before validation), a heap overflow in the Python binding
    (wrench buffer sized from a stale link count), and fg_set_solid after an even step changing the streaming of that step."""
    A = g._abi
    rng = np.random.default_rng(seed)
    kw, _, _ = random_case(g, rng)
    kw["nx"], kw["ny"], kw["nz"] = (max(kw["nx"], 4) if kw["nx"] < 100 else kw["nx"]), max(kw["ny"], 4), max(kw["nz"], 4)
    kw.update(max_markers=40, max_links=3)
    nx, ny, nz = kw["nx"], kw["ny"], kw["nz"]
    sims = [g.Sim(backend="oracle", **kw), g.Sim(backend=backend, **kw)]
    solid, worst, log = None, 0.0, []

    def both(fn, name):
        res = []
        for s in sims:
            try:
                res.append(("ok", fn(s)))
            except g.FgError as e:
                res.append(("err", e.code))
        log.append((name, res[0][0], res[1][0]))
        assert res[0][0] == res[1][0] and (res[0][0] == "ok" or res[0][1] == res[1][1]), (seed, name, res, log[-6:])
        return res[0][0] == "ok", res[0][1], res[1][1]

    try:
        for _ in range(14):
            op = int(rng.integers(0, 11))
            if op == 0:
                rho, u = util.smooth_fields((nz, ny, nx))
                rho = (rho + 0.002 * rng.standard_normal(rho.shape)).astype(np.float32)
                both(lambda s: s.set_fields(rho, u.astype(np.float32)), "set_fields")
            elif op == 1:
                n = int(rng.integers(0, 50))                                   # sometimes more than max_markers
                X = (rng.uniform(-0.2, 1.2, (n, 3)) * [nx, ny, nz]).astype(np.float32)
                U = rng.uniform(-0.02, 0.02, (n, 3)).astype(np.float32)
                link = rng.integers(0, int(rng.integers(1, 5)), n).astype(np.int32)   # sometimes a link id >= max_links
                both(lambda s: s.set_markers(X, U, rng.uniform(0.2, 1, n).astype(np.float32) * 0 + 0.5, link), "set_markers")
            elif op == 2:
                o = rng.uniform(0, 10, (int(rng.integers(1, 5)), 3))          # sometimes more than max_links
                both(lambda s: s.set_link_origins(o), "set_link_origins")
            elif op == 3:
                k = int(rng.integers(0, 4))
                both(lambda s: s.step(k), "step")
                # the momentum-exchange read-out after whatever came before (obstacles changed after an even or odd step, a
                # reset, populations set by hand, no obstacles at all); not drawn from the generator's stream
                if solid_force:
                    ok, fa, fb = both(lambda s: s.get_solid_force([nx / 2, ny / 2, nz / 2]), "get_solid_force")
                    if ok and np.isfinite(fa).all() and np.abs(fa).max() < 1e3:
                        worst = max(worst, float(np.abs(fa - fb).max() / max(np.abs(fa).max(), 0.1)))
            elif op == 4:
                both(lambda s: s.reset(0), "reset")
            elif op == 5:
                solid = (rng.random((nz, ny, nx)) < 0.05).astype(np.uint8) if rng.random() < 0.7 else None
                both(lambda s: s.set_solid(solid), "set_solid")
            elif op == 6:
                f = (A.W[:, None, None, None] * (1 + 0.01 * rng.standard_normal((19, nz, ny, nx)))).astype(np.float32)
                both(lambda s: s.set_populations(f), "set_populations")
            elif op == 7:
                ok, wa, wb = both(lambda s: s.get_link_wrenches().copy(), "get_link_wrenches")
                assert not ok or wa.shape == wb.shape, (seed, wa.shape, wb.shape, log[-6:])
                if ok and wa.size and np.abs(wa).max() < 1e3:
                    worst = max(worst, float(np.abs(wa - wb).max() / max(np.abs(wa).max(), 1e-3)))
            elif op == 8:
                ok, fa, fb = both(lambda s: s.get_force_field(), "get_force_field")
                if ok and np.isfinite(fa).all() and np.abs(fa).max() < 10:
                    worst = max(worst, float(np.abs(fa - fb).max()))
            elif op == 9:
                pts = (rng.uniform(0, 1, (4, 3)) * [nx, ny, nz]).astype(np.float32)
                both(lambda s: s.probe(pts), "probe")
            else:
                ok, (ra, ua), (rb, ub) = both(lambda s: s.get_fields(f64=True), "get_fields")
                keep = 1 if solid is None else (solid == 0)
                if not stable(ra, keep):
                    return None
                worst = max(worst, float(np.abs((ua - ub) * keep).max()))
        return worst
    finally:
        for s in sims:
            s.close()


def test_random_api_sequences_same_outcome_on_both_backends(g, emu):
    bad, ran = [], 0
    for seed in range(300):
        w = run_api_sequence_case(g, emu, seed, solid_force=True)
        if w is None:
            continue
        ran += 1
        if w > 2e-5:
            bad.append((seed, w))
    assert not bad, bad[:5]
    assert ran >= 250


def test_link_count_can_grow_without_the_marker_count_changing(g, emu):
    """Regression (found by the API fuzz under AddressSanitizer): the binding sized its wrench buffer from a link count it
    only refreshed when the MARKER count changed."""
    for backend in ("oracle", emu):
        s = g.Sim(backend=backend, nx=12, ny=10, nz=12, max_markers=8, max_links=4)
        X = util.sphere_markers((6, 5, 6), 2.0, 8)
        U = np.full_like(X, 0.01)
        s.set_markers(X, U, np.ones(8, np.float32), np.zeros(8, np.int32))
        s.step(1)
        assert s.get_link_wrenches().shape == (1, 6)
        s.set_markers(X, U, np.ones(8, np.float32), np.arange(8, dtype=np.int32) % 4)      # same 8 markers, 4 links now
        assert np.array_equal(s.get_link_wrenches(), np.zeros((4, 6)))                      # nothing stepped yet: zeros
        s.step(1)
        w = s.get_link_wrenches()
        assert w.shape == (4, 6) and np.abs(w).min(axis=1).max() > 0
        s.close()


@pytest.mark.parametrize("steps_before", [1, 2])
def test_obstacles_may_change_after_an_even_or_an_odd_step(g, emu, steps_before):
    """fg_set_solid between steps: after an even AA step the streaming of that step is still pending and must be completed
    with the OLD mask (sim.hpp finish_pending_streaming) — the oracle streams inside fg_step."""
    P, Wl = g.BC_PERIODIC, g.BC_WALL
    kw = dict(nx=10, ny=8, nz=9, tau=0.8, collision=g.MRT, bc=[P, P, Wl, Wl, P, P], body_force=[1e-4, 0, 2e-4])
    a, b = g.Sim(backend="oracle", **kw), g.Sim(backend=emu, **kw)
    rho, u = util.smooth_fields(a.shape)
    rng = np.random.default_rng(11)
    s1 = (rng.random(a.shape) < 0.1).astype(np.uint8)
    s2 = (rng.random(a.shape) < 0.1).astype(np.uint8)
    for s in (a, b):
        s.set_solid(s1)
        s.set_fields(rho, u)
        s.step(steps_before)
        s.set_solid(s2)
        s.step(3)
    keep = (s1 == 0) & (s2 == 0)
    assert np.abs((a.get_populations() - b.get_populations()) * keep).max() < 2e-7
    assert b.stats().parity == (3 if steps_before == 1 else 5) % 2       # the flush brought an odd parity back to even
    a.close()
    b.close()


def test_random_fish_across_slab_faces(g, emu):
    """A random free or pinned fish placed anywhere along z (so it usually straddles a slab face or the periodic wrap),
    2-3 slabs stepped from threads: replicated host integrators bit-identical on every rank, observations and populations
    equal to the unsplit run to round-off."""
    import test_slabs
    P, Wl = g.BC_PERIODIC, g.BC_WALL
    bad, ran = [], 0
    for seed in range(24):
        rng = np.random.default_rng(seed)
        n_ranks, h = int(rng.integers(2, 4)), int(rng.integers(14, 22))
        nx, ny, nz = int(rng.integers(16, 24)), int(rng.integers(14, 20)), h * n_ranks
        kw = dict(nx=nx, ny=ny, nz=nz, tau=float(rng.uniform(0.7, 1.0)), collision=int(rng.integers(0, 2)),
                  bc=[Wl] * 4 + [P] * 2 if rng.random() < 0.5 else [P] * 6, max_markers=4000, max_links=8)
        if rng.random() < 0.5:
            kw["split_min_cells"] = 1
        whole = g.Sim(backend=emu, **kw)
        parts = [g.Sim(backend=emu, n_ranks=n_ranks, rank=r, **kw) for r in range(n_ranks)]
        handles = [s.peer_export() for s in parts]
        for s in parts:
            s.peer_connect_all(handles)
        links = tuple((float(rng.uniform(5, 8)), float(rng.uniform(1.5, 2.5))) for _ in range(int(rng.integers(2, 5))))
        d = util.fish_desc(g, root=(nx / 2, ny / 2, float(rng.uniform(0, nz))), links=links, free=int(rng.random() < 0.7),
                           heading=float(rng.uniform(-0.4, 0.4)))
        for s in [whole] + parts:
            s.add_fish(d)
        worst, ok = 0.0, True
        for _ in range(4):
            act, k = rng.uniform(-1, 1, whole.action_size()).astype(np.float32), int(rng.integers(1, 7))
            whole.set_action(act)
            whole.step(k)
            for s in parts:
                s.set_action(act)
            test_slabs._run_threads([lambda s=s: s.step(k) for s in parts])
            ow, r0 = whole.get_obs(), whole.get_fields(f64=True)[0]
            if not (np.isfinite(ow).all() and np.isfinite(r0).all() and 0.85 < r0.min() and r0.max() < 1.15):
                ok = False
                break
            worst = max([worst] + [float(np.abs(s.get_obs() - ow).max()) for s in parts])
            assert all(np.array_equal(parts[0].get_obs(), s.get_obs()) for s in parts), seed
        if ok:
            ran += 1
            f, fs = whole.get_populations(), np.concatenate([s.get_populations() for s in parts], axis=1)
            if worst > 1e-4 or np.abs(f - fs).max() > 1e-6:
                bad.append((seed, worst, float(np.abs(f - fs).max()), kw))
        for s in [whole] + parts:
            s.close()
    assert not bad, bad[:3]
    assert ran >= 16


@pytest.mark.gpu
def test_random_api_sequences_cuda_vs_oracle(g, cuda):
    """The same call sequences through the real library: CUDA graphs captured and replayed under changing marker sets,
    obstacle masks, flags of the random case, resets and read-outs in any order."""
    bad = [(seed, w) for seed in range(3000, 3120) for w in [run_api_sequence_case(g, cuda, seed)] if w is not None and w > 2e-5]
    assert not bad, bad[:5]
