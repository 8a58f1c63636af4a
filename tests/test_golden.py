"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the fp64 oracle).

CPU: the oracle still reproduces them to round-off (pins the checker against drift) and the emulated kernel bodies
match them.  GPU: the CUDA library matches them without needing the oracle at all."""
import os

import numpy as np
import pytest

import util

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["bgk_periodic", "mrt_force", "bgk_ywall_moving", "mrt_all_walls_lid", "mrt_inlet_outlet_ywalls", "mrt_ragged", "mrt_xwalls_moving",
         "mrt_xwalls_moving_nx256", "mrt_inlet_outlet_ywalls_nx128", "mrt_xy_walls_nx64"]
# fixtures on rows where the two-cell kernels are the default: (case of util.parity_cases, shape)
WIDE = {"mrt_xwalls_moving_nx256": ("mrt_xwalls_moving", dict(nx=256, ny=6, nz=8)),
        "mrt_inlet_outlet_ywalls_nx128": ("mrt_inlet_outlet_ywalls", dict(nx=128, ny=7, nz=8)),
        "mrt_xy_walls_nx64": ("mrt_xy_walls", dict(nx=64, ny=9, nz=6))}


def run_case(g, backend, name):
    cases = util.parity_cases(g)
    kw = dict(cases[WIDE[name][0]], **WIDE[name][1]) if name in WIDE else cases[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    s = g.Sim(backend=backend, **kw)
    rho, u = util.smooth_fields(s.shape)
    s.set_fields(rho, u)
    s.step(int(gold["steps"]))
    r, v = s.get_fields(f64=True)
    s.close()
    return util.rel_l2(v, gold["u"]), util.rel_l2(r, gold["rho"])


def run_ib(g, backend, fixture="ib_sphere"):
    P, Wl, IN, OUT = g.BC_PERIODIC, g.BC_WALL, g.BC_INLET, g.BC_OUTLET
    gold = np.load(os.path.join(GOLD, fixture + ".npz"))
    kw = dict(nx=20, ny=18, nz=24, tau=0.8, collision=g.MRT, max_markers=512, max_links=1, bc=[P, P, Wl, Wl, IN, OUT], inlet_u=[0, 0, 0.05],
              ib_iterations=int(gold["passes"]) if "passes" in gold.files else 1)
    s = g.Sim(backend=backend, **kw)
    X = util.sphere_markers((10.3, 9.1, 8.2), 4.0, 200)
    s.set_markers(X, np.zeros_like(X), np.ones(200, np.float32), np.zeros(200, np.int32))
    s.set_link_origins([[10.3, 9.1, 8.2]])
    u = np.zeros((3,) + s.shape)
    u[2] = 0.05
    s.set_fields(np.ones(s.shape), u)
    s.step(int(gold["steps"]))
    r, v = s.get_fields(f64=True)
    base, owner = s.get_index_map()
    out = dict(u=util.rel_l2(v, gold["u"]), rho=util.rel_l2(r, gold["rho"]), base=np.array_equal(base, gold["base"]),
               owner=np.array_equal(owner, gold["owner"]), wrench=float(np.abs(s.get_link_wrenches() - gold["wrench"]).max() / np.abs(gold["wrench"]).max()),
               Fm=util.rel_l2(s.get_marker_forces(), gold["Fm"]))
    s.close()
    return out


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(g, name):
    eu, er = run_case(g, "oracle", name)
    assert eu < 1e-12 and er < 1e-13


@pytest.mark.parametrize("name", CASES)
def test_emulated_kernels_match_golden(g, emu, name):
    eu, er = run_case(g, emu, name)
    assert eu <= 1e-5 and er <= 1e-5


@pytest.mark.parametrize("fixture", ["ib_sphere", "ib_sphere_mdf3"])
def test_ib_golden_oracle_and_emulation(g, emu, fixture):
    o = run_ib(g, "oracle", fixture)
    assert o["u"] < 1e-12 and o["base"] and o["owner"] and o["wrench"] < 1e-6 and o["Fm"] < 1e-6
    e = run_ib(g, emu, fixture)
    assert e["u"] <= 1e-5 and e["rho"] <= 1e-5 and e["base"] and e["owner"] and e["wrench"] <= 1e-4 and e["Fm"] <= 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_golden(g, cuda, name):
    eu, er = run_case(g, cuda, name)
    assert eu <= 1e-5 and er <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", ["ib_sphere", "ib_sphere_mdf3"])
def test_cuda_ib_matches_golden(g, cuda, fixture):
    e = run_ib(g, cuda, fixture)
    assert e["u"] <= 1e-5 and e["rho"] <= 1e-5 and e["base"] and e["owner"] and e["wrench"] <= 1e-4 and e["Fm"] <= 1e-4


def test_poiseuille_steady_state_fixture_is_a_fixed_point(g, emu):
    """The converged NY = 128 Poiseuille column (oracle, wall-exact MRT rates) tiled into a thin slab: both the oracle
    and the emulated fp32 kernels must keep the analytic parabola (the GPU test does the same at 256x128x128)."""
    P, Wl = g.BC_PERIODIC, g.BC_WALL
    gold = np.load(os.path.join(GOLD, "poiseuille_ny128.npz"))
    gf = float(gold["gf"])
    kw = dict(nx=6, ny=128, nz=3, tau=float(gold["tau"]), collision=g.MRT, bc=[P, P, Wl, Wl, P, P], body_force=[0, 0, gf],
              mrt_rates=list(gold["rates"]))
    for backend, tol in (("oracle", 3e-6), (emu, 1e-5)):      # the fixture itself went through a float32 interface
        s = g.Sim(backend=backend, **kw)
        s.set_populations(np.ascontiguousarray(np.broadcast_to(gold["f"][:, None, :, None], (19,) + s.shape), dtype=np.float32))
        s.step(501)
        _, uu = s.get_fields(f64=True)
        assert util.rel_l2(uu[2].mean(axis=(0, 2)) + gf / 2, gold["analytic"]) <= tol
        s.close()
