"""The oracle against analytic / classical results (BASELINE.json north_star (1)).
The reference has no tests to mirror, so these pin the checker itself."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import util


@pytest.mark.parametrize("coll,plane", [("bgk", "xy"), ("bgk", "yz"), ("mrt", "xz")])
def test_taylor_green_decay_rate(g, coll, plane):
    n, tau = 32, 0.8
    s = g.Sim(backend="oracle", nx=n, ny=n, nz=n, tau=tau, collision=g.MRT if coll == "mrt" else g.BGK)
    rho, u = util.taylor_green(n, plane)
    s.set_fields(rho, u)
    t, amp = [], []
    for it in range(0, 400, 50):
        _, uu = s.get_fields(f64=True)
        t.append(it)
        amp.append(np.sqrt((uu ** 2).mean()))
        s.step(50)
    rate = -np.polyfit(t, np.log(amp), 1)[0]
    nu, k = (tau - 0.5) / 3, 2 * np.pi / n
    assert abs(rate / (2 * nu * k * k) - 1) < 5e-3      # 0.2-0.3 % at 32^3
    s.close()


def _poiseuille(g, coll, tau, NY, magic, axis):
    P, Wl = g.BC_PERIODIC, g.BC_WALL
    gf, nu = 1e-6, (tau - 0.5) / 3
    kw = dict(nx=4, ny=NY, nz=4, tau=tau, collision=coll, bc=[P, P, Wl, Wl, P, P],
              body_force=[gf, 0, 0] if axis == 0 else [0, 0, gf])
    if magic:
        sn = 1 / tau
        sq = 8 * (2 - sn) / (8 - sn)
        kw["mrt_rates"] = [0, 1.19, 1.4, 0, sq, 0, sq, 0, sq, sn, 1.4, sn, 1.4, sn, sn, sn, sq, sq, sq]
    s = g.Sim(backend="oracle", **kw)
    y = np.arange(NY)
    ana = gf / (2 * nu) * (y + 0.5) * (NY - 0.5 - y)
    u = np.zeros((3, 4, NY, 4))
    u[axis] = (ana - gf / 2)[None, :, None]      # start from the analytic profile, assert it is the fixed point
    s.set_fields(np.ones((4, NY, 4)), u)
    s.step(int(3 * NY * NY / nu))
    _, uu = s.get_fields(f64=True)
    prof = uu[axis][0, :, 0] + gf / 2            # physical velocity includes the half force (Guo)
    s.close()
    return util.rel_l2(prof, ana)


def test_poiseuille_exact_with_magic_rates(g):
    assert _poiseuille(g, g.MRT, 0.8, 16, True, 2) < 1e-9
    assert _poiseuille(g, g.MRT, 1.2, 16, True, 0) < 1e-9


def test_poiseuille_wall_slip_of_default_rates_is_second_order(g):
    assert _poiseuille(g, g.BGK, 0.8, 16, False, 2) < 2 * 4e-3
    assert _poiseuille(g, g.MRT, 0.8, 16, False, 2) < 2 * 4.7e-3


def test_mrt_all_rates_omega_equals_bgk(g):
    kw = dict(nx=10, ny=8, nz=6, tau=0.8, body_force=[1e-4, 0, -1e-4])
    a = g.Sim(backend="oracle", **kw)
    b = g.Sim(backend="oracle", collision=g.MRT, mrt_rates=[1.25] * 19, **kw)
    rho, u = util.smooth_fields(a.shape)
    for s in (a, b):
        s.set_fields(rho, u)
        s.step(20)
    assert util.rel_l2(b.get_fields(f64=True)[1], a.get_fields(f64=True)[1]) < 1e-12


@settings(max_examples=8, deadline=None)
@given(st.integers(0, 1000), st.sampled_from(["bgk", "mrt"]), st.floats(0.55, 1.5))
def test_periodic_box_conserves_mass_and_momentum(g, seed, coll, tau):
    s = g.Sim(backend="oracle", nx=9, ny=7, nz=6, tau=tau, collision=g.MRT if coll == "mrt" else g.BGK)
    rng = np.random.default_rng(seed)
    rho = 1 + 0.02 * rng.standard_normal(s.shape)
    u = 0.03 * rng.standard_normal((3,) + s.shape)
    s.set_fields(rho, u)
    r0, u0 = s.get_fields(f64=True)
    s.step(13)
    r1, u1 = s.get_fields(f64=True)
    assert abs(r1.sum() - r0.sum()) < 1e-11 * r0.sum()
    assert np.abs((r1 * u1).sum(axis=(1, 2, 3)) - (r0 * u0).sum(axis=(1, 2, 3))).max() < 1e-11
    s.close()


def test_guo_force_adds_exactly_F_per_step(g):
    F = np.array([1e-4, -2e-4, 3e-4])
    s = g.Sim(backend="oracle", nx=6, ny=6, nz=6, tau=0.7, collision=g.MRT, body_force=list(F))
    s.set_fields(np.ones(s.shape), np.zeros((3,) + s.shape))
    s.step(10)
    r, u = s.get_fields(f64=True)
    assert np.allclose((r * u).mean(axis=(1, 2, 3)), 10 * F, rtol=1e-10)


def test_spread_force_equals_marker_force_and_wrench_signs(g):
    kw = dict(nx=24, ny=24, nz=24, tau=0.8, max_markers=1000, max_links=2)
    s = g.Sim(backend="oracle", **kw)
    X = util.sphere_markers((12.2, 11.7, 12.4), 5.0, 300)
    U = np.zeros_like(X)
    dV = np.full(300, 4 * np.pi * 25 / 300, np.float32)
    s.set_markers(X, U, dV, np.zeros(300, np.int32))
    s.set_link_origins([[12.2, 11.7, 12.4]])
    u = np.zeros((3,) + s.shape)
    u[2] = 0.05
    s.set_fields(np.ones(s.shape), u)
    s.step(1)
    Fm = (s.get_marker_forces().astype(np.float64) * dV[:, None]).sum(0)
    Fg = s.get_force_field().astype(np.float64).sum(axis=(1, 2, 3))
    assert np.allclose(Fg, Fm, rtol=1e-5, atol=1e-9)          # sum of delta weights is 1 per marker
    w = s.get_link_wrenches()[0]
    assert np.allclose(w[:3], -Fm, rtol=1e-6)                 # force ON the body
    assert w[2] > 0                                           # a body at rest in a +z stream is dragged along +z
    base, owner = s.get_index_map()
    assert np.array_equal(base, np.floor(X).astype(np.int32) - 1) and (owner == 0).all()
    s.close()


def test_sphere_drag_reduced_size(g):
    """Fixed IB sphere, Re = 20, reduced 48x48x96 channel (BASELINE.json configs[1] at 1/8 scale so the CPU suite
    stays short).  Blockage (D/H = 0.25) and the diffuse interface raise Cd above the unbounded
    Schiller-Naumann value 2.61; the full-size run on the GPU narrows this."""
    P, IN, OUT = g.BC_PERIODIC, g.BC_INLET, g.BC_OUTLET
    D, Uin, Re = 12.0, 0.05, 20.0
    nu = Uin * D / Re
    kw = dict(nx=48, ny=48, nz=96, tau=3 * nu + 0.5, collision=g.MRT, bc=[P, P, P, P, IN, OUT], inlet_u=[0, 0, Uin],
              max_markers=1000, max_links=1)
    s = g.Sim(backend="oracle", **kw)
    n = int(round(np.pi * D * D))
    X = util.sphere_markers((24.3, 24.2, 30.1), D / 2, n)
    s.set_markers(X, np.zeros_like(X), np.full(n, np.pi * D * D / n, np.float32), np.zeros(n, np.int32))
    u = np.zeros((3,) + s.shape)
    u[2] = Uin
    s.set_fields(np.ones(s.shape), u)
    s.step(1500)
    Fz = s.get_link_wrenches()[0, 2]
    Deff = D + 1.0   # the 4-point delta thickens the sphere by about half a cell on each side
    cd = Fz / (0.5 * Uin ** 2 * np.pi * Deff ** 2 / 4)
    sn = 24 / Re * (1 + 0.15 * Re ** 0.687)
    assert 0.9 * sn < cd < 1.6 * sn, (cd, sn)
    s.close()


def test_sphere_drag_by_momentum_exchange_on_a_staircase_sphere(g):
    """"... and via momentum-exchange on a staircase sphere as a cross-check".  The same channel and Re as
    the immersed-boundary test above with the sphere carved out of obstacle cells (half-way bounce-back), drag from
    fg_get_solid_force.  Measured 3.24 here against 3.5 for the immersed sphere (whose 4-point kernel makes it about half a
    cell larger; `tools/drag_crosscheck.py`, profiles/r2_cpu_drag_crosscheck.txt) and 2.61 for an unbounded fluid:
    the periodic array with D/L = 0.25 confines the flow.  Centred sphere: no lateral force, no torque."""
    P, IN, OUT = g.BC_PERIODIC, g.BC_INLET, g.BC_OUTLET
    D, Uin, Re = 12.0, 0.05, 20.0
    nu = Uin * D / Re
    s = g.Sim(backend="oracle", nx=48, ny=48, nz=96, tau=3 * nu + 0.5, collision=g.MRT, bc=[P, P, P, P, IN, OUT], inlet_u=[0, 0, Uin])
    c = (23.5, 23.5, 30.5)
    z, y, x = np.meshgrid(np.arange(96), np.arange(48), np.arange(48), indexing="ij")
    solid = (((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) <= (D / 2) ** 2).astype(np.uint8)
    assert abs(int(solid.any(axis=0).sum()) - np.pi * D * D / 4) < 3          # frontal area of the staircase
    s.set_solid(solid)
    u = np.zeros((3,) + s.shape)
    u[2] = Uin
    s.set_fields(np.ones(s.shape), u)
    s.step(1000)
    F = s.get_solid_force(c)
    cd = F[2] / (0.5 * Uin ** 2 * np.pi * D ** 2 / 4)
    sn = 24 / Re * (1 + 0.15 * Re ** 0.687)
    assert 1.05 * sn < cd < 1.45 * sn, (cd, sn)
    assert np.abs(F[:2]).max() < 1e-9 * F[2] and np.abs(F[3:]).max() < 1e-9 * F[2] * D
    s.close()


def _hasimoto(c):
    """Stokes drag of a simple cubic array of spheres, volume fraction c: F = 6 pi mu a U K with U the mean velocity over the whole
    cell and F the total force on a sphere INCLUDING the mean pressure gradient's buoyancy (H. Hasimoto, J. Fluid Mech. 5 (1959) 317,
    series extended by Sangani & Acrivos, Int. J. Multiphase Flow 8 (1982) 343; K(0.027) = 2.008 and K(0.064) = 2.810 are the tabulated
    values of Zick & Homsy, J. Fluid Mech. 115 (1982) 13).  A published analytic result — nothing of it comes from this repository."""
    return 1.0 / (1 - 1.7601 * c ** (1 / 3) + c - 1.5593 * c ** 2 + 3.9799 * c ** (8 / 3) - 3.0734 * c ** (10 / 3))


_MAGIC = lambda tau: [0, 1.19, 1.4, 0, 8 * (2 - 1 / tau) / (8 - 1 / tau), 0, 8 * (2 - 1 / tau) / (8 - 1 / tau), 0, 8 * (2 - 1 / tau) / (8 - 1 / tau),   # noqa: E731
                      1 / tau, 1.4, 1 / tau, 1.4, 1 / tau, 1 / tau, 1 / tau] + [8 * (2 - 1 / tau) / (8 - 1 / tau)] * 3


def test_hasimoto_series_reproduces_the_tabulated_values():
    assert abs(_hasimoto(0.027) - 2.008) < 2e-3 and abs(_hasimoto(0.064) - 2.810) < 2e-3


@pytest.mark.parametrize("tau", [1.0, 0.8])
def test_stokes_drag_of_a_periodic_array_of_spheres_matches_hasimoto(g, tau):
    """A sphere of obstacle cells in a fully periodic cube IS a simple cubic array.  A uniform body force g on the fluid drives the
    flow (a mean pressure gradient whose buoyancy g V_sphere the body force does not exert on the obstacle, hence F = g L^3 in
    Hasimoto's sense while the bounce-back links carry g x fluid cells — checked with fg_get_solid_force).  With the sphere's
    volume-equivalent radius the drag coefficient K comes out 1.1 % above the series at c = 0.066 (24^3, a = 6; 1.3 % at 32^3,
    a = 8; the same for both viscosities with the wall-exact rates): bounce-back, forcing, viscosity and the momentum-exchange read-out
    in one number that is pinned from outside."""
    L, a, gf = 24, 6.0, 1e-6
    nu = (tau - 0.5) / 3
    s = g.Sim(backend="oracle", nx=L, ny=L, nz=L, tau=tau, collision=g.MRT, body_force=[0, 0, gf], mrt_rates=_MAGIC(tau))
    c0 = L / 2 - 0.5
    z, y, x = np.meshgrid(np.arange(L), np.arange(L), np.arange(L), indexing="ij")
    solid = (((x - c0) ** 2 + (y - c0) ** 2 + (z - c0) ** 2) <= a * a).astype(np.uint8)
    s.set_solid(solid)
    s.step(6000)            # the mean flow relaxes with U / g = 250 (tau 1.0) and 420 (tau 0.8) steps: e^-14 at least
    r, v = s.get_fields(f64=True)
    fluid = solid == 0
    U = ((v[2] + gf / 2 / r) * fluid).sum() / L ** 3           # mean over the whole cell; Guo: the physical velocity carries half the force
    F = s.get_solid_force()
    assert abs(F[2] / (gf * fluid.sum()) - 1) < 1e-5 and np.abs(F[:2]).max() < 1e-12
    vol = float(solid.sum())
    a_eq = (3 * vol / (4 * np.pi)) ** (1 / 3)
    K = gf * L ** 3 / (6 * np.pi * nu * a_eq * U)
    assert abs(K / _hasimoto(vol / L ** 3) - 1) < 0.02, (K, _hasimoto(vol / L ** 3))
    s.close()


@pytest.mark.parametrize("passes", [1, 3])
def test_hydrodynamic_radius_of_an_immersed_sphere_from_hasimoto(g, passes):
    """The same array with the sphere as an immersed boundary of markers at radius a (the fluid inside is held by the markers, so
    they carry the whole g L^3).  Solving Hasimoto's relation for the radius that explains the measured mean velocity gives
    a + 0.47 (24^3, a = 6) and a + 0.45 (32^3, a = 8) with one direct-forcing pass — the "half a cell" by which the 4-point kernel
    thickens a body, used as D + 1 in the drag tests — and a + 0.57 / a + 0.54 with three passes (no-slip enforced more fully)."""
    from scipy.optimize import brentq
    L, a, gf, tau = 24, 6.0, 1e-6, 1.0
    nu = (tau - 0.5) / 3
    n = int(round(4 * np.pi * a * a))
    s = g.Sim(backend="oracle", nx=L, ny=L, nz=L, tau=tau, collision=g.MRT, body_force=[0, 0, gf], max_markers=n, max_links=1,
              ib_iterations=passes, mrt_rates=_MAGIC(tau))
    c0 = L / 2 - 0.5
    X = util.sphere_markers((c0, c0, c0), a, n)
    s.set_markers(X, np.zeros_like(X), np.full(n, 4 * np.pi * a * a / n, np.float32), np.zeros(n, np.int32))
    s.step(2500)
    r, v = s.get_fields(f64=True)
    Fz = s.get_force_field().astype(np.float64)[2]
    U = (v[2] + (gf + Fz) / 2 / r).sum() / L ** 3
    assert abs(s.get_link_wrenches()[0][2] / (gf * L ** 3) - 1) < 1e-3          # steady: the markers carry the whole body force
    a_h = brentq(lambda ah: 6 * np.pi * nu * ah * U * _hasimoto(4 / 3 * np.pi * ah ** 3 / L ** 3) - gf * L ** 3, 0.5 * a, 1.5 * a)
    lo, hi = (0.40, 0.52) if passes == 1 else (0.50, 0.64)
    assert a + lo < a_h < a + hi, a_h
    s.close()


@pytest.mark.parametrize("coll,tau,tol", [("bgk", 0.8, 2.5e-3), ("mrt", 0.8, 4e-4), ("mrt", 0.6, 2.5e-3)])
def test_stokes_first_problem_impulsively_started_wall(g, coll, tau, tol):
    """Fluid at rest, the y-low wall moves at U from t = 0 on: u(y, t) = U erfc(y / (2 sqrt(nu t))), the classical similarity
    solution (the far wall is 20 diffusion lengths away at the last checkpoint).  Unsteady viscous diffusion and the moving-wall
    term of half-way bounce-back against an analytic TRANSIENT; the error falls with time as the start-up layer is resolved."""
    from scipy.special import erfc
    P, Wl = g.BC_PERIODIC, g.BC_WALL
    NY, Uw = 96, 0.01
    nu = (tau - 0.5) / 3
    kw = dict(mrt_rates=_MAGIC(tau)) if coll == "mrt" else {}
    s = g.Sim(backend="oracle", nx=3, ny=NY, nz=3, tau=tau, collision=g.MRT if coll == "mrt" else g.BGK, bc=[P, P, Wl, Wl, P, P],
              wall_u={g._abi.YLO: [Uw, 0, 0]}, **kw)
    y = np.arange(NY) + 0.5                    # the wall sits half a cell below the first node
    done, errs = 0, []
    for t in (100, 400, 1000):
        s.step(t - done)
        done = t
        _, u = s.get_fields(f64=True)
        errs.append(float(np.abs(u[0][0, :, 0] - Uw * erfc(y / (2 * np.sqrt(nu * t)))).max() / Uw))
        assert np.abs(u[1]).max() < 1e-5 * Uw and np.abs(u[2]).max() < 1e-12      # wall-normal: O(Ma^2 U) compressibility only (measured 2e-6 U)
    assert errs[0] < tol and errs[1] < errs[0] / 3 and errs[2] < errs[1] / 2, errs
    s.close()


@pytest.mark.parametrize("coll", ["bgk", "mrt"])
def test_sound_speed_of_a_standing_wave(g, coll):
    """A standing density wave of wavelength 64 oscillates with period lambda / c_s = 64 sqrt(3) = 110.85 steps: the isothermal
    equation of state p = rho / 3 of the lattice (measured 110.88 / 110.87; the damping shifts the period by 1.5e-4)."""
    N = 64
    s = g.Sim(backend="oracle", nx=N, ny=2, nz=2, tau=0.55, collision=g.MRT if coll == "mrt" else g.BGK)
    mode = np.cos(2 * np.pi * np.arange(N) / N)
    s.set_fields(np.ones(s.shape) + 1e-4 * mode[None, None, :], np.zeros((3,) + s.shape))
    amp = []
    for _ in range(400):
        r, _u = s.get_fields(f64=True)
        amp.append(float(((r - 1)[0, 0, :] * mode).sum() * 2 / N))
        s.step(1)
    zero = [i + amp[i] / (amp[i] - amp[i + 1]) for i in range(len(amp) - 1) if amp[i] * amp[i + 1] < 0]
    assert len(zero) >= 6
    c_s = N / (2 * np.mean(np.diff(zero)))
    assert abs(c_s * np.sqrt(3) - 1) < 5e-4, c_s
    s.close()


@pytest.mark.parametrize("NX,NY,tau,tol", [(16, 12, 0.8, 2e-3), (24, 16, 1.0, 1e-3)])
def test_rectangular_duct_flow_matches_the_series_solution(g, NX, NY, tau, tol):
    """Walls on x AND y, body force along z: the classical Fourier-series solution of Poiseuille flow in a rectangular duct
    (e.g. White, Viscous Fluid Flow, eq. 3-48), walls half a cell outside the first / last nodes.  Edges and corners of the
    half-way bounce-back (links that cross two wall faces) against an analytic field: 1.2e-3 at 16 x 12, 5e-4 at 24 x 16 (second order)."""
    P, Wl = g.BC_PERIODIC, g.BC_WALL
    gf, nu = 1e-6, (tau - 0.5) / 3
    s = g.Sim(backend="oracle", nx=NX, ny=NY, nz=3, tau=tau, collision=g.MRT, bc=[Wl, Wl, Wl, Wl, P, P], body_force=[0, 0, gf], mrt_rates=_MAGIC(tau))
    s.step(int(1.5 * max(NX, NY) ** 2 / nu))
    _, u = s.get_fields(f64=True)
    a, b = NX / 2, NY / 2
    X, Y = np.meshgrid(np.arange(NX) + 0.5 - a, np.arange(NY) + 0.5 - b, indexing="xy")
    ana = gf / (2 * nu) * (a * a - X * X)
    for n in range(200):
        k = (2 * n + 1) * np.pi / (2 * a)
        ana -= gf / nu * 2 * (-1) ** n / (a * k ** 3) * np.cos(k * X) * np.cosh(k * Y) / np.cosh(k * b)
    num = u[2][0] + gf / 2
    assert util.rel_l2(num, ana) < tol and abs(num.sum() / ana.sum() - 1) < tol, (util.rel_l2(num, ana), num.sum() / ana.sum())
    assert np.abs(u[2] - u[2][:1]).max() < 1e-15 and np.abs(u[0]).max() < 1e-5 * num.max() and np.abs(u[1]).max() < 1e-5 * num.max()
    s.close()


@pytest.mark.parametrize("passes", [1, 3])
def test_taylor_couette_between_immersed_cylinders(g, passes):
    """Circular Couette flow: an inner cylinder of markers rotating at Omega inside an outer one at rest (axis along z, 4 periodic
    planes).  The analytic steady flow is u_theta = A r + B / r and the torque per length on either cylinder is 4 pi mu B.  Checked:
    the profile in the gap has that form (residual 5e-4 of the wall speed); the radii at which it meets the two wall speeds lie
    0.40 cells inside the fluid from the marker radii (0.45 with three passes; 0.39 / 0.44 at 72^2 with R = 16 / 28) — the same
    thickening the Hasimoto test finds for a sphere; and the TORQUE of the link wrench on the inner cylinder is 4 pi mu B of that
    profile to 1e-3: the first check of the torque read-out and of moving markers (U_d = Omega x r) against an analytic flow."""
    N, nz, R1, R2, tau = 48, 4, 10.0, 19.0, 1.0
    nu, Om, c = (tau - 0.5) / 3, 0.01 / R1, N / 2 - 0.5

    def ring(R):
        n = int(round(2 * np.pi * R))
        th = 2 * np.pi * (np.arange(n) + 0.5) / n
        return np.concatenate([np.stack([c + R * np.cos(th), c + R * np.sin(th), np.full(n, float(zz))], 1) for zz in range(nz)]).astype(np.float32), 2 * np.pi * R / n
    (X1, ds1), (X2, ds2) = ring(R1), ring(R2)
    X = np.concatenate([X1, X2])
    U = np.zeros_like(X)
    U[:len(X1), 0], U[:len(X1), 1] = -Om * (X1[:, 1] - c), Om * (X1[:, 0] - c)
    dV = np.concatenate([np.full(len(X1), ds1), np.full(len(X2), ds2)]).astype(np.float32)
    link = np.concatenate([np.zeros(len(X1)), np.ones(len(X2))]).astype(np.int32)
    s = g.Sim(backend="oracle", nx=N, ny=N, nz=nz, tau=tau, collision=g.MRT, max_markers=len(X), max_links=2, ib_iterations=passes, mrt_rates=_MAGIC(tau))
    s.set_markers(X, U, dV, link)
    s.set_link_origins([[c, c, 0], [c, c, 0]])
    s.step(4000)                                   # gap^2 / nu = 490 steps
    W = s.get_link_wrenches()
    r, v = s.get_fields(f64=True)
    F3 = s.get_force_field().astype(np.float64)
    ux, uy = (v[0] + F3[0] / 2 / r)[0], (v[1] + F3[1] / 2 / r)[0]          # physical velocity (Guo: half the force)
    yy, xx = np.meshgrid(np.arange(N) - c, np.arange(N) - c, indexing="ij")
    rr = np.hypot(xx, yy)
    ut = (-yy * ux + xx * uy) / np.maximum(rr, 1e-9)
    m = (rr > R1 + 2.5) & (rr < R2 - 2.5)                                  # clear of the smeared shells
    basis = np.stack([rr[m], 1 / rr[m]], 1)
    (A, B), *_ = np.linalg.lstsq(basis, ut[m], rcond=None)
    assert np.abs(basis @ [A, B] - ut[m]).max() < 1e-3 * Om * R1
    r1h, r2h = np.sqrt(B / (Om - A)), np.sqrt(-B / A)
    lo, hi = (0.34, 0.46) if passes == 1 else (0.39, 0.51)
    assert lo < r1h - R1 < hi and lo < R2 - r2h < hi and abs((r1h - R1) - (R2 - r2h)) < 0.02, (r1h, r2h)
    Tin, Tout = W[0][5] / nz, W[1][5] / nz
    assert Tin < 0 < Tout                                                  # the fluid brakes the rotor and drags the stator along
    assert abs(-Tin / (4 * np.pi * nu * B) - 1) < 1e-3, (Tin, 4 * np.pi * nu * B)
    assert np.abs(W[:, :2]).max() < 1e-3 * abs(W[0][5]) / R1                # no net in-plane force on either cylinder (torque / R = the shear force)
    s.close()


@pytest.mark.parametrize("coll,tau", [("bgk", 1.0), ("mrt", 0.7)])
def test_couette_linear_profile_is_exact_with_a_moving_wall(g, coll, tau):
    """Plane Couette flow between y walls, the upper one moving tangentially in x AND z: the linear profile is an exact
    solution of half-way bounce-back with the moving-wall term (walls half a cell outside the first / last node), for BGK
    (to round-off) and for MRT (to third order in the wall speed) — this pins the 6 w_i c_i.u_w correction and the wall location."""
    P, Wl = g.BC_PERIODIC, g.BC_WALL
    NY, Uw, Ww = 12, 0.04, -0.02
    s = g.Sim(backend="oracle", nx=4, ny=NY, nz=4, tau=tau, collision=g.MRT if coll == "mrt" else g.BGK,
              bc=[P, P, Wl, Wl, P, P], wall_u={g._abi.YHI: [Uw, 0, Ww]})
    nu = (tau - 0.5) / 3
    s.step(int(6 * NY * NY / nu))
    rho, u = s.get_fields(f64=True)
    lin = (np.arange(NY) + 0.5) / NY
    # BGK reproduces the line to round-off of the steady-state iteration; MRT with the default (non-magic) rates keeps an
    # O(u_w^3) kink of ~1e-7 in the two nodes next to the moving wall
    tol = 5e-9 if coll == "bgk" else 5e-7
    assert np.abs(u[0][0, :, 0] - Uw * lin).max() < tol
    assert np.abs(u[2][0, :, 0] - Ww * lin).max() < tol
    assert np.abs(u[1]).max() < 1e-12 and np.abs(rho - 1).max() < (1e-9 if coll == "bgk" else 1e-5)   # MRT: O(u_w^2) density layer
    s.close()


def test_taylor_green_decay_error_is_second_order_in_resolution(g):
    """Fitted decay rate against 2 nu k^2 at 16^3 and 32^3 (same tau): the error falls by ~4 when the resolution doubles."""
    err = {}
    for n in (16, 32):
        tau = 0.8
        s = g.Sim(backend="oracle", nx=n, ny=n, nz=2, tau=tau)
        k = 2 * np.pi / n
        z, y, x = np.meshgrid(np.arange(2), np.arange(n), np.arange(n), indexing="ij")
        u = np.zeros((3, 2, n, n))
        U0 = 0.01
        u[0] = U0 * np.sin(k * x) * np.cos(k * y)
        u[1] = -U0 * np.cos(k * x) * np.sin(k * y)
        rho = 1 + 3 * (U0 ** 2 / 4) * (np.cos(2 * k * x) + np.cos(2 * k * y))
        s.set_fields(rho, u)
        t, amp = [], []
        steps = n * n // 8
        for it in range(8):
            _, uu = s.get_fields(f64=True)
            t.append(it * steps)
            amp.append(np.sqrt((uu ** 2).mean()))
            s.step(steps)
        rate = -np.polyfit(t, np.log(amp), 1)[0]
        err[n] = abs(rate / (2 * (tau - 0.5) / 3 * k * k) - 1)
        s.close()
    assert 3.0 < err[16] / err[32] < 5.5, err


def test_lid_driven_cavity_matches_ghia_re100(g):
    """Published benchmark numbers (Ghia, Ghia & Shin 1982, Re = 100) pin walls on four sides, the moving-wall term and the
    non-linear advection together.  48^2 nodes, 10 000 steps (the flow is steady to 1e-5 by then): both centre-line
    profiles within 1.2 % of the lid speed (measured 0.72 % / 0.65 %; 0.59 % / 0.73 % at 64^2)."""
    du, dv, w = util.cavity_vs_ghia(g, "oracle", 48, 10000)
    assert du < 0.012 and dv < 0.012, (du, dv)
    assert w < 1e-12      # nothing drives the periodic direction
