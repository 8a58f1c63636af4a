"""A second, independently written restatement of the coupled step (numpy, moment-space MRT, roll streaming) against the oracle.

The reference ships nothing to diff against (/root/reference/README.md:1-15 is the whole checkout), so the oracle's parity
is pinned by analytic results (test_oracle_physics.py).  Analytic cases exercise one mechanism at a time; this file pins
the whole update rule — collide, Guo source, streaming, half-way bounce-back with a moving wall, inlet / outlet planes,
direct-forcing immersed boundary, link wrench — against an implementation that shares nothing with oracle/fg_oracle.cpp
but SURVEY.md Appendix A:

  oracle                                  here
  --------------------------------------  ---------------------------------------------------------------
  f* = f - A (f - f_eq) + B Phi           relaxation in MOMENT space, m_eq from the closed forms of A3,
    with A = M^-1 S M built numerically     M Phi from the closed form of A4, f* = M^T diag(1/|row|^2) m*
  per-cell pull() with a rule per face    whole-array np.roll, then the wall / inlet / outlet planes are patched
  marker loops                            one dense 4x4x4 weight block per marker (np.einsum / np.add.at)

Populations agree to ~1e-15 per step, so both have the same update rule; the oracle is then the checker for the CUDA path.
"""
import numpy as np
import pytest

CX = np.array([0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0])
CY = np.array([0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1])
CZ = np.array([0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1])
W = np.array([1 / 3] + [1 / 18] * 6 + [1 / 36] * 12)
OPP = np.array([0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15])


def moment_matrix():
    x, y, z = CX.astype(float), CY.astype(float), CZ.astype(float)
    c2 = x * x + y * y + z * z
    return np.stack([
        np.ones(19), 19 * c2 - 30, (21 * c2 * c2 - 53 * c2 + 24) / 2,
        x, (5 * c2 - 9) * x, y, (5 * c2 - 9) * y, z, (5 * c2 - 9) * z,
        3 * x * x - c2, (3 * c2 - 5) * (3 * x * x - c2), y * y - z * z, (3 * c2 - 5) * (y * y - z * z),
        x * y, y * z, x * z, (y * y - z * z) * x, (z * z - x * x) * y, (x * x - y * y) * z])


def equilibrium(rho, u):
    cu = CX[:, None, None, None] * u[0] + CY[:, None, None, None] * u[1] + CZ[:, None, None, None] * u[2]
    return W[:, None, None, None] * rho * (1 + 3 * cu + 4.5 * cu * cu - 1.5 * (u * u).sum(0))


def peskin(r):
    r = np.abs(r)
    inner = (3 - 2 * r + np.sqrt(np.maximum(1 + 4 * r - 4 * r * r, 0))) / 8
    outer = (5 - 2 * r - np.sqrt(np.maximum(-7 + 12 * r - 4 * r * r, 0))) / 8
    return np.where(r < 1, inner, np.where(r < 2, outer, 0.0))


class NumpyStep:
    """Fields are [.., nz, ny, nx]; x periodic; y periodic or walls (wall_u per face); z periodic or inlet(lo) / outlet(hi)."""

    def __init__(self, shape, tau, mrt, rates=None, g=(0, 0, 0), ywalls=False, wall_u_ylo=(0, 0, 0), wall_u_yhi=(0, 0, 0),
                 inlet_u=None):
        self.shape, self.tau, self.mrt, self.g = shape, tau, mrt, np.asarray(g, float)
        self.ywalls, self.uw = ywalls, (np.asarray(wall_u_ylo, float), np.asarray(wall_u_yhi, float))
        self.inlet_u = None if inlet_u is None else np.asarray(inlet_u, float)
        self.M = moment_matrix()
        self.norm2 = (self.M * self.M).sum(1)
        sn = 1 / tau
        s = np.array([0, 1.19, 1.4, 0, 1.2, 0, 1.2, 0, 1.2, sn, 1.4, sn, 1.4, sn, sn, sn, 1.98, 1.98, 1.98]) if rates is None else np.asarray(rates, float)
        self.s = s if mrt else np.full(19, sn)      # BGK: one rate on every row, the conserved ones included
        self.markers = None
        self.ib_passes = 1

    def init(self, rho, u):
        self.f = equilibrium(np.asarray(rho, float), np.asarray(u, float))

    def set_markers(self, X, U, dV, origin):
        self.markers = (np.asarray(X, np.float32).astype(float), np.asarray(U, np.float32).astype(float), np.asarray(dV, np.float32).astype(float),
                        np.asarray(origin, float))

    # ---- A6/A7 (2)-(5),(7): direct forcing from the unforced velocity of the arriving populations
    def ib_force(self):
        X, U, dV, origin = self.markers
        rho = self.f.sum(0)
        us = np.stack([(CX[:, None, None, None] * self.f).sum(0), (CY[:, None, None, None] * self.f).sum(0), (CZ[:, None, None, None] * self.f).sum(0)]) / rho
        F = np.zeros((3,) + self.shape)
        self.Fm = np.zeros_like(X)
        self.base = np.floor(X).astype(np.int32) - 1
        for k in range(len(X)):
            i0, j0, k0 = self.base[k]
            wx, wy, wz = (peskin(X[k, d] - (self.base[k, d] + np.arange(4))) for d in range(3))
            w3 = wz[:, None, None] * wy[None, :, None] * wx[None, None, :]
            block = (slice(None), slice(k0, k0 + 4), slice(j0, j0 + 4), slice(i0, i0 + 4))
            ustar = np.einsum("dzyx,zyx->d", us[block], w3)
            self.Fm[k] = 2.0 * (U[k] - ustar)
            F[block] += self.Fm[k][:, None, None, None] * (w3 * dV[k])
        if self.ib_passes > 1:
            # A7 (4) "n_iter > 1", multi-direct forcing, stated with matrices instead of the oracle's gather / spread loops:
            # D [markers x cells] holds the delta weights, so interpolation is D, spreading is D^T diag(dV), and with
            # b = 2 (U_d - U*) every further pass is F_m <- F_m + b - D (D^T (dV F_m))   (what the markers see is U* + D F(x) / 2)
            n, ncell = len(X), int(np.prod(self.shape))
            D = np.zeros((n, ncell))
            nz, ny, nx = self.shape
            for k in range(n):
                i0, j0, k0 = self.base[k]
                wx, wy, wz = (peskin(X[k, d] - (self.base[k, d] + np.arange(4))) for d in range(3))
                zz, yy, xx = np.meshgrid(k0 + np.arange(4), j0 + np.arange(4), i0 + np.arange(4), indexing="ij")
                D[k, ((zz * ny + yy) * nx + xx).ravel()] = (wz[:, None, None] * wy[None, :, None] * wx[None, None, :]).ravel()
            b = self.Fm.copy()
            for _ in range(self.ib_passes - 1):
                self.Fm = self.Fm + b - D @ (D.T @ (dV[:, None] * self.Fm))
            F = (D.T @ (dV[:, None] * self.Fm)).T.reshape((3,) + self.shape)
        r = X - origin
        f = -self.Fm * dV[:, None]
        self.wrench = np.concatenate([f.sum(0), np.cross(r, f).sum(0)])
        return F

    # ---- A3/A4: relax in moment space
    def collide(self, F):
        f = self.f
        m = np.tensordot(self.M, f, 1)
        rho = m[0]
        j = np.stack([m[3], m[5], m[7]])
        ru = j + 0.5 * F                       # rho u includes half the force (A4)
        u = ru / rho
        jj = (ru * ru).sum(0) / rho
        meq = np.zeros_like(m)
        meq[0] = rho
        meq[1] = -11 * rho + 19 * jj
        meq[2] = 3 * rho - 5.5 * jj
        meq[3], meq[5], meq[7] = ru
        meq[4], meq[6], meq[8] = -2 / 3 * ru
        meq[9] = (2 * ru[0] ** 2 - ru[1] ** 2 - ru[2] ** 2) / rho
        meq[10] = -0.5 * meq[9]
        meq[11] = (ru[1] ** 2 - ru[2] ** 2) / rho
        meq[12] = -0.5 * meq[11]
        meq[13] = ru[0] * ru[1] / rho
        meq[14] = ru[1] * ru[2] / rho
        meq[15] = ru[0] * ru[2] / rho
        uF = (u * F).sum(0)
        fxx = 2 * u[0] * F[0] - u[1] * F[1] - u[2] * F[2]
        fww = u[1] * F[1] - u[2] * F[2]
        z = np.zeros_like(rho)
        src = np.stack([z, 38 * uF, -11 * uF, F[0], -2 / 3 * F[0], F[1], -2 / 3 * F[1], F[2], -2 / 3 * F[2],
                        2 * fxx, -fxx, 2 * fww, -fww, u[0] * F[1] + u[1] * F[0], u[1] * F[2] + u[2] * F[1], u[0] * F[2] + u[2] * F[0], z, z, z])
        s = self.s[:, None, None, None]
        # momentum rows: j - s (j - (j + F/2)) + (1 - s/2) F = j + F for any rate s, as it must be
        mstar = m - s * (m - meq) + (1 - 0.5 * s) * src
        return np.tensordot((self.M / self.norm2[:, None]).T, mstar, 1)

    # ---- A5: streaming as a whole-array shift, then the faces
    def stream(self, fs):
        nz, ny, nx = self.shape
        f = np.empty_like(fs)
        for i in range(19):
            f[i] = np.roll(fs[i], (CZ[i], CY[i], CX[i]), axis=(0, 1, 2))
        if self.ywalls:
            for i in range(19):
                if CY[i] == 0:
                    continue
                row, uw = (0, self.uw[0]) if CY[i] > 0 else (ny - 1, self.uw[1])    # population i arrives from beyond that wall
                f[i][:, row, :] = fs[OPP[i]][:, row, :] + 6 * W[i] * (CX[i] * uw[0] + CY[i] * uw[1] + CZ[i] * uw[2])
        if self.inlet_u is not None:
            feq_in = equilibrium(np.ones((1, 1, 1)), self.inlet_u.reshape(3, 1, 1, 1))[:, 0, 0, 0]
            for i in range(19):
                wall_first = self.ywalls and CY[i] != 0      # rows where the same link also crosses a y wall keep the bounce-back value
                rows = slice(None) if not wall_first else (slice(1, ny) if CY[i] > 0 else slice(0, ny - 1))
                if CZ[i] > 0:      # enters through the low z face: equilibrium of the inlet state
                    f[i][0, rows, :] = feq_in[i]
                elif CZ[i] < 0:    # enters through the high z face: zero gradient, i.e. pulled from the clamped plane
                    src = np.roll(fs[i][nz - 1], (CY[i], CX[i]), axis=(0, 1))
                    f[i][nz - 1, rows, :] = src[rows, :]
        return f

    def step(self, n=1):
        for _ in range(n):
            F = np.broadcast_to(self.g[:, None, None, None], (3,) + self.shape).copy()
            if self.markers is not None:
                self.last_F = self.ib_force()
                F += self.last_F
            self.f = self.stream(self.collide(F))

    def fields(self):
        rho = self.f.sum(0)
        j = np.stack([(CX[:, None, None, None] * self.f).sum(0), (CY[:, None, None, None] * self.f).sum(0), (CZ[:, None, None, None] * self.f).sum(0)])
        return rho, j / rho


def start_fields(shape, seed):
    rng = np.random.default_rng(seed)
    nz, ny, nx = shape
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    u = 0.02 * np.stack([np.sin(2 * np.pi * x / nx) * np.cos(2 * np.pi * y / ny), -np.cos(2 * np.pi * x / nx) * np.sin(2 * np.pi * z / nz),
                         0.5 * np.cos(2 * np.pi * y / ny)]) + 1e-3 * rng.standard_normal((3,) + shape)
    rho = 1 + 0.01 * np.cos(2 * np.pi * x / nx) + 1e-3 * rng.standard_normal(shape)
    return rho.astype(np.float32).astype(float), u.astype(np.float32).astype(float)   # the ABI takes float32 fields


def both(g, kw_np, kw_abi, shape, steps, markers=None, f_tol=5e-15):
    P = g.BC_PERIODIC
    rho, u = start_fields(shape, 7)
    ref = NumpyStep(shape, **kw_np)
    ref.ib_passes = max(1, int(kw_abi.get("ib_iterations", 1)))
    ref.init(rho, u)
    nz, ny, nx = shape
    s = g.Sim(backend="oracle", nx=nx, ny=ny, nz=nz, max_markers=0 if markers is None else len(markers[0]), max_links=1,
              **dict(dict(bc=[P] * 6), **kw_abi))
    s.set_fields(rho, u)
    if markers is not None:
        X, U, dV, origin = markers
        ref.set_markers(X, U, dV, origin)
        s.set_markers(X, U, dV, np.zeros(len(X), np.int32))
        s.set_link_origins([origin])
    # the oracle initialises from the same float32 fields in fp64: the starting populations must already agree
    f0 = s.get_populations()
    assert np.abs(f0 - ref.f).max() < 1e-7     # get_populations travels as float32
    worst = 0.0
    for _ in range(steps):
        ref.step()
        s.step()
        ro, uo = s.get_fields(f64=True)
        rn, un = ref.fields()
        worst = max(worst, float(np.abs(ro - rn).max()), float(np.abs(uo - un).max()))
    assert worst < f_tol, worst
    return s, ref


@pytest.mark.parametrize("mrt", [False, True])
@pytest.mark.parametrize("force", [(0, 0, 0), (1e-4, -2e-4, 3e-4)])
def test_periodic_box(g, mrt, force):
    kw = dict(tau=0.73, mrt=mrt, g=force)
    s, _ = both(g, kw, dict(tau=0.73, collision=g.MRT if mrt else g.BGK, body_force=list(force)), (6, 7, 9), 12)
    s.close()


def test_custom_rate_vector(g):
    sn = 1 / 0.9
    magic = 8 * (2 - sn) / (8 - sn)
    rates = [0, 1.1, 1.3, 0, magic, 0, magic, 0, magic, sn, 1.5, sn, 1.5, sn, sn, sn, magic, magic, magic]
    s, _ = both(g, dict(tau=0.9, mrt=True, rates=rates, g=(2e-4, 0, 0)), dict(tau=0.9, collision=g.MRT, mrt_rates=rates, body_force=[2e-4, 0, 0]), (5, 6, 8), 10)
    s.close()


@pytest.mark.parametrize("mrt", [False, True])
def test_moving_y_walls(g, mrt):
    P, Wl, A = g.BC_PERIODIC, g.BC_WALL, g._abi
    kw = dict(tau=0.8, mrt=mrt, ywalls=True, wall_u_ylo=(-0.01, 0, 0.03), wall_u_yhi=(0.05, 0, 0.02), g=(0, 0, 1e-4))
    s, _ = both(g, kw, dict(tau=0.8, collision=g.MRT if mrt else g.BGK, bc=[P, P, Wl, Wl, P, P], body_force=[0, 0, 1e-4],
                            wall_u={A.YLO: [-0.01, 0, 0.03], A.YHI: [0.05, 0, 0.02]}), (6, 7, 9), 12)
    s.close()


def test_inlet_outlet_with_y_walls(g):
    P, Wl, IN, OUT = g.BC_PERIODIC, g.BC_WALL, g.BC_INLET, g.BC_OUTLET
    kw = dict(tau=0.8, mrt=True, ywalls=True, inlet_u=(0.01, 0, 0.04))
    s, _ = both(g, kw, dict(tau=0.8, collision=g.MRT, bc=[P, P, Wl, Wl, IN, OUT], inlet_u=[0.01, 0, 0.04]), (8, 6, 7), 12)
    s.close()


def test_immersed_sphere_direct_forcing(g):
    """A3-A7 together: index map (bit-exact), marker forces, spread force field, link wrench, and the flow they produce."""
    import util
    shape = (16, 14, 15)
    X = util.sphere_markers((7.3, 6.6, 8.2), 3.0, 60)
    U = np.tile(np.float32([0.01, -0.02, 0.015]), (60, 1))
    dV = np.full(60, 4 * np.pi * 9 / 60, np.float32)
    origin = [7.3, 6.6, 8.2]
    s, ref = both(g, dict(tau=0.8, mrt=True), dict(tau=0.8, collision=g.MRT), shape, 8, markers=(X, U, dV, origin), f_tol=2e-14)
    # after the loop both sides hold the forces of the LAST step they took
    base, _ = s.get_index_map()
    assert np.array_equal(base, ref.base)
    assert np.abs(s.get_marker_forces() - ref.Fm).max() < 1e-7            # float32 read-out
    assert np.abs(s.get_link_wrenches()[0] - ref.wrench).max() < 1e-12 * max(1.0, np.abs(ref.wrench).max())
    s.close()


@pytest.mark.parametrize("passes", [2, 5])
def test_immersed_sphere_multi_direct_forcing(g, passes):
    """FgConfig.ib_iterations > 1: the oracle's gather / spread passes against the matrix form F <- F + b - D D^T dV F."""
    import util
    shape = (16, 14, 15)
    X = util.sphere_markers((7.3, 6.6, 8.2), 3.0, 60)
    U = np.tile(np.float32([0.01, -0.02, 0.015]), (60, 1))
    dV = np.full(60, 4 * np.pi * 9 / 60, np.float32)
    origin = [7.3, 6.6, 8.2]
    s, ref = both(g, dict(tau=0.8, mrt=True), dict(tau=0.8, collision=g.MRT, ib_iterations=passes), shape, 8, markers=(X, U, dV, origin), f_tol=5e-14)
    assert np.abs(s.get_marker_forces() - ref.Fm).max() < 2e-7            # float32 read-out
    assert np.abs(s.get_link_wrenches()[0] - ref.wrench).max() < 1e-12 * max(1.0, np.abs(ref.wrench).max())
    assert util.rel_l2(s.get_force_field(), ref.last_F) < 1e-6
    s.close()
