"""Known-answer tests of the algorithm specification, in numpy.

The reference ships no tests or golden vectors (/root/reference/README.md is the whole checkout), so these
algebraic identities are what pins the lattice, the moment basis, the Guo source and the Peskin kernel that both
the oracle (matrix form) and the CUDA kernels (closed forms in lbm_core.cuh) implement."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def lat(g):
    A = g._abi
    c = np.stack([A.CX, A.CY, A.CZ], 1).astype(float)
    return A, c, A.W.astype(float)


def moment_matrix(c):
    x, y, z = c[:, 0], c[:, 1], c[:, 2]
    c2 = x * x + y * y + z * z
    return np.stack([
        np.ones(19), 19 * c2 - 30, (21 * c2 * c2 - 53 * c2 + 24) / 2,
        x, (5 * c2 - 9) * x, y, (5 * c2 - 9) * y, z, (5 * c2 - 9) * z,
        3 * x * x - c2, (3 * c2 - 5) * (3 * x * x - c2), y * y - z * z, (3 * c2 - 5) * (y * y - z * z),
        x * y, y * z, x * z, (y * y - z * z) * x, (z * z - x * x) * y, (x * x - y * y) * z])


def feq(c, w, rho, u):
    cu = c @ u
    return w * rho * (1 + 3 * cu + 4.5 * cu * cu - 1.5 * (u @ u))


def test_lattice_constants(lat):
    A, c, w = lat
    assert abs(w.sum() - 1) < 1e-15
    assert np.allclose(w @ c, 0, atol=1e-16)
    assert np.allclose((w[:, None, None] * c[:, :, None] * c[:, None, :]).sum(0), np.eye(3) / 3, atol=1e-15)
    assert np.array_equal(c[A.OPP], -c)
    assert sorted(np.where(c[:, 2] == 1)[0]) == [5, 11, 12, 15, 16]
    assert sorted(np.where(c[:, 2] == -1)[0]) == [6, 13, 14, 17, 18]
    assert sorted(np.where(c[:, 0] == 1)[0]) == [1, 7, 9, 11, 13]


def test_moment_basis_orthogonal_with_published_norms(lat):
    _, c, _ = lat
    M = moment_matrix(c)
    G = M @ M.T
    assert np.allclose(G - np.diag(np.diag(G)), 0, atol=1e-12)
    assert np.allclose(np.diag(G), [19, 2394, 252, 10, 40, 10, 40, 10, 40, 36, 72, 12, 24, 4, 4, 4, 8, 8, 8])
    # shifted populations: M w is non-zero only on rho, e, eps (lbm_core.cuh collide_mrt relies on it)
    assert np.allclose(M @ lat[2], [1, -11, 3] + [0] * 16, atol=1e-14)


def test_equilibrium_moments_closed_form(lat):
    _, c, w = lat
    M = moment_matrix(c)
    rng = np.random.default_rng(0)
    for _ in range(5):
        rho, u = 1 + 0.1 * rng.standard_normal(), 0.1 * rng.standard_normal(3)
        j = rho * u
        jj = j @ j / rho
        pxx = (2 * j[0] ** 2 - j[1] ** 2 - j[2] ** 2) / rho
        pww = (j[1] ** 2 - j[2] ** 2) / rho
        meq = np.array([rho, -11 * rho + 19 * jj, 3 * rho - 5.5 * jj, j[0], -2 / 3 * j[0], j[1], -2 / 3 * j[1], j[2], -2 / 3 * j[2],
                        pxx, -0.5 * pxx, pww, -0.5 * pww, j[0] * j[1] / rho, j[1] * j[2] / rho, j[0] * j[2] / rho, 0, 0, 0])
        assert np.allclose(M @ feq(c, w, rho, u), meq, atol=1e-14)


def test_guo_source_moments_closed_form(lat):
    _, c, w = lat
    M = moment_matrix(c)
    rng = np.random.default_rng(1)
    u, F = 0.1 * rng.standard_normal(3), 1e-2 * rng.standard_normal(3)
    phi = w * ((3 * (c - u) + 9 * (c @ u)[:, None] * c) @ F)
    uF = u @ F
    a = 2 * u[0] * F[0] - u[1] * F[1] - u[2] * F[2]
    b = u[1] * F[1] - u[2] * F[2]
    closed = np.array([0, 38 * uF, -11 * uF, F[0], -2 / 3 * F[0], F[1], -2 / 3 * F[1], F[2], -2 / 3 * F[2], 2 * a, -a, 2 * b, -b,
                       u[0] * F[1] + u[1] * F[0], u[1] * F[2] + u[2] * F[1], u[0] * F[2] + u[2] * F[0], 0, 0, 0])
    assert np.allclose(M @ phi, closed, atol=1e-16)
    assert abs(phi.sum()) < 1e-17 and np.allclose(phi @ c, F, atol=1e-16)


def test_mrt_with_equal_rates_is_bgk(lat):
    _, c, w = lat
    M = moment_matrix(c)
    Minv = M.T / (M * M).sum(1)
    rng = np.random.default_rng(2)
    f = feq(c, w, 1.02, np.array([0.03, -0.02, 0.01])) + 1e-3 * rng.standard_normal(19)
    rho, u = f.sum(), f @ c / f.sum()
    om = 1.3
    bgk = f - om * (f - feq(c, w, rho, u))
    mrt = Minv @ (M @ f - om * (M @ f - M @ feq(c, w, rho, u)))
    assert np.allclose(bgk, mrt, atol=1e-15)


def peskin(r):
    r = np.abs(r)
    return np.where(r < 1, (3 - 2 * r + np.sqrt(np.maximum(0, 1 + 4 * r - 4 * r * r))) / 8,
                    np.where(r < 2, (5 - 2 * r - np.sqrt(np.maximum(0, -7 + 12 * r - 4 * r * r))) / 8, 0.0))


@pytest.mark.parametrize("X", [0.0, 0.3, 0.5, 0.77, 12.999])
def test_peskin_identities(X):
    nodes = np.floor(X) - 1 + np.arange(4)
    ph = peskin(X - nodes)
    assert abs(ph.sum() - 1) < 1e-14
    assert abs(((nodes - X) * ph).sum()) < 1e-14
    assert abs((ph ** 2).sum() - 3 / 8) < 1e-14
    assert abs(ph[0::2].sum() - 0.5) < 1e-14 and abs(ph[1::2].sum() - 0.5) < 1e-14
