"""Shared helpers for the parity tests (seeded synthetic inputs, norms, marker clouds, fish descriptions)."""
import os

import numpy as np

ORACLE_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "libfishgym_oracle.so")


def register_oracle(g):
    """The package does not know the CPU oracle (it is the checker, not a backend of the product): tests name it."""
    g.register_backend("oracle", ORACLE_LIB)
    return "oracle"


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300)))


def smooth_fields(shape, amp=0.02):
    nz, ny, nx = shape
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    u = np.zeros((3,) + tuple(shape))
    u[0] = amp * np.sin(2 * np.pi * x / nx) * np.cos(2 * np.pi * y / ny) + 0.3 * amp * np.sin(2 * np.pi * z / nz)
    u[1] = -amp * np.cos(2 * np.pi * x / nx) * np.sin(2 * np.pi * y / ny)
    u[2] = 0.5 * amp * np.cos(2 * np.pi * y / ny) * np.sin(2 * np.pi * z / nz)
    rho = 1 + 0.01 * np.cos(2 * np.pi * x / nx) * np.cos(2 * np.pi * z / nz)
    return rho, u


def taylor_green(n, plane="xy", U0=0.02):
    """2-D Taylor-Green vortex extruded along the third axis."""
    k = 2 * np.pi / n
    z, y, x = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    a, b = {"xy": (x, y), "yz": (y, z), "xz": (x, z)}[plane]
    ia, ib = {"xy": (0, 1), "yz": (1, 2), "xz": (0, 2)}[plane]
    u = np.zeros((3, n, n, n))
    u[ia] = U0 * np.sin(k * a) * np.cos(k * b)
    u[ib] = -U0 * np.cos(k * a) * np.sin(k * b)
    p = (U0 ** 2 / 4) * (np.cos(2 * k * a) + np.cos(2 * k * b))
    return 1 + 3 * p, u


def sphere_markers(center, radius, n):
    i = np.arange(n)
    z = 1 - (2 * i + 1) / n
    r = np.sqrt(1 - z * z)
    ph = np.pi * (3 - np.sqrt(5)) * i
    X = np.stack([center[0] + radius * r * np.cos(ph), center[1] + radius * r * np.sin(ph), center[2] + radius * z], 1)
    return X.astype(np.float32)


def fish_desc(g, root=(10, 9, 10), links=((8, 2.5), (7, 2.5), (6, 2), (5, 1.5)), free=1, heading=0.1):
    d = g.FgFishDesc()
    d.n_links = len(links)
    for k, (length, rad) in enumerate(links):
        d.link_len[k] = length
        d.link_rad[k] = rad
    for i in range(3):
        d.root_pos[i] = root[i]
    d.heading = heading
    d.density_ratio = 1.0
    d.joint_gain = 0.2
    d.joint_limit = 0.6
    d.joint_rate_max = 0.02
    d.free_root = free
    return d


# the parity case table shared by the emulation (CPU) and the CUDA (GPU) tests
def parity_cases(g):
    P, Wl, IN, OUT = g.BC_PERIODIC, g.BC_WALL, g.BC_INLET, g.BC_OUTLET
    A = g._abi
    N = dict(nx=12, ny=10, nz=8)
    return {
        "bgk_periodic": dict(N, tau=0.8),
        "mrt_periodic": dict(N, tau=0.7, collision=g.MRT),
        "mrt_force": dict(N, tau=0.7, collision=g.MRT, body_force=[1e-4, -2e-4, 3e-4]),
        "bgk_force": dict(N, tau=0.9, body_force=[1e-4, -2e-4, 3e-4]),
        "bgk_ywall_moving": dict(N, tau=0.8, bc=[P, P, Wl, Wl, P, P], wall_u={A.YHI: [0.05, 0, 0.02]}),
        "mrt_xy_walls": dict(N, tau=0.8, collision=g.MRT, bc=[Wl, Wl, Wl, Wl, P, P]),
        "mrt_all_walls_lid": dict(N, tau=0.8, collision=g.MRT, bc=[Wl] * 6, wall_u={A.ZHI: [0.03, 0.01, 0]}),
        "mrt_xwalls_moving": dict(N, tau=0.8, collision=g.MRT, bc=[Wl, Wl, P, P, P, P], wall_u={A.XLO: [0, 0.03, -0.02], A.XHI: [0, -0.01, 0.04]},
                                  body_force=[0, 1e-4, 0]),
        "bgk_inlet_outlet": dict(N, tau=0.8, bc=[P, P, P, P, IN, OUT], inlet_u=[0, 0, 0.04]),
        "mrt_inlet_outlet_ywalls": dict(N, tau=0.8, collision=g.MRT, bc=[P, P, Wl, Wl, IN, OUT], inlet_u=[0.01, 0, 0.04]),
        "mrt_outlet_inlet_xwalls": dict(N, tau=0.8, collision=g.MRT, bc=[Wl, Wl, P, P, OUT, IN], inlet_u=[0, 0.01, -0.04]),
        "mrt_ragged": dict(nx=7, ny=5, nz=3, tau=0.75, collision=g.MRT, bc=[P, P, Wl, Wl, P, P], body_force=[0, 0, 1e-4]),
        "bgk_wide_row": dict(nx=150, ny=3, nz=2, tau=0.8),
    }


def solid_block(kw):
    s = np.zeros((kw["nz"], kw["ny"], kw["nx"]), np.uint8)
    s[3:6, 2:5, 4:7] = 1
    s[0, 0, 0] = 1
    s[-1, 3, 3] = 1
    return s


def run_pair(g, ref_backend, test_backend, kw, steps=(1, 2, 3, 10, 11), solid=None):
    """Step both backends from the same seeded state; return the worst errors over the checkpoints."""
    a = g.Sim(backend=ref_backend, **kw)
    b = g.Sim(backend=test_backend, **kw)
    rho, u = smooth_fields(a.shape)
    if solid is not None:
        a.set_solid(solid)
        b.set_solid(solid)
    a.set_fields(rho, u)
    b.set_fields(rho, u)
    worst = dict(u=0.0, rho=0.0, f=0.0)
    done = 0
    for n in steps:
        a.step(n - done)
        b.step(n - done)
        done = n
        ra, ua = a.get_fields(f64=True)
        rb, ub = b.get_fields(f64=True)
        fa, fb = a.get_populations(), b.get_populations()
        if solid is not None:
            m = solid == 0
            ua, ub, ra, rb, fa, fb = ua * m, ub * m, ra * m, rb * m, fa * m, fb * m
        worst["u"] = max(worst["u"], rel_l2(ub, ua))
        worst["rho"] = max(worst["rho"], rel_l2(rb, ra))
        worst["f"] = max(worst["f"], float(np.abs(fa - fb).max()))
    a.close()
    b.close()
    return worst


# Lid-driven cavity at Re = 100: velocities on the two centre lines, in units of the lid speed, as tabulated by
# Ghia, Ghia & Shin, J. Comput. Phys. 48 (1982) 387, tables I and II (129 x 129 multigrid solution).  A published
# numerical benchmark — the one set of "golden numbers" for this path that does not come from this repository.
GHIA_RE100_U = np.array([   # (y, u) along x = 1/2
    [1.0000, 1.00000], [0.9766, 0.84123], [0.9688, 0.78871], [0.9609, 0.73722], [0.9531, 0.68717], [0.8516, 0.23151],
    [0.7344, 0.00332], [0.6172, -0.13641], [0.5000, -0.20581], [0.4531, -0.21090], [0.2813, -0.15662], [0.1719, -0.10150],
    [0.1016, -0.06434], [0.0703, -0.04775], [0.0625, -0.04192], [0.0547, -0.03717], [0.0000, 0.00000]])
GHIA_RE100_V = np.array([   # (x, v) along y = 1/2
    [1.0000, 0.00000], [0.9688, -0.05906], [0.9609, -0.07391], [0.9531, -0.08864], [0.9453, -0.10313], [0.9063, -0.16914],
    [0.8594, -0.22445], [0.8047, -0.24533], [0.5000, 0.05454], [0.2344, 0.17527], [0.2266, 0.17507], [0.1563, 0.16077],
    [0.0938, 0.12317], [0.0781, 0.10890], [0.0703, 0.10091], [0.0625, 0.09233], [0.0000, 0.00000]])


def cavity_vs_ghia(g, backend, n, steps, lid=0.1):
    """Lid-driven cavity in the x-y plane (lid = y-high wall moving along +x, z periodic and 2 cells deep), Re = 100.
    Returns the largest deviations of the centre-line profiles from Ghia et al., in units of the lid speed."""
    P, Wl = g.BC_PERIODIC, g.BC_WALL
    nu = lid * n / 100.0
    s = g.Sim(backend=backend, nx=n, ny=n, nz=2, tau=3 * nu + 0.5, collision=g.MRT, bc=[Wl, Wl, Wl, Wl, P, P],
              wall_u={g._abi.YHI: [lid, 0, 0]})
    s.step(steps)
    _, u = s.get_fields(f64=True)
    s.close()
    assert np.isfinite(u).all()
    # walls sit half a cell outside the first / last node; the centre line falls between two columns of nodes
    at = np.concatenate([[0.0], (np.arange(n) + 0.5) / n, [1.0]])
    ucl = np.concatenate([[0.0], 0.5 * (u[0][0, :, n // 2 - 1] + u[0][0, :, n // 2]) / lid, [1.0]])
    vcl = np.concatenate([[0.0], 0.5 * (u[1][0, n // 2 - 1, :] + u[1][0, n // 2, :]) / lid, [0.0]])
    du = np.abs(np.interp(GHIA_RE100_U[:, 0], at, ucl) - GHIA_RE100_U[:, 1]).max()
    dv = np.abs(np.interp(GHIA_RE100_V[:, 0], at, vcl) - GHIA_RE100_V[:, 1]).max()
    return float(du), float(dv), float(np.abs(u[2]).max())
