"""fg_get_solid_force: momentum-exchange force of the fluid on obstacle cells (SURVEY.md Appendix A5 "Momentum-exchange
force on solids: sum over boundary links of c_i (f_i* + f_opp,in)"; with half-way bounce-back off a resting obstacle the
two populations of a link are equal, so every reflected population f hands the obstacle 2 f along the link).

CPU: the oracle against a numpy evaluation from its own populations, a steady-channel force balance and the zero net
pressure force on a closed body; the product's kernel (SolidForce<parity>, emulated) against the oracle on every
boundary kind, both parities, z-slabs.  GPU: the CUDA library against the oracle through the C ABI."""
import numpy as np
import pytest

import util


def numpy_force(g, f, solid, origin):
    """Fully periodic box, one rank: f [19][nz][ny][nx] = populations arriving at the cells (fg_get_populations)."""
    A = g._abi
    F, T = np.zeros(3), np.zeros(3)
    nz, ny, nx = solid.shape
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    for i in range(1, 19):
        came_from_obstacle = np.roll(solid, (A.CZ[i], A.CY[i], A.CX[i]), axis=(0, 1, 2)) == 1      # flag of x - c_i
        m = (solid == 0) & came_from_obstacle
        c = np.array([A.CX[i], A.CY[i], A.CZ[i]], dtype=np.float64)
        Fi = -c[None, :] * (2.0 * f[i][m].astype(np.float64))[:, None]
        r = np.stack([x[m] - 0.5 * c[0] - origin[0], y[m] - 0.5 * c[1] - origin[1], z[m] - 0.5 * c[2] - origin[2]], 1)
        F += Fi.sum(0)
        T += np.cross(r, Fi).sum(0)
    return np.concatenate([F, T])


@pytest.mark.parametrize("collision", ["bgk", "mrt"])
def test_oracle_against_numpy_from_its_own_populations(g, collision):
    rng = np.random.default_rng(5)
    kw = dict(nx=11, ny=9, nz=8, tau=0.7, collision=g.MRT if collision == "mrt" else g.BGK, body_force=[1e-4, -2e-4, 3e-4])
    s = g.Sim(backend="oracle", **kw)
    solid = (rng.random(s.shape) < 0.15).astype(np.uint8)
    rho, u = util.smooth_fields(s.shape)
    s.set_solid(solid)
    s.set_fields(rho, u)
    o = [4.5, 3.0, 2.5]
    for n in (0, 1, 1, 5):
        s.step(n)
        got = s.get_solid_force(o)
        ref = numpy_force(g, s.get_populations(), solid, o)          # float32 populations: ~1e-7 relative
        assert np.abs(got - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()), (n, got, ref)
        plain = s.get_solid_force()                 # no origin: no torque (the OpenMP reduction may associate differently per call)
        assert np.abs(plain[:3] - got[:3]).max() < 1e-13 and not plain[3:].any()
    s.close()


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("collision", ["bgk", "mrt"])
def test_oracle_momentum_balance_every_step(g, seed, collision):
    """Newton's third law, exactly: in a periodic box with obstacles the fluid's momentum changes per step by the body force on
    its cells minus what the bounce-back links handed the obstacles — P(t+1) - P(t) + F_obstacles = g x (fluid cells) to
    round-off, in an unsteady flow, for any mask.  Independent of how the read-out is evaluated."""
    rng = np.random.default_rng(seed)
    gf = np.array([1e-4, -2e-4, 3e-4])
    s = g.Sim(backend="oracle", nx=11, ny=9, nz=8, tau=float(rng.uniform(0.6, 1.2)), collision=g.MRT if collision == "mrt" else g.BGK,
              body_force=list(gf))
    solid = (rng.random(s.shape) < rng.uniform(0.05, 0.3)).astype(np.uint8)
    rho, u = util.smooth_fields(s.shape)
    s.set_solid(solid)
    s.set_fields(rho + 0.01 * rng.standard_normal(s.shape), u + 0.01 * rng.standard_normal((3,) + s.shape))
    fluid = solid == 0

    def momentum():
        r, v = s.get_fields(f64=True)
        return (r * v * fluid).sum(axis=(1, 2, 3))
    p0 = momentum()
    for _ in range(6):
        s.step(1)
        p1 = momentum()
        assert np.abs((p1 - p0) + s.get_solid_force()[:3] - gf * fluid.sum()).max() < 1e-13
        p0 = p1
    s.close()


def test_oracle_steady_channel_force_balance(g):
    """Two obstacle layers as channel walls, body force along z: once steady, the walls take exactly what the force puts
    into the fluid, g x (number of fluid cells), and nothing in the other directions."""
    nx, ny, nz, gf = 3, 12, 3, 1e-5
    s = g.Sim(backend="oracle", nx=nx, ny=ny, nz=nz, tau=0.9, collision=g.MRT, body_force=[0, 0, gf])
    solid = np.zeros(s.shape, np.uint8)
    solid[:, 0, :] = 1
    solid[:, -1, :] = 1
    s.set_solid(solid)
    s.step(6000)                       # H^2 / nu = 100 / 0.133: many diffusion times
    F = s.get_solid_force()
    n_fluid = int((solid == 0).sum())
    assert abs(F[2] / (gf * n_fluid) - 1.0) < 1e-6, F
    assert abs(F[0]) < 1e-12 and abs(F[1]) < 1e-9       # the two walls take equal and opposite pressure forces
    s.close()


def test_closed_body_in_fluid_at_rest_feels_no_force(g, emu):
    for backend in ("oracle", emu):
        s = g.Sim(backend=backend, nx=10, ny=9, nz=8, tau=0.8)
        solid = np.zeros(s.shape, np.uint8)
        solid[2:5, 3:6, 4:8] = 1
        s.set_solid(solid)
        for n in (0, 1, 2):
            s.step(n)
            F = s.get_solid_force([3.0, 4.0, 5.5])
            assert np.abs(F).max() < (1e-12 if backend == "oracle" else 2e-6), (backend, n, F)
        s.close()


def test_without_obstacles_the_force_is_zero(g, emu):
    for backend in ("oracle", emu):
        s = g.Sim(backend=backend, nx=6, ny=5, nz=4, tau=0.8, bc=[g.BC_WALL] * 6)
        s.step(2)
        assert not s.get_solid_force([1, 1, 1]).any()
        s.close()


def _pair(g, backend, kw, solid, origin=(5.0, 4.0, 3.0), steps=(0, 1, 1, 1, 7, 1)):
    a, b = g.Sim(backend="oracle", **kw), g.Sim(backend=backend, **kw)
    rho, u = util.smooth_fields(a.shape)
    for s in (a, b):
        s.set_solid(solid)
        s.set_fields(rho, u)
    worst = 0.0
    for n in steps:                      # read-outs at both storage parities of the AA pattern
        a.step(n)
        b.step(n)
        fa, fb = a.get_solid_force(origin), b.get_solid_force(origin)
        worst = max(worst, float(np.abs(fb - fa).max() / max(np.abs(fa).max(), 1e-3)))
    a.close()
    b.close()
    return worst


CASES = ["bgk_periodic", "mrt_force", "bgk_ywall_moving", "mrt_xy_walls", "mrt_all_walls_lid", "mrt_xwalls_moving", "bgk_inlet_outlet",
         "mrt_inlet_outlet_ywalls", "mrt_outlet_inlet_xwalls"]


@pytest.mark.parametrize("name", CASES)
def test_emulated_kernel_matches_oracle(g, emu, name):
    """Every boundary kind next to obstacles (walls are not obstacles; the outlet clamps its plane, the inlet has none),
    obstacle cells on the first and last plane and in a corner."""
    kw = util.parity_cases(g)[name]
    assert _pair(g, emu, kw, util.solid_block(kw)) <= 1e-5


@pytest.mark.parametrize("name", ["mrt_periodic", "mrt_xy_walls", "bgk_inlet_outlet"])
def test_emulated_random_masks(g, emu, name):
    rng = np.random.default_rng(11)
    kw = dict(util.parity_cases(g)[name], nx=34, ny=7, nz=6)         # rows that end inside a warp
    solid = (rng.random((kw["nz"], kw["ny"], kw["nx"])) < 0.2).astype(np.uint8)
    assert _pair(g, emu, kw, solid) <= 1e-5


def test_host_staged_slabs_sum_to_the_unsplit_force(g, emu):
    """Obstacles across the slab face and on the periodic seam: every rank returns the share of its own fluid cells."""
    A = g._abi
    kw = dict(nx=10, ny=8, nz=12, tau=0.8, collision=g.MRT, body_force=[1e-4, 0, 2e-4])
    solid = np.zeros((12, 8, 10), np.uint8)
    solid[4:8, 2:5, 3:7] = 1          # straddles the face between slabs 0 and 1 (planes 5 | 6)
    solid[0, 6, 1] = 1
    solid[11, 6, 1] = 1               # neighbours across the periodic seam
    one = g.Sim(backend="oracle", **kw)
    rho, u = util.smooth_fields(one.shape)
    one.set_solid(solid)
    one.set_fields(rho, u)
    for backend in ("oracle", emu):
        ranks = [g.Sim(backend=backend, n_ranks=2, rank=r, **kw) for r in range(2)]
        for r, s in enumerate(ranks):
            s.set_solid(solid)
            s.set_fields(rho[6 * r:6 * r + 6], u[:, 6 * r:6 * r + 6])
        ref = g.Sim(backend="oracle", **kw)
        ref.set_solid(solid)
        ref.set_fields(rho, u)
        o = [4.0, 3.0, 5.0]
        for _ in range(5):
            ref.step(1)
            for s in ranks:
                s.step(1)
            msgs = [(s.halo_pack(A.ZLO), s.halo_pack(A.ZHI)) for s in ranks]
            ranks[0].halo_unpack(A.ZHI, msgs[1][0])
            ranks[0].halo_unpack(A.ZLO, msgs[1][1])
            ranks[1].halo_unpack(A.ZLO, msgs[0][1])
            ranks[1].halo_unpack(A.ZHI, msgs[0][0])
            total = ranks[0].get_solid_force(o) + ranks[1].get_solid_force(o)
            want = ref.get_solid_force(o)
            assert np.abs(total - want).max() <= 1e-5 * max(np.abs(want).max(), 1e-3), (backend, total, want)
        for s in ranks + [ref]:
            s.close()
    one.close()


@pytest.mark.parametrize("overlap", [True, False])
def test_peered_slabs_sum_to_the_unsplit_force(g, emu, overlap):
    """The product's multi-GPU layout (halo pushes into the neighbour's lattice, emulated in one process): the read-out waits
    for the neighbours' halos of the last step like every other read-out, then each rank sums its own fluid cells."""
    kw = dict(nx=10, ny=8, nz=16, tau=0.8, collision=g.MRT, body_force=[1e-4, 0, 2e-4])
    solid = np.zeros((16, 8, 10), np.uint8)
    solid[6:10, 2:5, 3:7] = 1         # across the face between the slabs (planes 7 | 8)
    solid[0, 6, 1] = 1
    solid[15, 6, 1] = 1               # neighbours across the periodic seam
    whole = g.Sim(backend=emu, **kw)
    parts = [g.Sim(backend=emu, n_ranks=2, rank=r, flags=0 if overlap else g._abi.FLAG_NO_OVERLAP, **kw) for r in range(2)]
    rho, u = util.smooth_fields(whole.shape)
    whole.set_solid(solid)
    whole.set_fields(rho, u)
    for r, s in enumerate(parts):
        s.set_solid(solid)
        s.set_fields(rho[8 * r:8 * r + 8], u[:, 8 * r:8 * r + 8])
    h = [s.peer_export() for s in parts]
    parts[0].peer_connect(h[1], h[1])
    parts[1].peer_connect(h[0], h[0])
    o = [4.0, 3.0, 7.5]
    for it in range(6):
        whole.step(1)
        for s in parts:
            s.step(1)
        total = parts[0].get_solid_force(o) + parts[1].get_solid_force(o)
        want = whole.get_solid_force(o)
        assert np.abs(total - want).max() <= 1e-9 * max(np.abs(want).max(), 1e-3), (it, total, want)      # same fp32 populations, fp64 sums
    assert np.array_equal(whole.get_populations(), np.concatenate([s.get_populations() for s in parts], axis=1))


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mrt_force", "mrt_all_walls_lid", "mrt_inlet_outlet_ywalls", "mrt_outlet_inlet_xwalls"])
def test_gpu_matches_oracle(g, cuda, name):
    kw = util.parity_cases(g)[name]
    assert _pair(g, cuda, kw, util.solid_block(kw)) <= 1e-5


@pytest.mark.gpu
def test_gpu_staircase_sphere_in_a_channel(g, cuda):
    """A staircase sphere of 12 cells diameter in a 64 x 48 x 96 channel with inflow (rows of two whole warps, an obstacle
    far larger than a warp's cells): drag and torque of the momentum exchange against the oracle's, both parities."""
    P, Wl, IN, OUT = g.BC_PERIODIC, g.BC_WALL, g.BC_INLET, g.BC_OUTLET
    kw = dict(nx=64, ny=48, nz=96, tau=0.6, collision=g.MRT, bc=[P, P, Wl, Wl, IN, OUT], inlet_u=[0, 0, 0.04])
    z, y, x = np.meshgrid(np.arange(96), np.arange(48), np.arange(64), indexing="ij")
    solid = (((x - 31.5) ** 2 + (y - 23.5) ** 2 + (z - 30.5) ** 2) <= 36.0).astype(np.uint8)
    a, b = g.Sim(backend="oracle", **kw), g.Sim(backend=cuda, **kw)
    u = np.zeros((3,) + a.shape)
    u[2] = 0.04
    for s in (a, b):
        s.set_solid(solid)
        s.set_fields(np.ones(a.shape), u)
    o = [31.5, 23.5, 30.5]
    for n in (30, 1):
        a.step(n)
        b.step(n)
        fa, fb = a.get_solid_force(o), b.get_solid_force(o)
        assert fa[2] > 0                                      # drag along the flow
        assert np.abs(fb[:3] - fa[:3]).max() <= 1e-4 * np.abs(fa[:3]).max(), (fa, fb)
        assert np.abs(fb[3:] - fa[3:]).max() <= 1e-4 * max(np.abs(fa[3:]).max(), np.abs(fa[:3]).max()), (fa, fb)
    a.close()
    b.close()


def _hasimoto_array(g, backend, L=24, a=6.0, tau=1.0, steps=6000):
    """Stokes drag of a simple cubic array of spheres (tests/test_oracle_physics.py: Hasimoto 1959) through `backend`: an obstacle
    sphere in a fully periodic cube, body force g, drag by momentum exchange.  Returns (K / K_Hasimoto, F_z / (g x fluid cells),
    lateral force / F_z)."""
    import test_oracle_physics as T
    gf, nu = 1e-6, (tau - 0.5) / 3
    s = g.Sim(backend=backend, nx=L, ny=L, nz=L, tau=tau, collision=g.MRT, body_force=[0, 0, gf], mrt_rates=T._MAGIC(tau))
    c0 = L / 2 - 0.5
    z, y, x = np.meshgrid(np.arange(L), np.arange(L), np.arange(L), indexing="ij")
    solid = (((x - c0) ** 2 + (y - c0) ** 2 + (z - c0) ** 2) <= a * a).astype(np.uint8)
    s.set_solid(solid)
    s.step(steps)
    r, v = s.get_fields(f64=True)
    fluid = solid == 0
    U = ((v[2] + gf / 2 / r) * fluid).sum() / L ** 3
    F = s.get_solid_force()
    s.close()
    vol = float(solid.sum())
    K = gf * L ** 3 / (6 * np.pi * nu * (3 * vol / (4 * np.pi)) ** (1 / 3) * U)
    return K / T._hasimoto(vol / L ** 3), F[2] / (gf * fluid.sum()), float(np.abs(F[:2]).max() / F[2])


def test_hasimoto_array_drag_with_the_fp32_kernels_emulated(g, emu):
    """The product's fp32 kernels (shifted populations, a body force of 1e-6 per step on velocities of 1e-4) reproduce the
    oracle's number to six digits (at 24^3, a = 6: K / K_Hasimoto = 1.0111102 against 1.0111102 — run here at 16^3 to keep the suite short)."""
    ratio, balance, lateral = _hasimoto_array(g, emu, L=16, a=4.0, steps=3000)        # (24^3, a = 6: 1.01111 on both, 13 s)
    want, _, _ = _hasimoto_array(g, "oracle", L=16, a=4.0, steps=3000)
    assert abs(ratio - want) < 2e-6 and abs(want - 1) < 0.03 and abs(balance - 1) < 1e-6 and lateral < 1e-9, (ratio, want, balance, lateral)


@pytest.mark.gpu
def test_hasimoto_array_drag_gpu(g, cuda):
    """(Written without a GPU, like the test below; the emulated fp32 kernels give 1.0111.)  The CUDA path alone against the
    published analytic result, as test_gpu_physics.py does for the cavity table."""
    ratio, balance, lateral = _hasimoto_array(g, cuda)
    print(f"Hasimoto array on the GPU: K / K_Hasimoto = {ratio:.5f} (oracle and emulation 1.01111), force balance {balance:.7f}, lateral {lateral:.1e}")
    assert abs(ratio - 1) < 0.02 and abs(balance - 1) < 1e-4 and lateral < 1e-4, (ratio, balance, lateral)


@pytest.mark.gpu
def test_staircase_sphere_drag_by_momentum_exchange_full_size(g, cuda):
    """(Last GPU test of the suite on purpose: written when the round's GPU budget was spent, so its expectation comes from the
    oracle alone.)  SURVEY.md §4b: "... and via momentum-exchange on a staircase sphere as a cross-check".  The channel of test_gpu_physics.py::test_sphere_drag_full_size at
    Re = 20 with the sphere carved out of obstacle cells (half-way bounce-back; frontal area 448 cells against pi D^2 / 4 = 452.4)
    and the drag read with fg_get_solid_force.  The fp64 oracle gives Cd = 2.970 after 2000 steps, 2.958 after 4000 and 2.956 after 6000 at
    this size (profiles/r2_cpu_staircase_sphere_256x128x128_oracle.txt; 13 % above the unbounded-fluid correlation: the periodic array of spheres confines the flow; the immersed sphere of that
    test, hydrodynamically half a cell larger, is formed with D + 1).  Centred sphere: no lateral force, no torque."""
    P, IN, OUT = g.BC_PERIODIC, g.BC_INLET, g.BC_OUTLET
    D, Uin, Re = 24.0, 0.05, 20.0
    nu = Uin * D / Re
    s = g.Sim(backend=cuda, nx=128, ny=128, nz=256, tau=3 * nu + 0.5, collision=g.MRT, bc=[P, P, P, P, IN, OUT], inlet_u=[0, 0, Uin])
    c = (63.5, 63.5, 71.5)
    z, y, x = np.meshgrid(np.arange(256), np.arange(128), np.arange(128), indexing="ij")
    solid = (((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) <= (D / 2) ** 2).astype(np.uint8)
    del x, y, z
    s.set_solid(solid)
    u = np.zeros((3,) + s.shape, np.float32)
    u[2] = Uin
    s.set_fields(np.ones(s.shape, np.float32), u)
    ref = 0.5 * Uin ** 2 * np.pi * D ** 2 / 4
    hist = []
    for _ in range(6):
        s.step(2000)
        hist.append(s.get_solid_force(c))
    cd = [float(F[2] / ref) for F in hist]
    sn = 24 / Re * (1 + 0.15 * Re ** 0.687)
    print(f"staircase sphere, momentum exchange: Cd history {[round(v, 4) for v in cd]} (Schiller-Naumann {sn:.3f}, ratio {cd[-1] / sn:.3f}; "
          f"oracle 2.970 @ 2000, 2.958 @ 4000, 2.956 @ 6000); lateral {hist[-1][:2]}, torque {hist[-1][3:]}")
    assert abs(cd[0] - 2.9704) < 0.01 and abs(cd[1] - 2.9576) < 0.01 and abs(cd[2] - 2.9563) < 0.01      # the oracle's values at 2000, 4000, 6000 steps
    assert abs(cd[-1] - cd[-2]) / cd[-1] < 0.01                               # converged
    assert 1.05 * sn < cd[-1] < 1.22 * sn, (cd, sn)
    F = hist[-1]
    assert np.abs(F[:2]).max() < 1e-3 * F[2] and np.abs(F[3:]).max() < 1e-3 * F[2] * D
    s.close()
