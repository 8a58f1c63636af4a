"""Generate tests/golden/*.npz from the fp64 oracle (run here, committed; the GPU box only reads the files).

The reference checkout holds no golden vectors (README.md only), so these pin OUR oracle against drift and give the
GPU tests fixed fp64 answers that do not depend on the oracle library being rebuilt identically.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import gym_fish_b200 as g  # noqa: E402
import util  # noqa: E402

STEPS = 11
CASES = ["bgk_periodic", "mrt_force", "bgk_ywall_moving", "mrt_all_walls_lid", "mrt_inlet_outlet_ywalls", "mrt_ragged"]


def main():
    cases = util.parity_cases(g)
    for name in CASES:
        kw = cases[name]
        s = g.Sim(backend="oracle", **kw)
        rho, u = util.smooth_fields(s.shape)
        s.set_fields(rho, u)
        s.step(STEPS)
        r, v = s.get_fields(f64=True)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rho=r, u=v, steps=STEPS)
        s.close()
    # immersed boundary: sphere in a channel, 7 steps
    P, Wl, IN, OUT = g.BC_PERIODIC, g.BC_WALL, g.BC_INLET, g.BC_OUTLET
    kw = dict(nx=20, ny=18, nz=24, tau=0.8, collision=g.MRT, max_markers=512, max_links=1, bc=[P, P, Wl, Wl, IN, OUT], inlet_u=[0, 0, 0.05])
    s = g.Sim(backend="oracle", **kw)
    X = util.sphere_markers((10.3, 9.1, 8.2), 4.0, 200)
    s.set_markers(X, np.zeros_like(X), np.ones(200, np.float32), np.zeros(200, np.int32))
    s.set_link_origins([[10.3, 9.1, 8.2]])
    u = np.zeros((3,) + s.shape)
    u[2] = 0.05
    s.set_fields(np.ones(s.shape), u)
    s.step(7)
    r, v = s.get_fields(f64=True)
    base, owner = s.get_index_map()
    np.savez_compressed(os.path.join(HERE, "ib_sphere.npz"), rho=r, u=v, base=base, owner=owner, wrench=s.get_link_wrenches(),
                        Fm=s.get_marker_forces(), steps=7)
    print("wrote", len(CASES) + 1, "fixtures")


if __name__ == "__main__":
    main()
