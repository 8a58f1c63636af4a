#!/usr/bin/env python
"""A body driven by the CALLER's own integrator: markers up, link wrench down, every step (the loop bench.py times as `e2e`).

    python examples/prescribed_body.py               # CUDA library, 128 x 128 x 256 channel (needs a B200)
    python examples/prescribed_body.py --small       # 32 x 32 x 64

A sphere on a spring in a channel flow: the library couples whatever markers it is given (fg_set_markers) to the fluid and
returns the hydrodynamic force on them (fg_get_link_wrenches); the rigid-body dynamics — here one line of symplectic Euler —
stay with the caller.  This is the path for users who bring their own articulated-body engine instead of the built-in fish.
--lib PATH loads another library that exports include/fishgym.h (the test suite passes the CPU checker there)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import gym_fish_b200 as g  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--small", action="store_true")
ap.add_argument("--lib", default="cuda")
ap.add_argument("--steps", type=int, default=400)
ap.add_argument("--passes", type=int, default=1, help="direct-forcing passes per step (FgConfig.ib_iterations; > 1: multi-direct forcing)")
args = ap.parse_args()

nx, ny, nz, D = (32, 32, 64, 8.0) if args.small else (128, 128, 256, 24.0)
U0, Re = 0.04, 40.0
sim = g.Sim(backend=args.lib, nx=nx, ny=ny, nz=nz, tau=3 * U0 * D / Re + 0.5, collision=g.MRT, max_markers=4096, max_links=1, ib_iterations=args.passes,
            bc=[g.BC_PERIODIC, g.BC_PERIODIC, g.BC_WALL, g.BC_WALL, g.BC_INLET, g.BC_OUTLET], inlet_u=[0, 0, U0])
u = np.zeros((3,) + sim.shape, np.float32)
u[2] = U0
sim.set_fields(np.ones(sim.shape, np.float32), u)

# Fibonacci sphere, one marker per unit area; body frame offsets stay fixed, the centre moves
n = int(round(np.pi * D * D))
i = np.arange(n)
zz = 1 - (2 * i + 1) / n
rr = np.sqrt(1 - zz * zz)
ph = np.pi * (3 - np.sqrt(5)) * i
offsets = (D / 2) * np.stack([rr * np.cos(ph), rr * np.sin(ph), zz], 1)
dV = np.full(n, np.pi * D * D / n, np.float32)
link = np.zeros(n, np.int32)
X, Um = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)          # reused every step: the binding keeps their addresses

rest = np.array([nx / 2 + 0.2, ny / 2 + 0.1, nz / 4])
pos, vel = rest.copy(), np.zeros(3)
mass, k_spring = 4 / 3 * np.pi * (D / 2) ** 3 * 2.0, 0.02        # a body twice as dense as the fluid on a soft spring
# The coupling is explicit: this step's force comes from the velocity the body had when its markers were sent.  Direct forcing
# F_k = 2 (U_d - U*) answers a change of the body's velocity with the stiffness sum 2 dV per pass, and an explicit update is
# unstable once that exceeds ~2 x mass (added-mass instability: light bodies, or several passes).  A virtual mass
# Mv = 1/2 x stiffness low-pass filters the momentum increment — (mass + Mv) dp_new = mass F + Mv dp_old, fixed point dp = F —
# which is what the built-in fish integrator does (csrc/body.hpp).  With n passes the forcing's own period-2 memory is damped
# less and less (a pass aims at u* + F/2 = U_d, the fluid leaves the step with u* + F): the virtual mass grows like 2^n - 1.
Mv = 0.5 * (2 ** max(1, args.passes) - 1) * 2.0 * float(dV.sum())
dp = np.zeros(3)

t0 = time.perf_counter()
for it in range(args.steps):
    X[:] = pos + offsets
    Um[:] = vel
    sim.set_markers(X, Um, dV, link)          # host -> device: this step's marker positions and velocities
    sim.set_link_origins([pos])
    sim.step(1)                               # IB coupling + fused stream-collide; returns when the wrench is there
    force = sim.get_link_wrenches()[0, :3]    # device -> host: hydrodynamic force on the body during this step
    dp = (mass * (force - k_spring * (pos - rest)) + Mv * dp) / (mass + Mv)
    vel += dp / mass
    pos += vel
    if it % max(1, args.steps // 8) == 0:
        print(f"step {it:5d}  drag Fz {force[2]:+.4e}  displacement z {pos[2] - rest[2]:+.4f}")
sim.sync()
dt = time.perf_counter() - t0
print(f"{args.steps} coupled steps in {dt:.2f} s = {nx * ny * nz * args.steps / dt / 1e6:.0f} MLUPS end to end; the sphere sits "
      f"{pos[2] - rest[2]:+.3f} cells downstream of its spring's rest point")
assert np.isfinite(pos).all() and pos[2] > rest[2]
sim.close()
