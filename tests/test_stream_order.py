"""Stream dependencies of the step orchestration, checked on the CPU.

The host emulation normally runs every operation at once, in submission order — which cannot notice a missing fork / join
between two launches that touch the same data.  With FG_EMU_SCHED=low | high | rand:<seed> it queues operations per
stream exactly as dev_cuda.cuh assigns them (launches on the current stream; copies, memsets, event records, neighbour
waits and signals on stream 0) and runs them only when the host waits, in an order that honours nothing but the recorded
event edges (tests/emu/dev_host.hpp).  The parity and bit-identity tests must hold under every policy: far-plane collide
beside the IB chain, thin wall-row branches, slab halos, bodies across faces, fused pairs, the wavefront pairs.
(Removing the join of the far branch in sim.hpp step() makes `low` and `rand` fail; that is how the harness was checked.)
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SELECT = ["tests/test_emu_parity.py", "tests/test_slabs.py", "tests/test_random_cases.py", "-k",
          "not gloo and not cuda and not 16_bit and not emulated_kernels_vs_oracle or moving_markers"]


def test_results_do_not_depend_on_the_order_streams_are_served_in(g, emu):
    procs = {}
    for policy in ("low", "high", "rand:1", "rand:2"):
        env = dict(os.environ, FG_EMU_SCHED=policy, OMP_NUM_THREADS="2")
        procs[policy] = subprocess.Popen([sys.executable, "-m", "pytest", "-x", "-q", "-m", "not gpu", "-p", "no:cacheprovider"] + SELECT,
                                         cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    for policy, p in procs.items():
        out, _ = p.communicate(timeout=900)
        assert p.returncode == 0, (policy, out[-3000:])
        assert " passed" in out and "failed" not in out, (policy, out[-500:])
