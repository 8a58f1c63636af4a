"""Stream dependencies of the step orchestration, checked on the CPU.

The host emulation normally runs every operation at once, in submission order — which cannot notice a missing fork / join
between two launches that touch the same data.  With FG_EMU_SCHED=low | high | rand:<seed> it queues operations per
stream exactly as dev_cuda.cuh assigns them (launches on the current stream; copies, memsets, event records, neighbour
waits and signals on stream 0) and runs them only when the host waits, in an order that honours nothing but the recorded
event edges (tests/emu/dev_host.hpp).  The parity and bit-identity tests must hold under every policy: far-plane collide
beside the IB chain, thin wall-row branches, slab halos, bodies across faces, fused pairs.
(Removing the join of the far branch in sim.hpp step() makes `low` and `rand` fail; that is how the harness was checked.)

FG_EMU_GRAPHS=1 adds CUDA-graph semantics: a captured substep is recorded with the launch arguments of that moment and
replayed for every later substep with the same key, while what the host code submits is dropped — so a launch argument
that is not part of the key goes stale here exactly as it would on the GPU (test_stale_graph_key_would_be_noticed).
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SELECT = ["tests/test_emu_parity.py", "tests/test_slabs.py", "tests/test_random_cases.py", "tests/test_multi_direct_forcing.py", "tests/test_solid_force.py", "-k",
          "not gloo and not cuda and not 16_bit and not emulated_kernels_vs_oracle and not hasimoto or moving_markers"]


def test_results_do_not_depend_on_the_order_streams_are_served_in(g, emu):
    procs = {}
    # the last two also with emulated CUDA graphs (capture once per key, replay the RECORDED arguments afterwards)
    for policy, graphs in (("low", False), ("high", False), ("rand:1", True), ("rand:2", True)):
        env = dict(os.environ, FG_EMU_SCHED=policy, OMP_NUM_THREADS="2")
        if graphs:
            env["FG_EMU_GRAPHS"] = "1"
        procs[policy] = subprocess.Popen([sys.executable, "-m", "pytest", "-x", "-q", "-m", "not gpu", "-p", "no:cacheprovider"] + SELECT,
                                         cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    for policy, p in procs.items():
        out, _ = p.communicate(timeout=900)
        assert p.returncode == 0, (policy, out[-3000:])
        assert " passed" in out and "failed" not in out, (policy, out[-500:])
