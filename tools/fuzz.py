#!/usr/bin/env python
"""Run the randomised test generators of tests/test_random_cases.py over any seed range (CPU emulation by default).

    python tools/fuzz.py static 0 2000                  # random configurations vs the oracle
    python tools/fuzz.py wide 0 2000                    # the same on rows where the two-cell kernels run (nx = 64 / 128 / 256, forced forms)
    python tools/fuzz.py moving|fish|api|slabs|xslabs 0 500
    FG_EMU_SCHED=rand:3 FG_EMU_GRAPHS=1 python tools/fuzz.py slabs 0 500     # queued streams, emulated graphs
    python tools/fuzz.py static 0 500 --lib cuda        # the real library on a GPU box
    python tools/fuzz.py moving|fish|xslabs 0 500 --passes 3 # ... with FgConfig.ib_iterations = 3 (multi-direct forcing)

Prints the seeds that fail; the committed tests run fixed ranges of the same generators."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gym_fish_b200 as g  # noqa: E402
import test_random_cases as t  # noqa: E402
import util  # noqa: E402

util.register_oracle(g)
kind, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
lib = sys.argv[sys.argv.index("--lib") + 1] if "--lib" in sys.argv else os.path.join(ROOT, "tests", "emu", "libfishgym_emu.so")
passes = int(sys.argv[sys.argv.index("--passes") + 1]) if "--passes" in sys.argv else 1
bad, ran = [], 0


def one(seed):
    if kind in ("static", "wide"):
        w, kw = t.run_case(g, lib, seed, wide=kind == "wide", solid_force=True)
        return (None if w is None else all(w[k] <= t.LIMITS[k] for k in w)), kw
    if kind == "moving":
        w, kw, _, _ = t.run_moving_markers_case(g, lib, seed, passes)
        return (None if w is None else all(w[k] <= dict(t.LIMITS, probe=5e-6)[k] for k in w)), kw
    if kind == "fish":
        w, kw = t.run_fish_case(g, lib, seed, passes)
        return (None if w is None else (w["obs"] <= 1e-4 and w["wrench"] <= 1e-4 and w["u"] <= 1e-5)), kw
    if kind == "api":
        w = t.run_api_sequence_case(g, lib, seed, solid_force=True)
        return (None if w is None else w <= 2e-5), None
    if kind == "slabs":
        return t.run_slab_case(g, lib, seed)
    if kind == "xslabs":
        w, kw, _ = t.run_bodies_across_slabs_case(g, lib, seed, passes)
        return (None if w is None else (w["f"] <= 1e-6 and w["wrench"] <= 1e-4)), kw
    raise SystemExit(__doc__)


for seed in range(lo, hi):
    try:
        ok, kw = one(seed)
    except g.FgError as e:      # a case that ends in an ABI error counts as failing (GPU runs: peer time-outs)
        ok, kw = False, f"{type(e).__name__}: {e}"
    if ok is None:
        continue
    ran += 1
    if not ok:
        bad.append(seed)
        print("SEED", seed, "FAILS", kw, flush=True)
print(f"{kind}: {ran} stable cases in [{lo}, {hi}), {len(bad)} failing: {bad[:20]}")
