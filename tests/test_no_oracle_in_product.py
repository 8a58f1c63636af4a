"""The oracle is test infrastructure: nothing under gym-fish_b200/ (the product) or examples/ may name, load or link it."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_sources_do_not_reference_the_oracle_library():
    """Python / Makefiles of the product and the examples hold no path to oracle/; C / C++ / CUDA sources include nothing
    from it (comments that say what the checker is are fine)."""
    bad = []
    for top in ("gym-fish_b200", "examples", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                path = os.path.join(d, f)
                if f.endswith(".py") or f == "Makefile":
                    pat = r"libfishgym_oracle|fg_oracle\.cpp|[\"'/]oracle[\"'/]"
                elif f.endswith((".cu", ".cuh", ".hpp", ".h", ".c")):
                    pat = r"#\s*include\s*[\"<][^\">]*oracle"
                else:
                    continue
                if re.search(pat, open(path, errors="replace").read()):
                    bad.append(os.path.relpath(path, ROOT))
    assert not bad, bad


def test_package_ships_only_the_cuda_backends(g):
    import importlib
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import gym_fish_b200 as g; print(sorted(g._abi.LIB_PATHS))" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], check=True, capture_output=True, text=True).stdout.strip()
    assert out == "['cuda', 'cuda_f16']", out
    assert importlib.import_module("gym_fish_b200") is g
