"""The drop-in boundary: both libraries export every symbol include/fishgym.h declares, structs match, errors are
loud, and the sim path never imports torch.  No compute call into the CUDA library here (no GPU)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "fishgym.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fg_[a-z0-9_]+)\s*\(", src)))


def test_binding_lists_exactly_the_header(g):
    assert sorted(n for n, _, _ in g._abi.SYMBOLS) == header_functions()


@pytest.mark.parametrize("backend", ["cuda", "cuda_f16", "oracle"])
def test_library_exports_every_symbol(g, backend):
    path = g._abi.LIB_PATHS[backend]
    if backend.startswith("cuda") and not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(path)
    for name in header_functions():
        assert hasattr(lib, name), f"{path} does not export {name}"
    lib.fg_abi_version.restype = ctypes.c_int
    assert lib.fg_abi_version() == g._abi.FG_ABI_VERSION
    lib.fg_backend_name.restype = ctypes.c_char_p
    assert lib.fg_backend_name() == {"cuda": b"cuda-sm100a", "cuda_f16": b"cuda-sm100a-f16", "oracle": b"oracle-fp64"}[backend]


@pytest.mark.parametrize("backend", ["cuda", "cuda_f16", "oracle"])
def test_struct_layout_matches_the_library(g, backend):
    lib = g.load_library(backend)
    cfg = g.FgConfig()
    assert lib.fg_config_default(ctypes.byref(cfg)) == 0
    assert cfg.struct_size == ctypes.sizeof(g.FgConfig)
    assert (cfg.nx, cfg.collision, cfg.n_ranks, cfg.tau, cfg.inlet_rho) == (32, g.BGK, 1, 0.8, 1.0)


@pytest.mark.parametrize("lib", ["libfishgym_cuda.so", "libfishgym_cuda_f16.so"])
def test_cuda_library_is_sm100a_only(lib):
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "gym-fish_b200", "csrc", lib)],
                         capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_[89]\d", out), out


def test_bad_configs_are_rejected_with_a_message(g):
    for kw, frag in [(dict(tau=0.5), "tau"), (dict(nz=10, n_ranks=3), "slab"), (dict(collision=7), "collision"),
                     (dict(bc=[2, 2, 0, 0, 0, 0]), "z faces"), (dict(bc=[0, 1, 0, 0, 0, 0]), "both faces")]:
        with pytest.raises(g.FgError) as e:
            g.Sim(backend="oracle", **kw)
        assert e.value.code == g._abi.FG_EINVAL and frag in str(e.value)
    cfg = g.default_config()
    cfg.struct_size = 8
    with pytest.raises(g.FgError):
        g.Sim(cfg, backend="oracle")


def test_state_errors(g):
    s = g.Sim(backend="oracle", nx=8, ny=8, nz=8)
    with pytest.raises(g.FgError) as e:
        s.set_markers(np.zeros((3, 3)), np.zeros((3, 3)), 1.0)
    assert "max_markers" in str(e.value)
    with pytest.raises(g.FgError):
        s.set_action([0.0])
    s.close()


def test_no_cpu_fallback_without_gpu(g):
    """On a box without a GPU the product must fail loudly, not compute on the CPU."""
    lib = g.load_library("cuda")
    import shutil
    if shutil.which("nvidia-smi"):
        probe = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True)
        if probe.returncode == 0 and "GPU" in probe.stdout:
            pytest.skip("a GPU is present")
    with pytest.raises(g.FgError) as e:
        g.Sim(backend="cuda", nx=8, ny=8, nz=8)
    assert e.value.code == g._abi.FG_ECUDA and "no CPU path" in str(e.value)


def test_sim_path_does_not_import_torch():
    code = ("import sys; sys.path.insert(0, %r); import gym_fish_b200 as g; from gym_fish_b200 import env; "
            "s = g.Sim(backend=%r, nx=8, ny=8, nz=8); s.step(1); assert 'torch' not in sys.modules") % (ROOT, os.path.join(ROOT, "oracle", "libfishgym_oracle.so"))
    subprocess.run([sys.executable, "-c", code], check=True)


def test_plain_c_program_links_and_runs_against_the_abi(g, tmp_path):
    """include/fishgym.h is C (not C++): examples/minimal.c compiles as C99 and drives the library without Python.
    Linked against the oracle here so it runs without a GPU; the same source links against libfishgym_cuda.so."""
    exe = str(tmp_path / "minimal")
    odir = os.path.join(ROOT, "oracle")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "minimal.c"), "-o", exe, "-L", odir, "-lfishgym_oracle", "-lm",
                    "-Wl,-rpath," + odir], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "backend oracle-fp64" in r.stdout
    # and it at least LINKS against the CUDA library (running it needs a GPU)
    cdir = os.path.join(ROOT, "gym-fish_b200", "csrc")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "minimal.c"),
                    "-o", exe + "_cuda", "-L", cdir, "-lfishgym_cuda", "-lm", "-Wl,-rpath," + cdir], check=True)
