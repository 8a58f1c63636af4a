"""pytest configuration: the `gpu` marker, library fixtures and shared helpers.

-m "not gpu": oracle vs analytic results and golden vectors, host logic, the CUDA kernel bodies run through the
              test-only host emulation (tests/emu), ABI symbol checks.  No compute call into the CUDA library.
-m gpu      : parity of the CUDA library against the oracle, through the C ABI, on a B200.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

ORACLE = os.path.join(ROOT, "oracle", "libfishgym_oracle.so")
EMU = os.path.join(ROOT, "tests", "emu", "libfishgym_emu.so")
EMU_F16 = os.path.join(ROOT, "tests", "emu", "libfishgym_emu_f16.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


_MADE = set()


def _ensure(path, directory):
    """Always `make` (once per session and directory): the Makefiles track header dependencies, so this is a no-op when
    the library is current and a rebuild when a header was edited after the (git-ignored) binary was built."""
    if directory not in _MADE:
        _MADE.add(directory)
        r = subprocess.run(["make", "-C", os.path.join(ROOT, directory)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0 and not os.path.exists(path):
            raise RuntimeError(f"make -C {directory} failed:\n{r.stdout[-3000:]}")
    return path


@pytest.fixture(scope="session")
def g():
    import gym_fish_b200
    # the package itself does not know the oracle: the checker is named here, by the tests
    gym_fish_b200.register_backend("oracle", _ensure(ORACLE, "oracle"))
    return gym_fish_b200


@pytest.fixture(scope="session")
def emu(g):
    return _ensure(EMU, "tests/emu")


@pytest.fixture(scope="session")
def emu_f16(g):
    """The 16-bit-storage build of the kernel bodies, emulated on the CPU (tests/emu, -DFG_POP16)."""
    return _ensure(EMU_F16, "tests/emu")


@pytest.fixture(scope="session")
def cuda_f16(g):
    _ensure(g._abi.LIB_PATHS["cuda_f16"], "gym-fish_b200/csrc")
    g.load_library("cuda_f16")
    return "cuda_f16"


@pytest.fixture(scope="session")
def cuda(g):
    """Backend name for GPU tests; fails (not skips) if the library is missing."""
    _ensure(g._abi.LIB_PATHS["cuda"], "gym-fish_b200/csrc")
    g.load_library("cuda")
    return "cuda"
