"""gym-fish_b200 — B200-native coupled fluid step behind a Fish-Gym-style underwater RL env.

Only what the hot path needs: ``csrc/`` (CUDA sm_100a kernels + the C ABI of
include/fishgym.h), the ctypes binding (``_abi``) and the Gym-style env (``env``).  numpy only —
importing this package must not import torch (BASELINE.json:5: "no PyTorch dependency on the sim path").
"""
from . import _abi
from ._abi import (BC_INLET, BC_OUTLET, BC_PERIODIC, BC_WALL, BGK, MRT, FgConfig, FgError, FgFishDesc, FgStats,
                   Sim, default_config, load_library, register_backend)

__all__ = ["_abi", "Sim", "FgConfig", "FgFishDesc", "FgStats", "FgError", "default_config", "load_library", "register_backend",
           "BGK", "MRT", "BC_PERIODIC", "BC_WALL", "BC_INLET", "BC_OUTLET"]
