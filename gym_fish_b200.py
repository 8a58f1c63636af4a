"""Import shim: the package directory is named ``gym-fish_b200`` (not a Python identifier), so
``import gym_fish_b200`` loads it from there under this importable name."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "gym-fish_b200")
_spec = _ilu.spec_from_file_location("gym_fish_b200", _os.path.join(_dir, "__init__.py"),
                                     submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["gym_fish_b200"] = _mod
_spec.loader.exec_module(_mod)
