"""ctypes binding of include/fishgym.h — numpy only, no torch on the sim path.

The product loads ``csrc/libfishgym_cuda.so`` and fails loudly when it is missing; there is no CPU
fallback, and this package does not know where the CPU oracle lives.  Checkers (tests/, ``__graft_entry__.smoke()``,
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``) name it themselves with ``register_backend``.

Boundary defined by BASELINE.json:5 (the reference ships no interface: /root/reference/README.md:14).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

FG_ABI_VERSION = 8
Q = 19

FG_OK, FG_EINVAL, FG_ENOMEM, FG_ECUDA, FG_ESTATE, FG_ENOTSUP, FG_EPEER = 0, -1, -2, -3, -4, -5, -6
BGK, MRT = 0, 1
XLO, XHI, YLO, YHI, ZLO, ZHI = range(6)
BC_PERIODIC, BC_WALL, BC_INLET, BC_OUTLET = range(4)
FLAG_NO_OVERLAP = 1
FLAG_PROFILE = 2
FLAG_NO_GRAPHS = 4
FLAG_FUSED_IB = 8
FLAG_NO_SPLIT = 16
FLAG_NO_SWEEP_FLIP = 32
FLAG_FUSED_PAIRS = 64
FLAG_NO_XWARP = 128
FLAG_SYNC_STEP = 256
FLAG_IB_TILE_SPREAD = 512
FLAG_EVEN_VEC4 = 1024
FLAG_EVEN_VEC2 = 2048
FLAG_EVEN_SCALAR = 8192
FLAG_ODD_VEC2 = 16384
FLAG_ODD_SCALAR = 32768

_ERR_NAMES = {FG_EINVAL: "FG_EINVAL", FG_ENOMEM: "FG_ENOMEM", FG_ECUDA: "FG_ECUDA", FG_ESTATE: "FG_ESTATE",
              FG_ENOTSUP: "FG_ENOTSUP", FG_EPEER: "FG_EPEER"}

# lattice constants, exported for tests and host-side helpers
CX = np.array([0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0])
CY = np.array([0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1])
CZ = np.array([0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1])
OPP = np.array([0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15])
W = np.array([1 / 3] + [1 / 18] * 6 + [1 / 36] * 12)


class FgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{_ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class FgConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("collision", C.c_int32),
        ("bc", C.c_int32 * 6),
        ("n_ranks", C.c_int32), ("rank", C.c_int32),
        ("device", C.c_int32),
        ("max_markers", C.c_int32),
        ("max_links", C.c_int32),
        ("flags", C.c_int32),
        ("split_min_cells", C.c_int32),
        ("pair_lag", C.c_int32),
        ("ib_iterations", C.c_int32),
        ("tau", C.c_double),
        ("mrt_rates", C.c_double * 19),
        ("wall_u", (C.c_double * 3) * 6),
        ("inlet_u", C.c_double * 3),
        ("inlet_rho", C.c_double),
        ("body_force", C.c_double * 3),
        ("reserved_d", C.c_double * 4),
    ]


class FgStats(C.Structure):
    _fields_ = [
        ("steps", C.c_int64), ("cells", C.c_int64),
        ("last_step_ms", C.c_double), ("last_mlups", C.c_double),
        ("kernel_launches", C.c_int64),
        ("n_markers", C.c_int32), ("n_links", C.c_int32),
        ("band_cells", C.c_int32), ("parity", C.c_int32),
        ("collide_ms", C.c_double), ("collide_launches", C.c_int64), ("ib_ms", C.c_double),
        ("collide_cells", C.c_int64), ("split_substeps", C.c_int64), ("pair_substeps", C.c_int64), ("graph_launches", C.c_int64),
    ]


class FgFishDesc(C.Structure):
    _fields_ = [
        ("n_links", C.c_int32), ("markers_per_link", C.c_int32),
        ("link_len", C.c_double * 8), ("link_rad", C.c_double * 8),
        ("root_pos", C.c_double * 3),
        ("heading", C.c_double), ("density_ratio", C.c_double),
        ("joint_gain", C.c_double), ("joint_limit", C.c_double), ("joint_rate_max", C.c_double),
        ("free_root", C.c_int32), ("reserved", C.c_int32),
    ]


class FgPeerHandle(C.Structure):
    _fields_ = [("bytes", C.c_ubyte * 320)]


_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATHS = {
    # FG_CUDA_LIB: another build of the same library (kernel experiments, e.g. different launch bounds)
    "cuda": os.environ.get("FG_CUDA_LIB") or os.path.join(_HERE, "csrc", "libfishgym_cuda.so"),
    # opt-in build of the same sources: 16-bit population storage, fp32 arithmetic (half the HBM bytes per update)
    "cuda_f16": os.path.join(_HERE, "csrc", "libfishgym_cuda_f16.so"),
}


def register_backend(name: str, path: str) -> None:
    """Give another library that exports the ABI of include/fishgym.h a short name (``Sim(backend=name)``).
    The product registers nothing: tests, smoke() and bench.py's CPU legs use this for the oracle."""
    if name in ("cuda", "cuda_f16"):
        raise ValueError("the product backends cannot be re-pointed; use FG_CUDA_LIB for kernel experiments")
    LIB_PATHS[name] = os.path.abspath(path)


# every symbol include/fishgym.h declares: (name, restype, argtypes)
_P = C.c_void_p
_f32 = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
SYMBOLS = [
    ("fg_abi_version", C.c_int, []),
    ("fg_backend_name", C.c_char_p, []),
    ("fg_last_error", C.c_char_p, [_P]),
    ("fg_config_default", C.c_int, [C.POINTER(FgConfig)]),
    ("fg_create", C.c_int, [C.POINTER(FgConfig), C.POINTER(_P)]),
    ("fg_destroy", C.c_int, [_P]),
    ("fg_reset", C.c_int, [_P, C.c_uint64]),
    ("fg_set_fields", C.c_int, [_P, _f32, _f32]),
    ("fg_get_fields", C.c_int, [_P, _f32, _f32]),
    ("fg_get_fields_f64", C.c_int, [_P, _f64, _f64]),
    ("fg_set_populations", C.c_int, [_P, _f32]),
    ("fg_get_populations", C.c_int, [_P, _f32]),
    ("fg_set_solid", C.c_int, [_P, C.c_void_p]),
    ("fg_set_markers", C.c_int, [_P, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("fg_set_link_origins", C.c_int, [_P, C.c_int32, C.c_void_p]),
    ("fg_get_index_map", C.c_int, [_P, _i32, _i32]),
    ("fg_get_marker_forces", C.c_int, [_P, _f32]),
    ("fg_get_marker_velocities", C.c_int, [_P, _f32]),
    ("fg_get_link_wrenches", C.c_int, [_P, C.c_void_p]),
    ("fg_get_force_field", C.c_int, [_P, _f32]),
    ("fg_probe", C.c_int, [_P, C.c_int32, _f32, _f32]),
    ("fg_add_fish", C.c_int, [_P, C.POINTER(FgFishDesc), C.POINTER(C.c_int32)]),
    ("fg_set_action", C.c_int, [_P, C.c_void_p, C.c_int32]),
    ("fg_get_obs", C.c_int, [_P, _f32, C.c_int32]),
    ("fg_obs_size", C.c_int, [_P]),
    ("fg_action_size", C.c_int, [_P]),
    ("fg_get_markers", C.c_int, [_P, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    ("fg_step", C.c_int, [_P, C.c_int32]),
    ("fg_sync", C.c_int, [_P]),
    ("fg_get_stats", C.c_int, [_P, C.POINTER(FgStats)]),
    ("fg_check_finite", C.c_int, [_P, C.POINTER(C.c_int64)]),
    ("fg_get_solid_force", C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    ("fg_set_flags", C.c_int, [_P, C.c_int32]),
    ("fg_halo_bytes", C.c_int64, [_P]),
    ("fg_halo_pack", C.c_int, [_P, C.c_int32, C.c_void_p]),
    ("fg_halo_unpack", C.c_int, [_P, C.c_int32, C.c_void_p]),
    ("fg_peer_export", C.c_int, [_P, C.POINTER(FgPeerHandle)]),
    ("fg_peer_connect", C.c_int, [_P, C.POINTER(FgPeerHandle), C.POINTER(FgPeerHandle)]),
    ("fg_peer_connect_all", C.c_int, [_P, C.POINTER(FgPeerHandle), C.c_int32]),
]

_LIBS: dict = {}


def load_library(backend: str = "cuda") -> C.CDLL:
    """dlopen a backend and bind every symbol of the ABI.  Raises if the library is missing:
    the product never degrades to another backend."""
    if backend in _LIBS:
        return _LIBS[backend]
    path = LIB_PATHS.get(backend, backend)
    if not os.path.exists(path):
        hint = ("run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a)" if backend.startswith("cuda")
                else "not a registered backend name nor the path of a library (register_backend)")
        raise FileNotFoundError(f"fishgym backend '{backend}': {path} is missing; {hint}")
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)   # AttributeError here means the library does not export the ABI
        fn.restype = restype
        fn.argtypes = argtypes
    ver = lib.fg_abi_version()
    if ver != FG_ABI_VERSION:
        raise RuntimeError(f"{path}: ABI version {ver}, binding expects {FG_ABI_VERSION}")
    _LIBS[backend] = lib
    return lib


def default_config(**kw) -> FgConfig:
    """FgConfig with the same defaults as fg_config_default(), overridable by keyword."""
    cfg = FgConfig()
    cfg.struct_size = C.sizeof(FgConfig)
    cfg.nx = cfg.ny = cfg.nz = 32
    cfg.collision = BGK
    cfg.n_ranks = 1
    cfg.tau = 0.8
    cfg.inlet_rho = 1.0
    for k, v in kw.items():
        if k == "bc":
            for i, b in enumerate(v):
                cfg.bc[i] = int(b)
        elif k == "wall_u":
            for face, vec in dict(v).items():
                for d in range(3):
                    cfg.wall_u[face][d] = float(vec[d])
        elif k in ("mrt_rates", "inlet_u", "body_force"):
            arr = getattr(cfg, k)
            for i, x in enumerate(v):
                arr[i] = float(x)
        else:
            if not hasattr(cfg, k):
                raise TypeError(f"FgConfig has no field '{k}'")
            setattr(cfg, k, v)
    return cfg


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Sim:
    """Thin object wrapper over one FgSim handle (reset/step plumbing for the env and the tests)."""

    def __init__(self, cfg: Optional[FgConfig] = None, backend: str = "cuda", **kw):
        self.lib = load_library(backend)
        self.backend = backend
        self.cfg = cfg if cfg is not None else default_config(**kw)
        h = C.c_void_p()
        rc = self.lib.fg_create(C.byref(self.cfg), C.byref(h))
        if rc != FG_OK:
            raise FgError(rc, (self.lib.fg_last_error(None) or b"").decode())
        self.h = h
        self.nx, self.ny = self.cfg.nx, self.cfg.ny
        self.nz = self.cfg.nz // self.cfg.n_ranks   # local slab height
        self.shape = (self.nz, self.ny, self.nx)
        self._counts = (0, 0)
        self._counts_stale = False
        self._marker_args = None
        # link wrenches come back into a buffer of FULL capacity: the library writes n_links rows, and n_links can grow
        # without the marker count changing (same markers, higher link ids)
        self._wrench_buf = np.zeros((max(self.cfg.max_links, 1), 6), dtype=np.float64)
        self._wrench_ptr = self._wrench_buf.ctypes.data

    # -- plumbing --
    def _ck(self, rc: int):
        if rc < 0:
            raise FgError(rc, (self.lib.fg_last_error(self.h) or b"").decode())
        return rc

    def close(self):
        if getattr(self, "h", None):
            self.lib.fg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def backend_name(self) -> str:
        return self.lib.fg_backend_name().decode()

    # -- fluid --
    def reset(self, seed: int = 0):
        self._ck(self.lib.fg_reset(self.h, seed))

    def set_fields(self, rho: np.ndarray, u: np.ndarray):
        rho = np.ascontiguousarray(rho, dtype=np.float32).reshape(self.shape)
        u = np.ascontiguousarray(u, dtype=np.float32).reshape((3,) + self.shape)
        self._ck(self.lib.fg_set_fields(self.h, rho, u))

    def get_fields(self, f64: bool = False):
        dt = np.float64 if f64 else np.float32
        rho = np.empty(self.shape, dtype=dt)
        u = np.empty((3,) + self.shape, dtype=dt)
        fn = self.lib.fg_get_fields_f64 if f64 else self.lib.fg_get_fields
        self._ck(fn(self.h, rho, u))
        return rho, u

    def set_populations(self, f: np.ndarray):
        f = np.ascontiguousarray(f, dtype=np.float32).reshape((Q,) + self.shape)
        self._ck(self.lib.fg_set_populations(self.h, f))

    def get_populations(self) -> np.ndarray:
        f = np.empty((Q,) + self.shape, dtype=np.float32)
        self._ck(self.lib.fg_get_populations(self.h, f))
        return f

    def set_solid(self, solid_global: Optional[np.ndarray]):
        if solid_global is None:
            self._ck(self.lib.fg_set_solid(self.h, None))
            return
        s = np.ascontiguousarray(solid_global, dtype=np.uint8).reshape((self.cfg.nz, self.ny, self.nx))
        self._ck(self.lib.fg_set_solid(self.h, _ptr(s)))

    # -- immersed boundary --
    @staticmethod
    def _ready(a, dtype, shape):
        return isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.c_contiguous and a.shape == shape

    def set_markers(self, X, U, dV, link_id=None):
        c = self._marker_args
        if c is not None and X is c[0] and U is c[1] and dV is c[2] and link_id is c[3] and X.shape[0] == c[4]:
            # the same (already validated) arrays as in the previous call: their addresses are known — `.ctypes.data`
            # alone costs 2 us per array, 8 of the 14 us this call took in a step-by-step loop
            n, px, pu, pd, pl = c[4:]
        else:
            n = len(X)
            f32, i32 = np.float32, np.int32
            ready = (self._ready(X, f32, (n, 3)) and self._ready(U, f32, (n, 3)) and self._ready(dV, f32, (n,))
                     and (link_id is None or self._ready(link_id, i32, (n,))))
            orig = (X, U, dV, link_id)
            if not ready:
                # slow path: convert; callers in a hot loop pass float32/int32 C-contiguous arrays and skip this
                X = np.ascontiguousarray(X, dtype=f32).reshape(-1, 3)
                n = X.shape[0]
                U = np.ascontiguousarray(U, dtype=f32).reshape(n, 3)
                dV = np.ascontiguousarray(np.broadcast_to(np.asarray(dV, dtype=f32), (n,)))
                link_id = None if link_id is None else np.ascontiguousarray(link_id, dtype=i32).reshape(n)
            px, pu, pd = X.ctypes.data, U.ctypes.data, dV.ctypes.data
            pl = None if link_id is None else link_id.ctypes.data
            # remembered only when the caller's own arrays are what the library reads (holding them keeps the addresses valid)
            self._marker_args = orig + (n, px, pu, pd, pl) if ready else None
        self._ck(self.lib.fg_set_markers(self.h, n, px, pu, pd, pl))
        if n != self._counts[0]:
            self._refresh_counts()
        else:
            self._counts_stale = True       # the link count may have changed: looked up again when a read-out needs it

    def set_link_origins(self, origins):
        o = np.ascontiguousarray(origins, dtype=np.float64).reshape(-1, 3)
        self._ck(self.lib.fg_set_link_origins(self.h, o.shape[0], _ptr(o)))
        self._refresh_counts()

    def _refresh_counts(self):
        st = self.stats()
        self._counts = (st.n_markers, st.n_links)
        self._counts_stale = False

    def stats(self) -> FgStats:
        st = FgStats()
        self._ck(self.lib.fg_get_stats(self.h, C.byref(st)))
        return st

    def get_index_map(self):
        n = self._counts[0]
        base = np.empty((n, 3), dtype=np.int32)
        owner = np.empty((n,), dtype=np.int32)
        self._ck(self.lib.fg_get_index_map(self.h, base, owner))
        return base, owner

    def get_marker_forces(self) -> np.ndarray:
        F = np.empty((self._counts[0], 3), dtype=np.float32)
        self._ck(self.lib.fg_get_marker_forces(self.h, F))
        return F

    def get_marker_velocities(self) -> np.ndarray:
        U = np.empty((self._counts[0], 3), dtype=np.float32)
        self._ck(self.lib.fg_get_marker_velocities(self.h, U))
        return U

    def get_link_wrenches(self) -> np.ndarray:
        """[n_links][6] hydrodynamic (force, torque) on each link; returns a view of a reused buffer."""
        if self._counts_stale:
            self._refresh_counts()
        self._ck(self.lib.fg_get_link_wrenches(self.h, self._wrench_ptr))
        return self._wrench_buf[: self._counts[1]]

    def get_force_field(self) -> np.ndarray:
        F = np.empty((3,) + self.shape, dtype=np.float32)
        self._ck(self.lib.fg_get_force_field(self.h, F))
        return F

    def probe(self, X) -> np.ndarray:
        """(rho, ux, uy, uz) of the fluid at points X[n][3] (4-point delta interpolation)."""
        X = np.ascontiguousarray(X, dtype=np.float32).reshape(-1, 3)
        out = np.empty((X.shape[0], 4), dtype=np.float32)
        self._ck(self.lib.fg_probe(self.h, X.shape[0], X, out))
        return out

    # -- bodies --
    def add_fish(self, desc: FgFishDesc) -> int:
        fid = C.c_int32(-1)
        self._ck(self.lib.fg_add_fish(self.h, C.byref(desc), C.byref(fid)))
        self._refresh_counts()
        return fid.value

    def action_size(self) -> int:
        return self._ck(self.lib.fg_action_size(self.h))

    def obs_size(self) -> int:
        return self._ck(self.lib.fg_obs_size(self.h))

    def set_action(self, a):
        a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
        self._ck(self.lib.fg_set_action(self.h, _ptr(a), a.shape[0]))

    def get_obs(self) -> np.ndarray:
        o = np.empty((self.obs_size(),), dtype=np.float32)
        self._ck(self.lib.fg_get_obs(self.h, o, o.shape[0]))
        return o

    def get_markers(self):
        n = self._ck(self.lib.fg_get_markers(self.h, None, None, None, 0))
        X = np.empty((n, 3), dtype=np.float32)
        U = np.empty((n, 3), dtype=np.float32)
        link = np.empty((n,), dtype=np.int32)
        self._ck(self.lib.fg_get_markers(self.h, _ptr(X), _ptr(U), _ptr(link), n))
        return X, U, link

    # -- stepping --
    def step(self, n_substeps: int = 1):
        self._ck(self.lib.fg_step(self.h, n_substeps))

    def sync(self):
        self._ck(self.lib.fg_sync(self.h))

    def check_finite(self) -> int:
        """Cells of the local slab whose rest population is non-finite or out of range (0 = the fluid has not diverged)."""
        n = C.c_int64(0)
        self._ck(self.lib.fg_check_finite(self.h, C.byref(n)))
        return int(n.value)

    def get_solid_force(self, origin=None) -> np.ndarray:
        """[6]: momentum-exchange force of the fluid on the obstacle cells (fg_set_solid) in the last step and, with `origin`
        (x, y, global z), its torque about that point; the share of this rank's fluid cells on z-slabs."""
        out = np.zeros(6, dtype=np.float64)
        o = None if origin is None else np.ascontiguousarray(origin, dtype=np.float64).reshape(3)
        self._ck(self.lib.fg_get_solid_force(self.h, None if o is None else o.ctypes.data_as(C.POINTER(C.c_double)),
                                             out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def set_flags(self, flags: int):
        self._ck(self.lib.fg_set_flags(self.h, int(flags)))
        self.cfg.flags = int(flags)

    # -- halos --
    def halo_bytes(self) -> int:
        return int(self._ck(self.lib.fg_halo_bytes(self.h)))

    def halo_pack(self, face: int) -> np.ndarray:
        buf = np.empty((self.halo_bytes(),), dtype=np.uint8)
        self._ck(self.lib.fg_halo_pack(self.h, face, _ptr(buf)))
        return buf

    def halo_unpack(self, face: int, buf: np.ndarray):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        if buf.nbytes != self.halo_bytes():
            raise ValueError("halo message has the wrong size")
        self._ck(self.lib.fg_halo_unpack(self.h, face, _ptr(buf)))

    def peer_export(self) -> bytes:
        h = FgPeerHandle()
        self._ck(self.lib.fg_peer_export(self.h, C.byref(h)))
        return bytes(h.bytes)

    def peer_connect(self, zlo: Optional[bytes], zhi: Optional[bytes]):
        def mk(b):
            if b is None:
                return None
            h = FgPeerHandle()
            C.memmove(h.bytes, b, 320)
            return C.byref(h)
        self._ck(self.lib.fg_peer_connect(self.h, mk(zlo), mk(zhi)))

    def peer_connect_all(self, handles):
        """handles[r] = peer_export() of rank r, for every rank (immersed bodies may then cross slab faces)."""
        arr = (FgPeerHandle * len(handles))()
        for i, b in enumerate(handles):
            C.memmove(arr[i].bytes, b, 320)
        self._ck(self.lib.fg_peer_connect_all(self.h, arr, len(handles)))
