"""Multi-GPU z-slabs on real hardware: one PROCESS per slab (torchrun), CUDA-IPC peer mapping, halos pushed by peer
stores by the library.  With >= 2 GPUs every process owns one (`gpurun --gpus 2 -- python -m pytest tests -m gpu -k
multi_gpu`); on a box with ONE GPU the two processes share it (`device = local_rank % visible GPUs`): CUDA IPC maps a
lattice of another process on the same device just as well, so the cross-process path — handle export / open, peer
stores, system-scope flags between two contexts — runs under the driver's 1-GPU test box, too (only the wire differs)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return sum(1 for line in out.splitlines() if line.startswith("GPU "))
    except OSError:
        return 0


def _layout():
    """(visible GPUs, processes): one process per GPU with 2 or 4 GPUs, two processes sharing the GPU when there is one."""
    n = gpu_count()
    if n < 1:
        pytest.skip("needs a GPU")
    return n, (2 if n < 4 else 4)


WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import numpy as np, torch, torch.distributed as dist
import gym_fish_b200 as g, util
util.register_oracle(g)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo", rank=rank, world_size=world)
case = {case!r}
P, Wl, IN, OUT = g.BC_PERIODIC, g.BC_WALL, g.BC_INLET, g.BC_OUTLET
kw = dict(nx=40, ny=24, nz=16 * world, tau=0.7, collision=g.MRT, body_force=[1e-4, 0, 2e-4])
if case == "channel":
    kw.update(bc=[P, P, Wl, Wl, IN, OUT], inlet_u=[0, 0.01, 0.04], body_force=[0, 0, 0])
periodic = case in ("periodic", "fish", "spheres")
fish = case == "fish"
if fish:
    kw = dict(nx=24, ny=20, nz=24 * world, tau=0.8, max_markers=4000, max_links=8)
spheres = case == "spheres"
if spheres:      # plane split on every rank: its own moving sphere, far interior planes collide beside the IB kernels
    kw = dict(nx=40, ny=36, nz=64 * world, tau=0.8, collision=g.MRT, max_markers=1000, max_links=1, split_min_cells=1)
s = g.Sim(backend="cuda", n_ranks=world, rank=rank, device=local % {ngpu}, flags={flags}, **kw)
h = kw["nz"] // world
rho, u = util.smooth_fields((kw["nz"], kw["ny"], kw["nx"]))
s.set_fields(rho[rank*h:(rank+1)*h], u[:, rank*h:(rank+1)*h])
handles = [None] * world
dist.all_gather_object(handles, s.peer_export())
lo = handles[(rank - 1) % world] if (rank > 0 or periodic) else None
hi = handles[(rank + 1) % world] if (rank < world - 1 or periodic) else None
if fish:
    s.peer_connect_all(handles)
    s.add_fish(util.fish_desc(g, root=(12, 10, 17)))     # straddles the face at z = 24 and swims
    obs = []
    for it in range(4):
        s.set_action(np.sin(0.5 * it + np.arange(3)))
        s.step(6)
        obs.append(s.get_obs())
    o = torch.from_numpy(np.stack(obs))
    allo = [torch.empty_like(o) for _ in range(world)]
    dist.all_gather(allo, o)
    if rank == 0:
        np.save({out!r} + ".obs.npy", torch.stack(allo).numpy())
elif spheres:
    s.peer_connect(lo, hi)
    dist.barrier()
    for it in range(10):
        zc = 64 * rank + 14.3 + 3.0 * it
        X = util.sphere_markers((20.2, 18.1, zc), 6.0, 450)
        U = np.zeros_like(X); U[:, 2] = 0.02
        s.set_markers(X, U, np.ones(450, np.float32), np.zeros(450, np.int32))
        s.set_link_origins([[20.2, 18.1, zc]])
        s.step(1)
    s.step(5)
    assert s.stats().split_substeps >= 10, s.stats().split_substeps
else:
    s.peer_connect(lo, hi)
    dist.barrier()
    s.step(25)          # many substeps in ONE call: the ranks run free, ordered only by the device-side neighbour flags
    s.step(8)
f = torch.from_numpy(s.get_populations())
parts = [torch.empty_like(f) for _ in range(world)]
dist.all_gather(parts, f)
if rank == 0:
    np.save({out!r}, torch.cat(parts, dim=1).numpy())
dist.barrier(); s.close(); dist.destroy_process_group()
"""


@pytest.mark.parametrize("case", ["periodic", "channel"])
@pytest.mark.parametrize("overlap", [True, False])
def test_multi_gpu_slabs_bit_identical_to_one_gpu(g, cuda, case, overlap, tmp_path):
    n, world = _layout()
    out = str(tmp_path / "f.npy")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, tests=os.path.join(ROOT, "tests"), case=case, out=out, ngpu=n,
                                    flags=0 if overlap else g._abi.FLAG_NO_OVERLAP))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
                        "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    P, Wl, IN, OUT = g.BC_PERIODIC, g.BC_WALL, g.BC_INLET, g.BC_OUTLET
    kw = dict(nx=40, ny=24, nz=16 * world, tau=0.7, collision=g.MRT, body_force=[1e-4, 0, 2e-4])
    if case == "channel":
        kw.update(bc=[P, P, Wl, Wl, IN, OUT], inlet_u=[0, 0.01, 0.04], body_force=[0, 0, 0])
    whole = g.Sim(backend=cuda, **kw)
    rho, u = util.smooth_fields(whole.shape)
    whole.set_fields(rho, u)
    whole.step(33)
    assert np.array_equal(whole.get_populations(), np.load(out))


def test_multi_gpu_fish_across_slab_faces(g, cuda, tmp_path):
    """Bodies crossing slab faces on real GPUs (fg_peer_connect_all): partial marker velocities and link wrenches travel
    by peer stores; every rank integrates the same fish and must end up with bit-identical observations, equal to the
    1-GPU run to round-off."""
    n, world = _layout()
    out = str(tmp_path / "f.npy")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, tests=os.path.join(ROOT, "tests"), case="fish", out=out, flags=0, ngpu=n))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
                        "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    kw = dict(nx=24, ny=20, nz=24 * world, tau=0.8, max_markers=4000, max_links=8)
    whole = g.Sim(backend=cuda, **kw)
    rho, u = util.smooth_fields(whole.shape)              # the workers start from the same moving fluid
    whole.set_fields(rho, u)
    whole.add_fish(util.fish_desc(g, root=(12, 10, 17)))
    ref = []
    for it in range(4):
        whole.set_action(np.sin(0.5 * it + np.arange(3)))
        whole.step(6)
        ref.append(whole.get_obs())
    obs = np.load(out + ".obs.npy")
    for rk in range(world):
        assert np.array_equal(obs[rk], obs[0])
    assert np.abs(obs[0] - np.stack(ref)).max() < 1e-4
    assert np.abs(whole.get_populations() - np.load(out)).max() < 1e-6


@pytest.mark.parametrize("overlap", [True, False])
def test_multi_gpu_plane_split_matches_one_gpu(g, cuda, overlap, tmp_path):
    """Plane split on z-slabs (each rank: far interior planes before it even waits for its neighbours, IB kernels and
    boundary planes at high priority): the fluid must match the unsplit one-GPU run up to the order of the spreading
    atomics."""
    n, world = _layout()
    out = str(tmp_path / "f.npy")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, tests=os.path.join(ROOT, "tests"), case="spheres", out=out, ngpu=n,
                                    flags=0 if overlap else g._abi.FLAG_NO_OVERLAP))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
                        "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    kw = dict(nx=40, ny=36, nz=64 * world, tau=0.8, collision=g.MRT, max_markers=1000 * world, max_links=world, flags=g._abi.FLAG_NO_SPLIT)
    whole = g.Sim(backend=cuda, **kw)
    rho, u = util.smooth_fields(whole.shape)
    whole.set_fields(rho, u)
    for it in range(10):
        Xs, Us, origins = [], [], []
        for rk in range(world):
            zc = 64 * rk + 14.3 + 3.0 * it
            X = util.sphere_markers((20.2, 18.1, zc), 6.0, 450)
            U = np.zeros_like(X); U[:, 2] = 0.02
            Xs.append(X); Us.append(U); origins.append([20.2, 18.1, zc])
        whole.set_markers(np.concatenate(Xs), np.concatenate(Us), np.ones(450 * world, np.float32), np.repeat(np.arange(world), 450).astype(np.int32))
        whole.set_link_origins(origins)
        whole.step(1)
    whole.step(5)
    assert np.abs(whole.get_populations() - np.load(out)).max() < 2e-7
