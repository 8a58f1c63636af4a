"""Gym-style env over the C ABI, on the oracle backend so it runs without a GPU."""
import numpy as np
import pytest


def small_env(g, backend):
    from gym_fish_b200.env import EnvConfig, FishEnv, FishSpec
    fish = FishSpec(links=((8, 2.5), (7, 2.5), (6, 2), (5, 1.5)), root=(12, 10, 12), joint_rate_max=0.02, joint_limit=0.6)
    return FishEnv(EnvConfig(grid=(24, 20, 40), tau=0.8, collision=g.BGK, n_substeps=5, max_episode_steps=4, fish=(fish,)), backend=backend)


def test_reset_step_contract_and_determinism(g):
    env = small_env(g, "oracle")
    obs, info = env.reset(seed=3)
    assert obs.dtype == np.float32 and obs.shape == env.observation_space.shape == (14,)
    assert env.action_space.shape == (3,) and info["n_markers"] > 100
    traj = []
    for t in range(4):
        a = env.action_space.sample()
        assert env.action_space.contains(a)
        obs, r, term, trunc, info = env.step(a)
        assert isinstance(r, float) and isinstance(term, bool) and isinstance(trunc, bool)
        assert np.isfinite(obs).all() and not info["diverged"]
        traj.append((obs.copy(), r))
    assert trunc                                        # max_episode_steps reached
    obs2, _ = env.reset(seed=3)
    for t in range(4):
        o, r, *_ = env.step(env.action_space.sample())  # same seed -> same actions -> identical trajectory
        assert np.array_equal(o, traj[t][0]) and r == traj[t][1]
    env.close()


def test_env_on_emulated_cuda_runtime_matches_oracle(g, emu):
    a, b = small_env(g, "oracle"), small_env(g, emu)
    a.reset(seed=1); b.reset(seed=1)
    for t in range(3):
        act = np.sin(t + np.arange(3)).astype(np.float32)
        oa, ra, *_ = a.step(act)
        ob, rb, *_ = b.step(act)
        assert np.abs(oa - ob).max() < 1e-4 and abs(ra - rb) < 1e-5
    a.close(); b.close()


@pytest.mark.gpu
def test_env_on_gpu_matches_oracle(g, cuda):
    a, b = small_env(g, "oracle"), small_env(g, cuda)
    a.reset(seed=1); b.reset(seed=1)
    for t in range(4):
        act = np.sin(t + np.arange(3)).astype(np.float32)
        oa, ra, *_ = a.step(act)
        ob, rb, *_ = b.step(act)
        assert np.abs(oa - ob).max() < 1e-4 and abs(ra - rb) < 1e-5
    a.close(); b.close()
