"""bench.py contract on a machine without a GPU: the `--impl reference` arm (the fp64 oracle on the host cores)
prints one JSON line with the same metric/unit/config keys as the GPU arm; under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3",
                        "--ref-budget-s", "8"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "MLUPS" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["steps"] == 3 and d["warmup"] == 3 and d["value"] > 0 and d["ms_per_step"] > 0
    # the default workload is the one the metric is quoted on (BASELINE.json configs[4]: 512^3 cells per GPU), and the
    # `config` object is built by the same function as the GPU arm's, so the two arms name the same workload
    import bench
    assert d["config"] == bench.base_config("box_512", 1) and d["config"]["grid_per_gpu_xyz"] == [512, 512, 512]
    assert d["run"]["warmup_steps_run"] >= 20          # the arm is warmed until steady whatever --warmup says
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "3"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
