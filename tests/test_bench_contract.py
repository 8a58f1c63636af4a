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


def _dry_run(*extra):
    """The b200 arm itself cannot run here (no GPU); FG_BENCH_SHRINK + FG_CUDA_LIB point the same script at the CPU emulation of
    the kernels on lattices 16x smaller — a dry run of the script's control flow, marked as such in its line."""
    emu = os.path.join(ROOT, "tests", "emu", "libfishgym_emu.so")
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "emu")], capture_output=True)
    env = dict(os.environ, FG_BENCH_SHRINK="16", FG_CUDA_LIB=emu)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "4", "--warmup", "3", *extra],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_b200_arm_dry_run_prints_the_contract_keys():
    d = _dry_run()
    assert d["dry_run_shrink"] == 16 and d["n_gpus"] == 1 and d["unit"] == "MLUPS" and d["scaling"] == "weak" and d["vs_baseline"] is None
    for k in ("metric", "value", "ms_per_step", "higher_is_better", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["config"]["name"] == "box_512" and d["gpu_launches"] > 0 and d["value"] > 0
    # the emulation has no event timing of the bulk launches, so the dry run carries no roofline object; on a GPU it must
    assert d["roofline"] is None or {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0 and "own process" in d["cpu_baseline"]["sample"]
    assert set(d["sub_records"]) == {"ib_overhead", "env", "sphere_256x128x128", "ib_multi_direct_forcing"} and "watchdog" not in d
    mdf = d["sub_records"]["ib_multi_direct_forcing"]
    assert mdf["ib_iterations"] == 3 and mdf["value"] > 0 and mdf["markers"] == d["sub_records"]["ib_overhead"]["markers"]
    assert not any("error" in v for v in d["sub_records"].values() if isinstance(v, dict))


def test_b200_arm_watchdog_prints_the_line_when_the_extras_overrun():
    """Sub-records and the CPU baseline are extras: when they exceed --sub-budget-s the main measurement is printed anyway."""
    d = _dry_run("--sub-budget-s", "0.01")
    assert d["value"] > 0 and d["e2e"] is not None and "watchdog" in d
