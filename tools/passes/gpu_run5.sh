# GPU pass 5: fused cooperative IB kernel inside per-substep graphs
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 300 --warmup 30 > gpurun_out/bench.log 2>&1
timeout 300 python bench.py --workload box_512_ib --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 20 > gpurun_out/bench_512_ib.log 2>&1
timeout 300 python bench.py --workload tank_512x256x256 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_tank.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 5 > gpurun_out/ncu_bench.log 2>&1
for f in pytest_gpu bench bench_512_ib bench_tank; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-700; done
