# GPU pass 3: parity tests + bench lines after the addressing rewrite
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 300 --warmup 30 > gpurun_out/bench.log 2>&1
timeout 300 python bench.py --workload box_512 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_512.log 2>&1
timeout 300 python bench.py --workload tank_512x256x256 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_tank.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:StreamCollide -s 6 -c 2 -f -o gpurun_out/prof_sc3 python bench.py --workload box_512 --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_full.log 2>&1
for f in pytest_gpu bench bench_512 bench_tank; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-900; done
