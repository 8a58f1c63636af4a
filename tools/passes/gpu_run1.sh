# first GPU pass: smoke, GPU parity tests, bench lines, ncu launch list + full capture of the stream-collide kernel
mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|Socket|Thread|Core" >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 300 --warmup 30 > gpurun_out/bench.log 2>&1
timeout 300 python bench.py --workload box_512 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_512.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 5 > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:StreamCollide -s 6 -c 4 -f -o gpurun_out/prof_sc python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/smoke.log gpurun_out/pytest_gpu.log gpurun_out/bench.log gpurun_out/bench_512.log
