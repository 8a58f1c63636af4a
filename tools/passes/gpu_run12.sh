mkdir -p gpurun_out
set -x
timeout 300 python tools/debug_pair.py > gpurun_out/debug_pair.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fused_step_pairs" > gpurun_out/pytest_pair.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pair.log
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on --kernel-name-base demangled -k regex:kern_ticket -s 4 -c 2 -f -o gpurun_out/prof_pair python bench.py --workload box_256 --steps 20 --warmup 4 --no-cpu-baseline --pair-lag 4 > gpurun_out/ncu_pair.log 2>&1
timeout 600 ncu --set full --clock-control none --cache-control none --kernel-name-base demangled -k regex:StreamCollide -s 8 -c 2 -f -o gpurun_out/prof_nopair python bench.py --workload box_256 --steps 20 --warmup 4 --no-cpu-baseline --no-pair > gpurun_out/ncu_nopair.log 2>&1
cat gpurun_out/debug_pair.log; tail -n 5 gpurun_out/pytest_pair.log
