# GPU pass 13: StreamCollidePair v3 (static wavefront order, warp-level completion counters)
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fused_step_pairs or case_table or solid" > gpurun_out/pytest_pair.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pair.log
for wl in box_256 box_512; do
  timeout 300 python bench.py --workload $wl --steps 100 --warmup 10 --no-cpu-baseline --no-pair > gpurun_out/p13_${wl}_nopair.log 2>&1
  for lag in 2 3 4 6; do
    timeout 300 python bench.py --workload $wl --steps 100 --warmup 10 --no-cpu-baseline --pair-lag $lag > gpurun_out/p13_${wl}_lag$lag.log 2>&1
  done
done
tail -n 4 gpurun_out/pytest_pair.log
