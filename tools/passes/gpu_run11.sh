# GPU pass 11: fused even+odd step pairs (StreamCollidePair), stage 1: no bodies, one rank
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_step_pairs or case_table" > gpurun_out/pytest_pair.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pair.log
for wl in box_256 box_512; do
  timeout 300 python bench.py --workload $wl --steps 100 --warmup 10 --no-cpu-baseline --no-pair > gpurun_out/p11_${wl}_nopair.log 2>&1
  for lag in 1 2 3 4 8; do
    timeout 300 python bench.py --workload $wl --steps 100 --warmup 10 --no-cpu-baseline --pair-lag $lag > gpurun_out/p11_${wl}_lag$lag.log 2>&1
  done
done
timeout 300 python bench.py --workload box_256 --steps 100 --warmup 10 --no-cpu-baseline --pair-lag 16 > gpurun_out/p11_box_256_lag16.log 2>&1
tail -n 3 gpurun_out/pytest_pair.log
