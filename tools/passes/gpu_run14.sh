# GPU pass 14: x-wall warp-uniform variant A/B on the tank, full GPU suite
mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for i in 1 2; do
timeout 300 python bench.py --workload tank_512x256x256 --steps 400 --warmup 40 --no-cpu-baseline > gpurun_out/p14_tank_$i.log 2>&1
timeout 300 python bench.py --workload tank_512x256x256 --steps 400 --warmup 40 --no-cpu-baseline --xwarp > gpurun_out/p14_tank_xwarp_$i.log 2>&1
done
tail -n 4 gpurun_out/pytest_gpu.log
