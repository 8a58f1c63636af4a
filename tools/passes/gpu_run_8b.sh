# 8-GPU pass 2: slab tests at world=4 (incl. plane split), weak scaling of the default workload at 1/2/4/8, 512^3 and school at 8
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_multi_8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi_8.log
run() { G=$1; shift; if [ $G = 1 ]; then timeout 600 python bench.py --gpus 1 "$@"; else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $G "$@"; fi; }
for G in 1 2 4 8; do
  run $G --no-cpu-baseline > gpurun_out/s8b_sphere_$G.log 2>&1
done
run 8 --workload box_512 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/s8b_box_8.log 2>&1
run 8 --workload school_1024x512x512 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/s8b_school_8.log 2>&1
tail -n 3 gpurun_out/pytest_multi_8.log
