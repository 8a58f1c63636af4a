# multi-GPU pass 5: final code (async fg_step on z-slabs, XWARP, pop_t): tests + scaling lines
N=${1:-2}
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_multi_$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi_$N.log
run() { G=$1; shift; if [ $G = 1 ]; then timeout 600 python bench.py --gpus 1 "$@"; else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $G "$@"; fi; }
run 1 --no-cpu-baseline > gpurun_out/m5_sphere_1.log 2>&1
run $N --no-cpu-baseline > gpurun_out/m5_sphere_$N.log 2>&1
run $N --workload box_512 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/m5_box_$N.log 2>&1
run $N --workload school_1024x512x512 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/m5_school_$N.log 2>&1
run $N --impl reference --steps 5 --warmup 3 > gpurun_out/m5_reference_$N.log 2>&1
tail -n 3 gpurun_out/pytest_multi_$N.log; tail -n 1 gpurun_out/m5_reference_$N.log | cut -c1-200
