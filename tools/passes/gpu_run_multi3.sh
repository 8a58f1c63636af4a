# multi-GPU pass 3: culled markers + merged boundary launch
N=${1:-2}
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_multi_$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi_$N.log
run() { G=$1; shift; if [ $G = 1 ]; then timeout 600 python bench.py --gpus 1 "$@"; else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $G "$@"; fi; }
run $N --workload school_1024x512x512 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/m3_school_$N.log 2>&1
run 1 --no-cpu-baseline > gpurun_out/m3_sphere_1.log 2>&1
run $N --no-cpu-baseline > gpurun_out/m3_sphere_$N.log 2>&1
run $N --no-cpu-baseline --no-overlap > gpurun_out/m3_sphere_${N}_nooverlap.log 2>&1
for f in gpurun_out/pytest_multi_$N.log gpurun_out/m3_*; do echo "== $f"; tail -n 2 $f | cut -c1-300; done
