# multi-GPU pass 2: slab + IB-across-slab tests on N GPUs, school strong scaling, default/box weak scaling with graphs
N=${1:-2}
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_multi_$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi_$N.log
run() { G=$1; shift; if [ $G = 1 ]; then timeout 600 python bench.py --gpus 1 "$@"; else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $G "$@"; fi; }
for G in 1 $N; do
  run $G --workload school_1024x512x512 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/school_$G.log 2>&1
  run $G --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/scale_sphere_$G.log 2>&1
  run $G --workload box_512 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/scale_box_$G.log 2>&1
done
for f in gpurun_out/pytest_multi_$N.log gpurun_out/school_* gpurun_out/scale_sphere_* gpurun_out/scale_box_*; do echo "== $f"; tail -n 4 $f | cut -c1-600; done
