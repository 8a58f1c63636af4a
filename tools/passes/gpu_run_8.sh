# 8-GPU pass: slab tests at world=4, weak scaling 1 vs 8 (default workload, 512^3), school strong scaling at 8
mkdir -p gpurun_out
set -x
nvidia-smi topo -m > gpurun_out/topo_8.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_multi_8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi_8.log
run() { G=$1; shift; if [ $G = 1 ]; then timeout 600 python bench.py --gpus 1 "$@"; else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $G "$@"; fi; }
for G in 1 4 8; do
  run $G --no-cpu-baseline > gpurun_out/s8_sphere_$G.log 2>&1
done
for G in 1 8; do
  run $G --workload box_512 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/s8_box_$G.log 2>&1
done
run 8 --workload box_512 --steps 100 --warmup 10 --no-cpu-baseline --no-overlap > gpurun_out/s8_box_8_nooverlap.log 2>&1
run 8 --workload school_1024x512x512 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/s8_school_8.log 2>&1
for f in gpurun_out/pytest_multi_8.log gpurun_out/s8_*; do echo "== $f"; tail -n 2 $f | cut -c1-400; done
