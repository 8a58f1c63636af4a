# 8-GPU pass 3 (final code of the session): weak scaling of the default workload at 1/4/8, 512^3 and the school at 8
mkdir -p gpurun_out
set -x
run() { G=$1; shift; if [ $G = 1 ]; then timeout 300 python bench.py --gpus 1 "$@"; else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $G "$@"; fi; }
for G in 1 4 8; do
  run $G --no-cpu-baseline > gpurun_out/s8c_sphere_$G.log 2>&1
done
run 8 --workload box_512 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/s8c_box_8.log 2>&1
run 8 --workload school_1024x512x512 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/s8c_school_8.log 2>&1
