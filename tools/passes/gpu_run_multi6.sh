# multi-GPU pass 6: 9 vs 8 CTAs/SM on two z-slabs (default workload)
mkdir -p gpurun_out
set -x
run() { G=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $G "$@"; }
for i in 1 2; do
  run 2 --no-cpu-baseline > gpurun_out/m6_sphere_2_occ9_$i.log 2>&1
  FG_CUDA_LIB=$PWD/build/libfg_occ8.so run 2 --no-cpu-baseline > gpurun_out/m6_sphere_2_occ8_$i.log 2>&1
done
