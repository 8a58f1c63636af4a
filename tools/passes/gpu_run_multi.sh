# multi-GPU pass: slab tests on N GPUs, weak-scaling bench lines (overlap on/off)
N=${1:-2}
mkdir -p gpurun_out
set -x
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_multi_$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi_$N.log
for W in sphere_256x128x128 box_512; do
for G in 1 $N; do
  if [ $G = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29611"; fi
  timeout 600 $L bench.py --gpus $G --workload $W --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/scale_${W}_$G.log 2>&1
done
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --workload box_512 --steps 200 --warmup 20 --no-cpu-baseline --no-overlap > gpurun_out/scale_box_512_${N}_nooverlap.log 2>&1
for f in gpurun_out/pytest_multi_$N.log gpurun_out/scale_*; do echo "== $f"; tail -n 4 $f | cut -c1-1500; done
