# multi-GPU pass 7: peered handles launch the 8-CTA/SM instantiation (default build): N=1 and N=2 lines + slab tests
mkdir -p gpurun_out
set -x
timeout 300 python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/m7_sphere_1.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/m7_sphere_2.log 2>&1
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "bit_identical or plane_split" > gpurun_out/pytest_multi_2.log 2>&1; tail -n 2 gpurun_out/pytest_multi_2.log
