# Round 2, multi-GPU pass 1 (gpurun --gpus 2): FG_FLAG_WAVEFRONT on peered z-slabs — only worth running if r2_pass2.sh showed a
# gain on one GPU.  N=1 and N=2 lines of the default workload and of 256^3 per GPU, plain and --wavefront.
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631"
for wl in sphere_256x128x128 box_256; do
  for mode in "" "--wavefront"; do
    echo "== $wl N=1 $mode" >> gpurun_out/wave_multi.log
    timeout 300 python bench.py --gpus 1 --workload $wl --steps 600 --warmup 100 --no-cpu-baseline $mode 2>&1 | tail -n 1 >> gpurun_out/wave_multi.log
    echo "== $wl N=2 $mode" >> gpurun_out/wave_multi.log
    timeout 300 $T bench.py --gpus 2 --workload $wl --steps 600 --warmup 100 --no-cpu-baseline $mode 2>&1 | tail -n 1 >> gpurun_out/wave_multi.log
  done
done
FG_TEST_EXPERIMENTS=1 timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; tail -n 2 gpurun_out/pytest_multi.log
python tools/summarize_bench.py gpurun_out/wave_multi.log
