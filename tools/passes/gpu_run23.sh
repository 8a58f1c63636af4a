# GPU pass 23: threads per stream-collide CTA: 128 x 9 (default), 64 x 18, 256 x 4
mkdir -p gpurun_out
set -x
for v in default cta64 cta256; do
  lib=""; [ $v != default ] && lib=$PWD/build/libfg_$v.so
  for wl in sphere_256x128x128 tank_512x256x256 box_512; do
    st=400; [ $wl = box_512 ] && st=100
    FG_CUDA_LIB=$lib timeout 300 python bench.py --workload $wl --steps $st --warmup 40 --no-cpu-baseline > gpurun_out/p23_${wl}_$v.log 2>&1
  done
done
FG_CUDA_LIB=$PWD/build/libfg_cta64.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "case_table or xwall" > gpurun_out/pytest_cta64.log 2>&1; tail -n 2 gpurun_out/pytest_cta64.log
