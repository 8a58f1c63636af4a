# GPU pass 21: launch bounds of the stream-collide kernels: 8 (64 regs), 9 (56 regs, no spills), 10 CTAs/SM (48 regs, spills)
mkdir -p gpurun_out
set -x
for occ in 8 9 10; do
  lib=""; [ $occ != 8 ] && lib=$PWD/build/libfg_occ$occ.so
  for wl in sphere_256x128x128 box_256 tank_512x256x256 box_512; do
    st=400; [ $wl = box_512 ] && st=100
    FG_CUDA_LIB=$lib timeout 300 python bench.py --workload $wl --steps $st --warmup 40 --no-cpu-baseline > gpurun_out/p21_${wl}_occ$occ.log 2>&1
  done
done
FG_CUDA_LIB=$PWD/build/libfg_occ9.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "case_table or taylor or plane_split_matches" > gpurun_out/pytest_occ9.log 2>&1; tail -n 2 gpurun_out/pytest_occ9.log
