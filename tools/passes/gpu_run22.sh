# GPU pass 22: cache-operator variants of the population loads / stores (default, ld.cs, ld.cg, st.cs, ld.cs+st.cs, ld.cg+st.cg)
mkdir -p gpurun_out
set -x
for v in default LDCS LDCG STCS LDCSSTCS LDCGSTCG; do
  lib=""; [ $v != default ] && lib=$PWD/build/libfg_$v.so
  for wl in sphere_256x128x128 tank_512x256x256 box_512; do
    st=400; [ $wl = box_512 ] && st=100
    FG_CUDA_LIB=$lib timeout 300 python bench.py --workload $wl --steps $st --warmup 40 --no-cpu-baseline > gpurun_out/p22_${wl}_$v.log 2>&1
  done
done
