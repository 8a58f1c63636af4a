# GPU pass 17: two rows per CTA (FG_FLAG_ROWS2) A/B
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_rows or case_table" > gpurun_out/pytest_rows2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_rows2.log
for wl in sphere_256x128x128 box_256 tank_512x256x256 box_512; do
  st=400; [ $wl = box_512 ] && st=100
  timeout 300 python bench.py --workload $wl --steps $st --warmup 40 --no-cpu-baseline > gpurun_out/p17_${wl}.log 2>&1
  timeout 300 python bench.py --workload $wl --steps $st --warmup 40 --no-cpu-baseline --rows2 > gpurun_out/p17_${wl}_rows2.log 2>&1
done
tail -n 3 gpurun_out/pytest_rows2.log
