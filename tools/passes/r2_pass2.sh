# Round 2, GPU pass 2 (one GPU): first measurement of FG_FLAG_WAVEFRONT (sim.hpp wave_pair) — even + odd step as a
# launch-level wavefront of plane chunks on two streams.  A/B per workload: default vs --wavefront at several chunk sizes,
# then one ncu capture of DRAM bytes per launch for the best chunk size (the point of the wavefront is traffic < 152 B/update).
mkdir -p gpurun_out
B="python bench.py --steps 600 --warmup 100 --no-cpu-baseline"
for wl in sphere_256x128x128 box_256 box_512 tank_512x256x256; do
  echo "== $wl default" >> gpurun_out/wave.log
  timeout 200 $B --workload $wl 2>&1 | tail -n 1 >> gpurun_out/wave.log
  for lag in 0 8 11 16 21 32; do   # 126 bulk CTAs per plane, 1332 resident: 10.6 planes per wave
    echo "== $wl --wavefront --pair-lag $lag" >> gpurun_out/wave.log
    timeout 200 $B --workload $wl --wavefront --pair-lag $lag 2>&1 | tail -n 1 >> gpurun_out/wave.log
  done
  echo "== $wl --wavefront --no-graphs" >> gpurun_out/wave.log
  timeout 200 $B --workload $wl --wavefront --no-graphs 2>&1 | tail -n 1 >> gpurun_out/wave.log
done
FG_TEST_EXPERIMENTS=1 timeout 300 python -m pytest tests/test_random_cases.py -m gpu -q -k wavefront > gpurun_out/pytest_wave.log 2>&1
# DRAM traffic of one pair with and without the wavefront (serialised under ncu: only the byte counts are meaningful)
for mode in "" "--wavefront"; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none \
    -k regex:StreamCollide -c 80 --csv --log-file gpurun_out/ncu_wave_${mode:-plain}.csv $B --workload box_256 --steps 8 --warmup 4 $mode > /dev/null 2>&1
done
python tools/summarize_bench.py gpurun_out/wave.log 2>/dev/null | tail -n 40; tail -n 3 gpurun_out/pytest_wave.log
