# Round 2, GPU pass 1 (one GPU): everything that was written in the GPU-less tail of round 1 gets its first run here.
#   1. full GPU suite (new: Ghia cavity on the CUDA path, slab that runs ahead -> FG_EPEER with the pinned time-out word)
#   2. smoke, default bench line, reference arm
#   3. e2e breakdown: where the 10 us between `value` (fg_step(K)) and `e2e` (send + step(1) + read) go — device time per
#      call vs wall per step, with / without re-sending markers, with FG_FLAG_SYNC_STEP and without the plane split
#   4. launch list of the default bench command (ncu, serialised)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench.log 2>&1
timeout 300 python bench.py --impl reference --steps 30 --warmup 3 > gpurun_out/bench_reference.log 2>&1
for f in "" "--sync-step" "--no-split" "--no-graphs"; do
  timeout 200 python tools/e2e_breakdown.py $f >> gpurun_out/e2e_breakdown.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 20 --warmup 5 > gpurun_out/bench_under_ncu.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/smoke.log; tail -n 1 gpurun_out/bench.log; cat gpurun_out/e2e_breakdown.log
