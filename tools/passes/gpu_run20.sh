# GPU pass 20 (final evidence of the session, HEAD): full GPU suite, smoke, bench lines, reference arm, launch list, memcheck subset
mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench.log 2>&1
timeout 400 python bench.py --impl reference --steps 50 --warmup 5 > gpurun_out/bench_reference.log 2>&1
timeout 300 python bench.py --workload tank_512x256x256 --steps 400 --warmup 40 --no-cpu-baseline > gpurun_out/bench_tank.log 2>&1
timeout 300 python bench.py --workload box_512 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_512.log 2>&1
timeout 300 python bench.py --workload box_512_ib --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 40 > gpurun_out/bench_512_ib.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --kernel-name-base demangled --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 5 > gpurun_out/ncu_bench.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_f16_storage.py -m gpu -q -x -k "plane_split_matches or immersed or fused_step_pairs or xwall or swimming or step_returns or f16_storage_case" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/memcheck.log
for f in pytest_gpu smoke memcheck; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-300; done
