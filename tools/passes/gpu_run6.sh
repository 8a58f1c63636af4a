# GPU pass 6 (evidence): full GPU test suite, compute-sanitizer memcheck on a subset, ncu traffic capture + launch list
# for the default workload, bench lines for every single-GPU workload
mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "case_table or solid or immersed or two_slabs or swimming" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/memcheck.log
timeout 400 python bench.py > gpurun_out/bench.log 2>&1
timeout 300 python bench.py --workload box_512 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_512.log 2>&1
timeout 300 python bench.py --workload box_512_ib --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 20 > gpurun_out/bench_512_ib.log 2>&1
timeout 300 python bench.py --workload tank_512x256x256 --steps 400 --warmup 40 --no-cpu-baseline > gpurun_out/bench_tank.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 5 > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:StreamCollide -s 8 -c 3 -f -o gpurun_out/prof_default python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_full_default.log 2>&1
for f in pytest_gpu memcheck bench bench_512 bench_512_ib bench_tank; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-400; done
