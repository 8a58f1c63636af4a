# GPU pass 18: fg_step returns after the wrench event (asynchronous collide tail): full suite, e2e numbers
mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench.log 2>&1
timeout 300 python bench.py --workload tank_512x256x256 --steps 400 --warmup 40 --no-cpu-baseline > gpurun_out/bench_tank.log 2>&1
timeout 300 python bench.py --workload box_512_ib --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 40 > gpurun_out/bench_512_ib.log 2>&1
timeout 300 python bench.py --workload box_256 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_256.log 2>&1
timeout 300 python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown.log 2>&1
for f in pytest_gpu smoke e2e_breakdown; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-300; done
