# GPU pass 8: plane split (far-plane collide beside the IB kernels, priority streams) A/B, new split tests
mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench.log 2>&1
timeout 300 python bench.py --no-split --no-cpu-baseline > gpurun_out/bench_nosplit.log 2>&1
timeout 300 python bench.py --no-graphs --no-cpu-baseline > gpurun_out/bench_nographs.log 2>&1
timeout 300 python bench.py --no-graphs --no-split --no-cpu-baseline > gpurun_out/bench_nographs_nosplit.log 2>&1
timeout 300 python bench.py --workload tank_512x256x256 --steps 400 --warmup 40 --no-cpu-baseline > gpurun_out/bench_tank.log 2>&1
timeout 300 python bench.py --workload tank_512x256x256 --steps 400 --warmup 40 --no-cpu-baseline --no-split > gpurun_out/bench_tank_nosplit.log 2>&1
timeout 300 python bench.py --workload box_512_ib --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 20 > gpurun_out/bench_512_ib.log 2>&1
for f in pytest_gpu smoke bench bench_nosplit bench_nographs bench_nographs_nosplit bench_tank bench_tank_nosplit bench_512_ib; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-330; done
