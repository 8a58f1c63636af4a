# GPU pass 7: full GPU suite after the probe / env additions, final bench line
mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench.log 2>&1
timeout 400 python bench.py --impl reference --steps 50 --warmup 5 > gpurun_out/bench_reference.log 2>&1
for f in pytest_gpu smoke bench bench_reference; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-600; done
