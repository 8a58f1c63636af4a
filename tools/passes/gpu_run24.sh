# GPU pass 24: last sanity of HEAD (full GPU suite, smoke, default bench line)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 200 python bench.py > gpurun_out/bench.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/smoke.log
