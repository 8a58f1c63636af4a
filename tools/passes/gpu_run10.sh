# GPU pass 10: direct launches for split substeps (default), e2e breakdown, IB kernel list
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench.log 2>&1
timeout 300 python bench.py --workload tank_512x256x256 --steps 400 --warmup 40 --no-cpu-baseline > gpurun_out/bench_tank.log 2>&1
timeout 300 python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown.log 2>&1
timeout 300 python tools/e2e_breakdown.py --no-split >> gpurun_out/e2e_breakdown.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:Ib -c 60 --csv --log-file gpurun_out/ib_launches.csv python bench.py --workload box_512_ib --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_ib.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; cat gpurun_out/e2e_breakdown.log
