# GPU pass 16: the opt-in 16-bit-storage build (libfishgym_cuda_f16.so): parity tests, throughput on every single-GPU workload,
# vector-env throughput
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_f16_storage.py tests/test_abi.py -m gpu -q > gpurun_out/pytest_f16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_f16.log
timeout 300 python bench.py --storage f16 --no-cpu-baseline > gpurun_out/f16_sphere.log 2>&1
timeout 300 python bench.py --storage f16 --workload tank_512x256x256 --steps 400 --warmup 40 --no-cpu-baseline > gpurun_out/f16_tank.log 2>&1
timeout 300 python bench.py --storage f16 --workload box_512 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/f16_box_512.log 2>&1
timeout 300 python bench.py --storage f16 --workload box_256 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/f16_box_256.log 2>&1
timeout 300 python bench.py --storage f16 --workload box_512_ib --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 20 > gpurun_out/f16_box_512_ib.log 2>&1
timeout 600 python tools/vec_env_bench.py > gpurun_out/vec_env.log 2>&1
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on --kernel-name-base demangled -k regex:StreamCollide -s 6 -c 2 -f -o gpurun_out/prof_f16 python bench.py --storage f16 --workload box_256 --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f16.log 2>&1
tail -n 4 gpurun_out/pytest_f16.log; cat gpurun_out/vec_env.log
