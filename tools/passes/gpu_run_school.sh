# school strong scaling (BASELINE configs[3]) on 1 and N GPUs
N=${1:-2}
mkdir -p gpurun_out
set -x
run() { G=$1; shift; if [ $G = 1 ]; then timeout 600 python bench.py --gpus 1 "$@"; else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $G "$@"; fi; }
for G in 1 $N; do
  run $G --workload school_1024x512x512 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/school_$G.log 2>&1
done
for f in gpurun_out/school_*; do echo "== $f"; tail -n 3 $f | cut -c1-900; done
