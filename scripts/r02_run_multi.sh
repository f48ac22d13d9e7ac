# usage: r02_run_multi.sh <ngpus> [tests]
N=$1
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
if [ "$2" = "tests" ]; then
  timeout 1500 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r02_pytest_multi_n$N.log 2>&1; tail -15 gpurun_out/r02_pytest_multi_n$N.log
fi
for c in C2 C3 C3m; do
  timeout 600 $TR bench.py --gpus $N --config $c --no-cpu --steps 3 --warmup 2 > gpurun_out/r02_bench_${c}_n$N.json 2> gpurun_out/r02_bench_${c}_n$N.err; tail -c 500 gpurun_out/r02_bench_${c}_n$N.err
done
TADEV_SUMMA_TRACE=1 timeout 600 $TR bench.py --gpus $N --config C2 --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/r02_trace_C2_n$N.json 2> gpurun_out/r02_trace_C2_n$N.log
grep '^{' gpurun_out/r02_bench_C*_n$N.json | cut -c1-700
