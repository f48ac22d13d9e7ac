N=4
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545"
TADEV_SUMMA_TRACE=1 timeout 600 $TR bench.py --gpus $N --config C2 --no-cpu --steps 3 --warmup 2 > gpurun_out/r02b_bench_C2_n$N.json 2> gpurun_out/r02b_trace_C2_n$N.log
TADEV_SUMMA_TRACE=1 timeout 600 $TR bench.py --gpus $N --config C3 --no-cpu --no-e2e --steps 3 --warmup 2 > gpurun_out/r02b_bench_C3_n$N.json 2> gpurun_out/r02b_trace_C3_n$N.log
grep -h '^{' gpurun_out/r02b_bench_C*_n$N.json | cut -c1-260
