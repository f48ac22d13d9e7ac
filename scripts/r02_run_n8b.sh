N=8
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29589"
TADEV_SUMMA_TRACE=1 timeout 300 $TR bench.py --gpus $N --config C3 --no-cpu --no-e2e --steps 3 --warmup 2 > gpurun_out/r02h_bench_C3_n${N}_auto.json 2> gpurun_out/r02h_trace_C3_n${N}_auto.log
TADEV_WIDE_COMM=0 timeout 300 $TR bench.py --gpus $N --config C3 --no-cpu --no-e2e --steps 3 --warmup 2 > gpurun_out/r02h_bench_C3_n${N}_thin.json 2> gpurun_out/r02h_bench_C3_n${N}_thin.err
timeout 400 $TR bench.py --gpus $N --config C2 --no-cpu --steps 3 --warmup 2 > gpurun_out/r02h_bench_C2_n$N.json 2> gpurun_out/r02h_bench_C2_n$N.err
grep -h '^{' gpurun_out/r02h_bench_C*_n$N*.json | cut -c1-200
