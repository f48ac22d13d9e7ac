N=2
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522"
TADEV_WIDE_COMM=1 TADEV_SUMMA_TRACE=1 timeout 300 $TR bench.py --gpus $N --config C3 --no-cpu --no-e2e --steps 3 --warmup 2 > gpurun_out/r02g_bench_C3_n${N}_wide.json 2> gpurun_out/r02g_trace_C3_n${N}_wide.log
timeout 300 $TR bench.py --gpus $N --config C3 --no-cpu --no-e2e --steps 3 --warmup 2 > gpurun_out/r02g_bench_C3_n$N.json 2> gpurun_out/r02g_bench_C3_n$N.err
grep -h '^{' gpurun_out/r02g_bench_C3_n$N*.json | cut -c1-220; grep -h "communicators" gpurun_out/r02g_trace_C3_n${N}_wide.log | head -2
