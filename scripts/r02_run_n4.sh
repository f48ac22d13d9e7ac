N=4
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x -k "test_summa_nccl or baseline_configs" > gpurun_out/r02_pytest_multi_n$N.log 2>&1; tail -8 gpurun_out/r02_pytest_multi_n$N.log
for res in 4 2 1; do
  TADEV_SM_RESERVE=$res timeout 600 $TR bench.py --gpus $N --config C2 --no-cpu --no-e2e --steps 2 --warmup 1 > gpurun_out/r02_bench_C2_n${N}_res$res.json 2> gpurun_out/r02_bench_C2_n${N}_res$res.err; tail -c 300 gpurun_out/r02_bench_C2_n${N}_res$res.err
done
for c in C3 C3m C5r; do
  timeout 600 $TR bench.py --gpus $N --config $c --no-cpu --steps 3 --warmup 2 > gpurun_out/r02_bench_${c}_n$N.json 2> gpurun_out/r02_bench_${c}_n$N.err; tail -c 500 gpurun_out/r02_bench_${c}_n$N.err
done
TADEV_SUMMA_TRACE=1 timeout 600 $TR bench.py --gpus $N --config C3 --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/r02_trace_C3_n$N.json 2> gpurun_out/r02_trace_C3_n$N.log
TADEV_SUMMA_TRACE=1 timeout 600 $TR bench.py --gpus $N --config C2 --no-cpu --steps 1 --warmup 1 > gpurun_out/r02_trace_C2_n$N.json 2> gpurun_out/r02_trace_C2_n$N.log
grep -h '^{' gpurun_out/r02_bench_C*_n$N*.json | cut -c1-300
