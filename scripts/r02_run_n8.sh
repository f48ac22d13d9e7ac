N=8
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29588"
nvidia-smi topo -m > gpurun_out/r02_topo_n8.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x -k "(test_summa_nccl and not host and not lazy) or (baseline and (c3 or c5s or c4))" > gpurun_out/r02_pytest_multi_n$N.log 2>&1; tail -8 gpurun_out/r02_pytest_multi_n$N.log
for res in 2 4; do
  TADEV_SM_RESERVE=$res timeout 600 $TR bench.py --gpus $N --config C2 --no-cpu --no-e2e --steps 3 --warmup 2 > gpurun_out/r02_bench_C2_n${N}_res$res.json 2> gpurun_out/r02_bench_C2_n${N}_res$res.err; tail -c 300 gpurun_out/r02_bench_C2_n${N}_res$res.err
done
TADEV_SUMMA_TRACE=1 timeout 600 $TR bench.py --gpus $N --config C2 --no-cpu --steps 2 --warmup 2 > gpurun_out/r02_bench_C2_n$N.json 2> gpurun_out/r02_trace_C2_n$N.log
for c in C3 C3m C5; do
  timeout 600 $TR bench.py --gpus $N --config $c --no-cpu --steps 3 --warmup 2 > gpurun_out/r02_bench_${c}_n$N.json 2> gpurun_out/r02_bench_${c}_n$N.err; tail -c 500 gpurun_out/r02_bench_${c}_n$N.err
done
timeout 900 $TR bench.py --gpus $N --config C4 --no-cpu > gpurun_out/r02_bench_C4_n$N.json 2> gpurun_out/r02_bench_C4_n$N.err; tail -c 500 gpurun_out/r02_bench_C4_n$N.err
grep -h '^{' gpurun_out/r02_bench_C*_n$N*.json | cut -c1-260
