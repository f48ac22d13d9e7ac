set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
TADEV_SUMMA_TRACE=1 $TR bench.py --gpus 4 --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_n4_base.json 2> gpurun_out/r02_n4_base_trace.log
TADEV_SM_RESERVE=2 $TR bench.py --gpus 4 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02_n4_res2.json 2> gpurun_out/r02_n4_res2.err
TADEV_SM_RESERVE=1 $TR bench.py --gpus 4 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02_n4_res1.json 2> gpurun_out/r02_n4_res1.err
$TR bench.py --gpus 4 --steps 2 --warmup 1 --no-cpu --no-e2e --spl 8 > gpurun_out/r02_n4_spl8.json 2> gpurun_out/r02_n4_spl8.err
TADEV_SM_RESERVE=2 $TR bench.py --gpus 4 --steps 2 --warmup 1 --no-cpu --no-e2e --spl 8 > gpurun_out/r02_n4_spl8_res2.json 2> gpurun_out/r02_n4_spl8_res2.err
TADEV_SM_RESERVE=0 $TR bench.py --gpus 4 --steps 2 --warmup 1 --no-cpu --no-e2e --spl 8 > gpurun_out/r02_n4_spl8_res0.json 2> gpurun_out/r02_n4_spl8_res0.err
tail -n 3 gpurun_out/r02_n4_*.json
