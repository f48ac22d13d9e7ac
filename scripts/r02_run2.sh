set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/r02b_pytest_gpu.log 2>&1; tail -30 gpurun_out/r02b_pytest_gpu.log
for c in C1 C3 C3m; do
  for hl in 0 1; do
    TADEV_HOST_LISTS=$hl timeout 600 python bench.py --config $c --no-cpu --no-e2e --steps 5 > gpurun_out/r02b_bench_${c}_hl$hl.json 2> gpurun_out/r02b_bench_${c}_hl$hl.err; tail -c 300 gpurun_out/r02b_bench_${c}_hl$hl.err
  done
done
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02b_bench_C2.json 2> gpurun_out/r02b_bench_C2.err; tail -c 300 gpurun_out/r02b_bench_C2.err
cat gpurun_out/r02b_bench_*.json | cut -c1-400
