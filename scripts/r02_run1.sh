set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu.log 2>&1; tail -15 gpurun_out/r02_pytest_gpu.log
for c in C1 C3 C3m C5r C4r; do
  timeout 600 python bench.py --config $c --no-cpu > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err; tail -c 600 gpurun_out/r02_bench_$c.err
done
timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/r02_bench_C2.json 2> gpurun_out/r02_bench_C2.err; tail -c 600 gpurun_out/r02_bench_C2.err
timeout 900 python bench.py --config C5 --no-cpu > gpurun_out/r02_bench_C5.json 2> gpurun_out/r02_bench_C5.err; tail -c 600 gpurun_out/r02_bench_C5.err
cat gpurun_out/r02_bench_C*.json | cut -c1-1500
