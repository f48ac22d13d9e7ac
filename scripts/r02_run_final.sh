set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r02_final_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02_final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02_final_bench_C2.json 2> gpurun_out/r02_final_bench_C2.err; tail -c 300 gpurun_out/r02_final_bench_C2.err
timeout 600 python bench.py --config C4 --no-cpu --no-traffic > gpurun_out/r02_final_bench_C4.json 2> gpurun_out/r02_final_bench_C4.err; tail -c 300 gpurun_out/r02_final_bench_C4.err
grep -h '^{' gpurun_out/r02_final_bench_C*.json | cut -c1-300
