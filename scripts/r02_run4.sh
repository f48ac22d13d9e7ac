set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/r02d_pytest_gpu.log 2>&1; tail -12 gpurun_out/r02d_pytest_gpu.log
LD_LIBRARY_PATH=tiledarray_b200 tests/cpp/build/test_tile_plugin 2>&1 | tail -4
for c in C1 C2 C3; do
  timeout 900 python bench.py --config $c > gpurun_out/r02d_bench_$c.json 2> gpurun_out/r02d_bench_$c.err; tail -c 300 gpurun_out/r02d_bench_$c.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_C2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-traffic > gpurun_out/r02_launches_C2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_C3.csv python bench.py --config C3 --steps 2 --warmup 1 --no-e2e --no-cpu --no-traffic > gpurun_out/r02_launches_C3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_grouped_f64_ws -c 1 -o gpurun_out/r02_gemm_C3 python bench.py --config C3 --steps 1 --warmup 0 --no-e2e --no-cpu --no-traffic --parity-tiles 0 > gpurun_out/r02_gemm_C3_ncu.log 2>&1
ncu --set full --clock-control none -k regex:tl_ -c 12 -o gpurun_out/r02_tilelist_C3 python bench.py --config C3 --steps 1 --warmup 0 --no-e2e --no-cpu --no-traffic --parity-tiles 0 > gpurun_out/r02_tilelist_C3_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
grep -h '^{' gpurun_out/r02d_bench_C*.json | cut -c1-200
