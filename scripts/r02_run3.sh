set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/r02c_pytest_gpu.log 2>&1; tail -30 gpurun_out/r02c_pytest_gpu.log
LD_LIBRARY_PATH=tiledarray_b200 tests/cpp/build/test_tile_plugin > gpurun_out/r02c_tile_plugin.log 2>&1; tail -5 gpurun_out/r02c_tile_plugin.log
TADEV_SUMMA_TRACE=1 timeout 600 python bench.py --config C3 --no-cpu --no-e2e --steps 2 --warmup 2 > gpurun_out/r02c_bench_C3_trace.json 2> gpurun_out/r02c_bench_C3_trace.err
TADEV_HOST_LISTS=1 TADEV_SUMMA_TRACE=1 timeout 600 python bench.py --config C3 --no-cpu --no-e2e --steps 2 --warmup 2 > gpurun_out/r02c_bench_C3_trace_hl1.json 2> gpurun_out/r02c_bench_C3_trace_hl1.err
grep "host:\|tables_up\|gemm_done" gpurun_out/r02c_bench_C3_trace.err | tail -12
grep "host:\|tables_up\|gemm_done" gpurun_out/r02c_bench_C3_trace_hl1.err | tail -8
python scripts/elementwise_bench.py gpurun_out/r02c_elementwise.json 2>&1 | tail -12
