set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/r02e_pytest_gpu.log 2>&1; tail -12 gpurun_out/r02e_pytest_gpu.log
LD_LIBRARY_PATH=tiledarray_b200 tests/cpp/build/test_tile_plugin 2>&1 | tail -4
KF='regex:gemm|tl_|probe|shape|zero|sqnorm|transpose|rowcopy|tiles_binary'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" -c 400 --csv --log-file gpurun_out/r02_launches_C2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-traffic > gpurun_out/r02_launches_C2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" -c 400 --csv --log-file gpurun_out/r02_launches_C3.csv python bench.py --config C3 --steps 2 --warmup 1 --no-e2e --no-cpu --no-traffic > gpurun_out/r02_launches_C3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" -c 400 --csv --log-file gpurun_out/r02_launches_C5r.csv python bench.py --config C5r --steps 1 --warmup 1 --no-e2e --no-cpu --no-traffic > gpurun_out/r02_launches_C5r.log 2>&1
tail -2 gpurun_out/r02_launches_C2.log | cut -c1-200
