"""BASELINE configs[0] on the host CPU, the way the reference runs it: `ta_dense 4096 256` (examples/gemm/
ta_dense.cpp) = c("m,n") = a("m,k") * b("k,n") with a.fill(1), b.fill(1), N=4096, tile=256, 5 repetitions, mean and
median wall time (util/time.h). The reference's algorithm is restated by the CPU oracle: SUMMA on a 1x1 grid, one
single-threaded vendor DGEMM per tile pair (tiledarray.cpp:112), one task thread per host core (MAD_NUM_THREADS).
python scripts/cpu_baseline_c1.py [json-out]   (no GPU needed; run on the GPU box's host for the record)"""
import json
import os
import statistics
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cpu as ocpu  # noqa: E402

N, T, REPEAT = 4096, 256, 5
nt = N // T
cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
ocpu.load()
ones = np.ones((T, T))
a_tiles = {(i, k): ones.copy() for i in range(nt) for k in range(nt)}
b_tiles = {(k, j): ones.copy() for k in range(nt) for j in range(nt)}
gflop = 2.0 * N ** 3 / 1e9
times = []
for rep in range(REPEAT + 1):  # one untimed warm-up (thread pool, page faults), like a first TA iteration
    out, secs, npairs = ocpu.cpu_contract(a_tiles, b_tiles, [T] * nt, [T] * nt, [T] * nt, 0, 0, 1.0, None, cores)
    if rep:
        times.append(secs)
        print(f"Iteration {rep}   time={secs:.6f}   GFLOPS={gflop / secs:.1f}", flush=True)
assert all(np.all(t == float(N)) for t in out.values()) and npairs == nt ** 3
mean, median = statistics.mean(times), statistics.median(times)
rec = {"config": "C1 dense N=4096 tile=256 on the host CPU (reference algorithm, oracle port)", "cores": cores, "repeat": REPEAT,
       "mean_s": mean, "median_s": median, "mean_gflops": gflop * statistics.mean(1.0 / t for t in times),
       "median_gflops": gflop / median, "tile_pairs": npairs}
print(json.dumps(rec))
if len(sys.argv) > 1:
    json.dump(rec, open(sys.argv[1], "w"), indent=1)
