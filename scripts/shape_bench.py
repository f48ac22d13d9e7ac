"""SparseShape screening + tile-list kernels at the BASELINE shapes (SURVEY §8 d): run under
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none
to get per-kernel time and DRAM bytes. Algorithmic bytes = 4*(Mt*Kt + Kt*Nt + Mt*Nt).
python scripts/shape_bench.py [json-out]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiledarray_b200 import Device  # noqa: E402

dev = Device(0)
out = {}
rng = np.random.default_rng(0)
CASES = [("C3 128x128x128", 128, 128, 128), ("C4 4x169x169", 4, 169, 169), ("C4-swapped 169x4x169", 169, 4, 169),
         ("big 2048^3", 2048, 2048, 2048)]
for name, Mt, Nt, Kt in CASES:
    a = (rng.random((Mt, Kt)) < 0.1).astype(np.float32) * rng.random((Mt, Kt)).astype(np.float32)
    b = (rng.random((Kt, Nt)) < 0.1).astype(np.float32) * rng.random((Kt, Nt)).astype(np.float32)
    ksz = np.full(Kt, 512.0, dtype=np.float32)
    for _ in range(2):
        c, nz = dev.shape_gemm(a, b, ksz, 1.0, 1e-6)
    pi, pj = dev.build_pairlist(Kt // 2, 1, 1, 0, 0, a, b, c, Mt, Nt, Kt, 1e-6)
    sc, nz2 = dev.shape_scale(a, np.full(Mt, 1 / 512.0, np.float32), np.full(Kt, 1 / 512.0, np.float32), 1e-6)
    out[name] = {"alg_bytes_gemm": 4 * (Mt * Kt + Kt * Nt + Mt * Nt), "nzero": nz, "npairs_mid_step": int(len(pi))}
    print(name, out[name], flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
dev.close()
