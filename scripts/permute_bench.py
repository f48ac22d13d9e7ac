"""HBM-bandwidth measurement of the tile permutation kernels (SURVEY §8 d, config 5 permutes).
Algorithmic bytes = 2 * tile bytes (read + write each element once). Operands > L2.
python scripts/permute_bench.py [json-out]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiledarray_b200 import Device  # noqa: E402

dev = Device(0)
peak = dev.probe_copy_gbs(1 << 30, 5)
out = {"copy_gbs": peak, "tile_env": os.environ.get("TADEV_PERM_TILE", "default"), "cases": {}}
CASES = [
    # (extent, perm, what)
    ((16, 16, 64, 64), (0, 3, 1, 2), "C5 A tile (i,k,a,c)->(i,a,c,k): tiled transpose"),
    ((16, 64, 16, 64), (2, 0, 1, 3), "C5 B tile (j,c,k,b)->(c,k,j,b): row copy"),
    ((64, 64, 64, 64), (2, 3, 0, 1), "C4 result tile (i,j,a,b)->(a,b,i,j): 4096x4096 transpose"),
    ((4096, 4096), (1, 0), "matrix transpose"),
    ((64, 64, 64, 64), (3, 2, 1, 0), "full reversal"),
    ((36, 64, 64, 36), (2, 3, 0, 1), "ragged C4 result tile"),
]
for ext, perm, what in CASES:
    n = int(np.prod(ext))
    ntiles = max(1, (1 << 31) // (n * 8))  # 2 GiB per side
    src = dev.alloc(n * 8 * ntiles)
    dst = dev.alloc(n * 8 * ntiles)
    dev.fill_uniform(src, n * ntiles, 3)
    srcs = [src.view(i * n * 8, n * 8) for i in range(ntiles)]
    dsts = [dst.view(i * n * 8, n * 8) for i in range(ntiles)]
    dev.permute_batched(ext, perm, 8, srcs, dsts)
    dev.sync()
    best = 1e9
    for _ in range(3):
        with dev.timer() as tm:
            dev.permute_batched(ext, perm, 8, srcs, dsts)
        best = min(best, tm.ms)
    gbs = 2.0 * n * 8 * ntiles / (best * 1e-3) / 1e9
    with dev.timer() as tm1:
        for i in range(min(ntiles, 64)):
            dev.permute(ext, perm, 8, srcs[i], dsts[i])
    gbs1 = 2.0 * n * 8 * min(ntiles, 64) / (tm1.ms * 1e-3) / 1e9
    key = "x".join(map(str, ext)) + "_" + "".join(map(str, perm))
    out["cases"][key] = {"what": what, "batched_gbs": gbs, "frac_of_copy": gbs / peak, "per_tile_launch_gbs": gbs1, "ntiles": ntiles}
    print(f"{key:28s} batched {gbs:7.0f} GB/s ({gbs / peak:5.2f} of copy {peak:.0f})   one-launch-per-tile {gbs1:7.0f} GB/s   [{what}]", flush=True)
    src.free()
    dst.free()
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
dev.close()
