"""HBM-bandwidth measurement of the batched element-wise tile kernel (tadev_tiles_binary_f64).
Algorithmic bytes per element: 8 * (2 read + 1 written). python scripts/elementwise_bench.py [json-out]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiledarray_b200 import Device, _lib  # noqa: E402

dev = Device(0)
peak = dev.probe_copy_gbs(1 << 30, 5)
out = {"copy_gbs": peak, "cases": {}}
for name, tile_elems, ntiles in (("1024x1024 tiles", 1 << 20, 256), ("64^3 tiles", 1 << 18, 1024), ("one 2 GiB tile", 1 << 28, 1)):
    n = tile_elems * ntiles
    x, y, z = dev.alloc(n * 8), dev.alloc(n * 8), dev.alloc(n * 8)
    dev.fill_uniform(x, n, 1)
    dev.fill_uniform(y, n, 2)
    offs = np.arange(ntiles, dtype=np.uint64) * np.uint64(tile_elems * 8)
    elems = np.full(ntiles, tile_elems, dtype=np.int64)
    for op, opname in ((_lib.EW_AXPBY, "axpby"), (_lib.EW_MULT, "mult")):
        dev.tiles_binary(op, z.ptr + offs, x.ptr + offs, y.ptr + offs, elems, 2.0, -1.0)
        dev.sync()
        best = 1e9
        for _ in range(3):
            with dev.timer() as tm:
                dev.tiles_binary(op, z.ptr + offs, x.ptr + offs, y.ptr + offs, elems, 2.0, -1.0)
            best = min(best, tm.ms)
        gbs = 3.0 * n * 8 / (best * 1e-3) / 1e9
        out["cases"][f"{opname} {name}"] = {"gbs": gbs, "frac_of_copy": gbs / peak, "ms": best}
        print(f"{opname:6s} {name:18s} {gbs:7.0f} GB/s ({gbs / peak:4.2f} of copy {peak:.0f})", flush=True)
    for b in (x, y, z):
        b.free()
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
dev.close()
