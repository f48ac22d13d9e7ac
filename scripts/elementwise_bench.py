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
# squared-norm reduction (truncate / true shapes): algorithmic bytes = 8 per element read
import ctypes as C  # noqa: E402
from tiledarray_b200._lib import check  # noqa: E402
for name, tile_elems, ntiles in (("512x512 tiles", 1 << 18, 1638), ("1024x1024 tiles", 1 << 20, 256), ("one 8 MiB tile", 1 << 20, 1)):
    n = tile_elems * ntiles
    x = dev.alloc(n * 8)
    dev.fill_uniform(x, n, 3)
    ptrs = dev.upload(x.ptr + np.arange(ntiles, dtype=np.uint64) * np.uint64(tile_elems * 8))
    sizes = dev.upload(np.full(ntiles, tile_elems, dtype=np.int64))
    o = dev.alloc(8 * ntiles)
    best = 1e9
    for _ in range(4):
        with dev.timer() as tm:
            check(dev.lib.tadev_tile_sqnorms_f64(dev.ctx, dev.stream, ntiles, ptrs.ptr, sizes.ptr, tile_elems, o.ptr))
        best = min(best, tm.ms)
    gbs = n * 8 / (best * 1e-3) / 1e9
    out["cases"][f"sqnorm {name}"] = {"gbs": gbs, "frac_of_copy": gbs / peak, "ms": best}
    print(f"sqnorm {name:18s} {gbs:7.0f} GB/s ({gbs / peak:4.2f} of copy {peak:.0f}; read-only)", flush=True)
    for b in (x, ptrs, sizes, o):
        b.free()
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
dev.close()
