"""One grouped DGEMM launch for ncu: python scripts/gemm_once.py [tile] [side] [ksteps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiledarray_b200 import Device, OP_N
tile = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
side = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ks = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = Device(0)
T = tile
bufA = dev.alloc(side * ks * T * T * 8); bufB = dev.alloc(side * ks * T * T * 8); bufC = dev.alloc(side * side * T * T * 8)
dev.fill_uniform(bufA, side * ks * T * T, 1); dev.fill_uniform(bufB, side * ks * T * T, 2)
tb = T * T * 8
groups = [(bufC.ptr + (i * side + j) * tb, T, T, 0,
           [(bufA.ptr + (i * ks + k) * tb, bufB.ptr + (k * side + j) * tb, T) for k in range(ks)])
          for i in range(side) for j in range(side)]
packed = dev.make_groups(groups)
for _ in range(2):
    with dev.timer() as tm:
        dev.gemm_grouped_packed(OP_N, OP_N, 1.0, packed)
    print("ms", tm.ms, "TF", 2.0 * (side * T) ** 2 * ks * T / tm.ms / 1e9)
dev.close()
