"""One grouped DGEMM launch for ncu / quick timing:
python scripts/gemm_once.py [tile] [side] [ksteps] [opA] [opB] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiledarray_b200 import Device
tile = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
side = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ks = int(sys.argv[3]) if len(sys.argv) > 3 else 4
opA = int(sys.argv[4]) if len(sys.argv) > 4 else 0
opB = int(sys.argv[5]) if len(sys.argv) > 5 else 0
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 2
dev = Device(0)
T = tile
bufA = dev.alloc(side * ks * T * T * 8); bufB = dev.alloc(side * ks * T * T * 8); bufC = dev.alloc(side * side * T * T * 8)
dev.fill_uniform(bufA, side * ks * T * T, 1); dev.fill_uniform(bufB, side * ks * T * T, 2)
tb = T * T * 8
groups = [(bufC.ptr + (i * side + j) * tb, T, T, 0,
           [(bufA.ptr + (i * ks + k) * tb, bufB.ptr + (k * side + j) * tb, T) for k in range(ks)])
          for i in range(side) for j in range(side)]
packed = dev.make_groups(groups)
best = 1e9
for _ in range(reps):
    with dev.timer() as tm:
        dev.gemm_grouped_packed(opA, opB, 1.0, packed)
    best = min(best, tm.ms)
print(f"tile={tile} side={side} ks={ks} opA={opA} opB={opB}: best {best:.3f} ms  {2.0 * (side * T) ** 2 * ks * T / best / 1e9:.2f} TF")
dev.close()
