"""One plain GEMM (single group) for L2/DRAM-traffic experiments: python scripts/gemm_single.py m n k [opA opB]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiledarray_b200 import Device
m, n, k = (int(x) for x in sys.argv[1:4])
opA = int(sys.argv[4]) if len(sys.argv) > 4 else 0
opB = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dev = Device(0)
A, B, Cb = dev.alloc(m * k * 8), dev.alloc(k * n * 8), dev.alloc(m * n * 8)
dev.fill_uniform(A, m * k, 1); dev.fill_uniform(B, k * n, 2)
for _ in range(2):
    with dev.timer() as tm:
        dev.gemm(opA, opB, m, n, k, 1.0, A, B, 0.0, Cb)
    print(f"m={m} n={n} k={k}: {tm.ms:.3f} ms {2.0*m*n*k/tm.ms/1e9:.2f} TF  unique operand bytes {(m*k+k*n)*8/1e6:.0f} MB")
dev.close()
