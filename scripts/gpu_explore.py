"""First-contact GPU script: peaks, kernel correctness spot checks, kernel timings.

Run on the GPU box:  python scripts/gpu_explore.py > gpurun_out/explore.log
Not a test and not the bench: numbers here steer kernel design (DESIGN.md cites them).
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiledarray_b200 import Device, OP_N, OP_T  # noqa: E402

out = {}
dev = Device(0)


def relerr(x, ref):
    return float(np.linalg.norm(x - ref) / max(np.linalg.norm(ref), 1e-300))


# ---- 1. peaks -------------------------------------------------------------------------------
for kind, name in ((0, "dmma"), (1, "dfma"), (2, "dmma+dfma")):
    best = 0.0
    for it in range(3):
        t, ms = dev.probe_fp64_peak(kind, 40000)
        best = max(best, t)
    out[f"peak_{name}_tflops"] = best
    print(f"probe {name}: {best:.2f} TFLOP/s (last {ms:.2f} ms)", flush=True)
# sustained: 2 s of back-to-back DMMA
t0 = time.time()
vals = []
while time.time() - t0 < 3.0:
    t, ms = dev.probe_fp64_peak(0, 200000)
    vals.append(t)
out["peak_dmma_sustained_tflops"] = float(np.median(vals[len(vals) // 2:]))
print("probe dmma sustained:", out["peak_dmma_sustained_tflops"], vals[:3], vals[-3:], flush=True)
out["copy_gbs"] = dev.probe_copy_gbs(1 << 30, 5)
print("copy GB/s:", out["copy_gbs"], flush=True)

# ---- 2. cuBLAS DGEMM through torch (comparison only) ----------------------------------------
try:
    import torch

    for n in (4096, 8192):
        a = torch.randn(n, n, dtype=torch.float64, device="cuda")
        b = torch.randn(n, n, dtype=torch.float64, device="cuda")
        torch.matmul(a, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(3):
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[f"cublas_dgemm_{n}_tflops"] = 2 * n ** 3 / (best * 1e-3) / 1e12
        print(f"cuBLAS DGEMM {n}: {out[f'cublas_dgemm_{n}_tflops']:.2f} TFLOP/s ({best:.2f} ms)", flush=True)
        del a, b
    torch.cuda.empty_cache()
except Exception as e:  # noqa: BLE001
    print("torch cublas comparison failed:", e)

# ---- 3. tile GEMM correctness -----------------------------------------------------------------
rng = np.random.default_rng(0)


def run_gemm(opA, opB, m, n, k, alpha=1.0, beta=0.0):
    A = rng.uniform(-1, 1, (m, k) if opA == OP_N else (k, m))
    B = rng.uniform(-1, 1, (k, n) if opB == OP_N else (n, k))
    C0 = rng.uniform(-1, 1, (m, n))
    dA, dB, dC = dev.upload(A), dev.upload(B), dev.upload(C0)
    dev.gemm(opA, opB, m, n, k, alpha, dA, dB, beta, dC)
    C = dev.download(dC, np.float64, (m, n))
    ref = alpha * ((A if opA == OP_N else A.T) @ (B if opB == OP_N else B.T)) + beta * C0
    for b in (dA, dB, dC):
        b.free()
    return relerr(C, ref)


worst = 0.0
for (m, n, k) in [(256, 256, 256), (128, 128, 16), (37, 53, 29), (1, 1, 1), (130, 257, 100), (64, 36, 4096 // 64),
                  (3, 5, 7), (512, 512, 512), (100, 100, 1000)]:
    for opA in (OP_N, OP_T):
        for opB in (OP_N, OP_T):
            for beta in (0.0, 1.0):
                e = run_gemm(opA, opB, m, n, k, alpha=1.5, beta=beta)
                worst = max(worst, e)
                if e > 1e-13:
                    print("GEMM MISMATCH", m, n, k, opA, opB, beta, e, flush=True)
out["gemm_worst_relerr"] = worst
print("gemm worst rel err:", worst, flush=True)

# chained group: C = sum_t A_t B_t
m = n = 256
ks = [64, 128, 16, 7, 256]
As = [rng.uniform(-1, 1, (m, k)) for k in ks]
Bs = [rng.uniform(-1, 1, (k, n)) for k in ks]
dAs, dBs = [dev.upload(a) for a in As], [dev.upload(b) for b in Bs]
dC = dev.alloc(m * n * 8)
dev.gemm_grouped(OP_N, OP_N, 1.0, [(dC.ptr, m, n, 0, [(a.ptr, b.ptr, k) for a, b, k in zip(dAs, dBs, ks)])])
C = dev.download(dC, np.float64, (m, n))
ref = sum(a @ b for a, b in zip(As, Bs))
out["gemm_chain_relerr"] = relerr(C, ref)
print("chain rel err:", out["gemm_chain_relerr"], flush=True)

# ---- 4. tile GEMM timing ------------------------------------------------------------------------
def time_grouped(tile, ntiles_side, ksteps, reps=3):
    """C (ntiles_side^2 tiles of tile^2) = sum over ksteps of A(i,k) B(k,j); one launch."""
    T = tile
    nA = ntiles_side * ksteps
    bufA = dev.alloc(nA * T * T * 8)
    bufB = dev.alloc(nA * T * T * 8)
    bufC = dev.alloc(ntiles_side * ntiles_side * T * T * 8)
    dev.fill_uniform(bufA, nA * T * T, 1)
    dev.fill_uniform(bufB, nA * T * T, 2)
    tb = T * T * 8
    groups = []
    for i in range(ntiles_side):
        for j in range(ntiles_side):
            tasks = [(bufA.ptr + (i * ksteps + k) * tb, bufB.ptr + (k * ntiles_side + j) * tb, T) for k in range(ksteps)]
            groups.append((bufC.ptr + (i * ntiles_side + j) * tb, T, T, 0, tasks))
    packed = dev.make_groups(groups)
    dev.gemm_grouped_packed(OP_N, OP_N, 1.0, packed)
    dev.sync()
    best = 1e9
    for _ in range(reps):
        with dev.timer() as tm:
            dev.gemm_grouped_packed(OP_N, OP_N, 1.0, packed)
        best = min(best, tm.ms)
    fl = 2.0 * (ntiles_side * T) ** 2 * (ksteps * T)
    for b in (bufA, bufB, bufC):
        b.free()
    return fl / (best * 1e-3) / 1e12, best


for (tile, side, ks_) in [(256, 16, 16), (512, 16, 16), (1024, 8, 8), (1024, 16, 16), (1024, 16, 1), (512, 16, 1)]:
    tf, ms = time_grouped(tile, side, ks_)
    out[f"gemm_t{tile}_s{side}_k{ks_}_tflops"] = tf
    print(f"grouped gemm tile={tile} side={side} ksteps={ks_}: {tf:.2f} TFLOP/s ({ms:.2f} ms)", flush=True)

# ---- 5. permute -------------------------------------------------------------------------------
def np_permute(x, perm):
    axes = [0] * len(perm)
    for i, p in enumerate(perm):
        axes[p] = i
    return np.ascontiguousarray(np.transpose(x, axes))


ok = True
for ext, perm in [((4, 5, 6), (2, 0, 1)), ((7, 3), (1, 0)), ((5, 4, 3, 2), (0, 3, 2, 1)), ((5, 4, 3, 2), (1, 0, 3, 2)),
                  ((16, 16, 64, 64), (0, 3, 1, 2)), ((16, 64, 16, 64), (2, 0, 1, 3)), ((3, 4, 5, 6, 7, 2), (5, 3, 1, 0, 2, 4)),
                  ((33, 65), (1, 0)), ((2, 3, 4), (0, 1, 2)), ((1, 9, 1, 5), (3, 2, 1, 0))]:
    x = rng.uniform(-1, 1, ext)
    dx = dev.upload(x)
    dy = dev.alloc(x.nbytes)
    dev.permute(ext, perm, 8, dx, dy)
    ref = np_permute(x, perm)
    y = dev.download(dy, np.float64, ref.shape)
    if not np.array_equal(y, ref):
        ok = False
        print("PERMUTE MISMATCH", ext, perm, flush=True)
    dx.free(); dy.free()
out["permute_ok"] = ok
print("permute ok:", ok, flush=True)

for ext, perm in [((16, 16, 64, 64), (0, 3, 1, 2)), ((16, 64, 16, 64), (2, 0, 1, 3)), ((4096, 4096), (1, 0)),
                  ((64, 64, 64, 64), (2, 3, 0, 1)), ((64, 64, 64, 64), (3, 2, 1, 0))]:
    n = int(np.prod(ext))
    reps_buf = max(1, (1 << 30) // (n * 8))  # many tiles back to back > L2
    src = dev.alloc(n * 8 * reps_buf)
    dst = dev.alloc(n * 8 * reps_buf)
    dev.fill_uniform(src, n * reps_buf, 3)
    dev.permute(ext, perm, 8, src, dst)
    dev.sync()
    with dev.timer() as tm:
        for rr in range(reps_buf):
            dev.permute(ext, perm, 8, src.view(rr * n * 8, n * 8), dst.view(rr * n * 8, n * 8))
    gbs = 2.0 * n * 8 * reps_buf / (tm.ms * 1e-3) / 1e9
    out[f"permute_{'x'.join(map(str, ext))}_{''.join(map(str, perm))}_gbs"] = gbs
    print(f"permute {ext} {perm}: {gbs:.0f} GB/s ({tm.ms / reps_buf * 1e3:.1f} us/tile)", flush=True)
    src.free(); dst.free()

# ---- 6. shapes ----------------------------------------------------------------------------------
a = rng.uniform(0, 1, (37, 29)).astype(np.float32)
b = rng.uniform(0, 1, (29, 41)).astype(np.float32)
ksz = rng.integers(1, 9, 29).astype(np.float32)
la = a * ksz[None, :]
rb = b * ksz[:, None]
acc = np.zeros((37, 41), np.float32)
for k in range(29):
    acc = acc + np.outer(la[:, k], rb[k, :]).astype(np.float32)
ref = np.float32(7.2) * acc
thr = np.float32(np.median(ref))
refz = np.where(ref < thr, np.float32(0), ref)
got, nz = dev.shape_gemm(a, b, ksz, 7.2, float(thr))
out["shape_gemm_bitexact"] = bool(np.array_equal(got.view(np.uint32), refz.view(np.uint32))) and nz == int((ref < thr).sum())
print("shape gemm bit-exact:", out["shape_gemm_bitexact"], nz, flush=True)

out["launches"] = dev.launch_count()
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/explore.json", "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
dev.close()
