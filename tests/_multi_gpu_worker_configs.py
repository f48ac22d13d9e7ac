"""Worker of tests/test_gpu_multi.py::test_baseline_configs_multi_rank: miniatures of BASELINE configs 3, 4 and 5 on
N GPUs, operands created by tiledarray.contraction_arrays (already in SUMMA's distribution: no redistribution), every
local result tile compared in full with a host einsum of the regenerated operands.
usage: _multi_gpu_worker_configs.py <case>     case in {c3, c3m, c4, c5, c5s, general, generalp}
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import ta_oracle as O  # noqa: E402
from tests import util_rng  # noqa: E402
from tiledarray_b200 import Device  # noqa: E402
from tiledarray_b200.tiledarray import (ContEngine, DistArray, SparseShape, TiledRange, TiledRange1, World,  # noqa: E402
                                        contraction_arrays)


def host_full(tr, seed, shape=None):
    out = np.zeros(tr.elements_shape)
    for o in range(tr.ntiles):
        if shape is not None and shape.is_zero(o):
            continue
        idx = tr.tile_index(o)
        ext = tr.tile_extent(idx)
        out[tr.tile_slices(idx)] = util_rng.tile_fill(o, int(np.prod(ext)), seed).reshape(ext)
    return out


def main():
    case = sys.argv[1]
    rank, size, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world = World(device=Device(local), rank=rank, size=size)
    mask = None
    shA = shB = None
    lazy = (None, None)
    ContEngine.stream_permutes = "auto"
    if case in ("c3", "c3m"):
        t = TiledRange1.make_uniform(24 * 32, 32)
        trA = trB = trC = TiledRange([t, t])
        target, lidx, ridx = "m,n", "m,k", "k,n"
        rng = np.random.default_rng(5)
        nA = np.where(rng.random((24, 24)) < 0.15, 32.0, 0.0).astype(np.float32)
        nB = np.where(rng.random((24, 24)) < 0.15, 32.0, 0.0).astype(np.float32)
        shA, shB = SparseShape(world, nA, trA), SparseShape(world, nB, trB)
        if case == "c3m":
            mask = shA
    elif case == "c4":  # R(a,b,i,j) = T(c,d,i,j) * V(a,b,c,d), V lazy; the engine exchanges the operands
        o1, v1 = TiledRange1(0, 6, 10), TiledRange1(0, 8, 16, 24, 28)
        trA, trB, trC = TiledRange([v1, v1, o1, o1]), TiledRange([v1, v1, v1, v1]), TiledRange([v1, v1, o1, o1])
        target, lidx, ridx = "a,b,i,j", "c,d,i,j", "a,b,c,d"
        lazy = (None, 808)
    elif case in ("general", "generalp"):  # fused (batch) index b: one SUMMA per slab and batch element on the shared grid
        bt, d1, d2 = TiledRange1(0, 2, 5), TiledRange1.make_uniform(24, 8), TiledRange1(0, 6, 16, 20)
        if case == "general":
            trA, trB, trC = TiledRange([bt, d1, d2]), TiledRange([bt, d2, d1]), TiledRange([bt, d1, d1])
            target, lidx, ridx = "b,i,j", "b,i,k", "b,k,j"
        else:  # non-canonical argument layouts: permuted first
            trA, trB, trC = TiledRange([d1, d2, bt]), TiledRange([d1, d2, bt]), TiledRange([bt, d1, d1])
            target, lidx, ridx = "b,i,j", "i,k,b", "j,k,b"
    else:  # c5 / c5s: both operands explicitly permuted (up front / streamed per SUMMA window)
        s1, b1 = TiledRange1.make_uniform(8, 2), TiledRange1.make_uniform(24, 8)
        trA, trB, trC = TiledRange([s1, s1, b1, b1]), TiledRange([s1, b1, s1, b1]), TiledRange([s1, b1, s1, b1])
        target, lidx, ridx = "i,a,j,b", "i,k,a,c", "j,c,k,b"
        ContEngine.stream_permutes = case == "c5s"
    a, b, (Pr, Pc) = contraction_arrays(world, target, lidx, trA, ridx, trB, shA, shB, lazy_seeds=lazy)
    world.init_comm(Pr, Pc)
    if lazy[0] is None:
        a.fill_random(707)
    if lazy[1] is None:
        b.fill_random(808)
    c = DistArray(world, trC)
    prod = a[lidx] * b[ridx]
    c[target] = prod.set_shape(mask) if mask is not None else prod
    st = ContEngine.last_stats
    A, B = host_full(trA, 707, shA), host_full(trB, 808, shB)
    spec = f"{lidx.replace(',', '')},{ridx.replace(',', '')}->{target.replace(',', '')}"
    ref = np.einsum(spec, A, B, optimize=True)
    worst = 0.0
    for o in sorted(c.tiles):
        sl = trC.tile_slices(trC.tile_index(o))
        worst = max(worst, O.rel_frobenius(c.find(o), ref[sl]))
    # absent tiles must be zero blocks of the reference (or masked away)
    absent_ok = True
    for o in range(trC.ntiles):
        if not c.shape.is_dense() and c.is_zero(o) and mask is None:
            absent_ok = absent_ok and not ref[trC.tile_slices(trC.tile_index(o))].any()
    cnt = torch.tensor([len(c.tiles), st.npairs, 0 if absent_ok else 1], device="cuda")
    dist.all_reduce(cnt)
    w = torch.tensor([worst], device="cuda")
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
    if rank == 0:
        want_tiles = trC.ntiles if c.shape.is_dense() else int((c.shape.norms >= np.float32(SparseShape.threshold())).sum())
        ok = w.item() < 1e-12 and int(cnt[0]) == want_tiles and int(cnt[2]) == 0
        if case in ("c3", "c3m"):
            pa, pb = (shA.norms > 0).astype(int), (shB.norms > 0).astype(int)
            want_pairs = int(((pa @ pb) * ((shA.norms > 0) if mask is not None else 1)).sum())
            ok = ok and int(cnt[1]) == want_pairs
        print(f"MULTI_GPU_RESULT ok={ok} case={case} worst={w.item():.3e} tiles={int(cnt[0])}/{want_tiles} pairs={int(cnt[1])} "
              f"grid={Pr}x{Pc} swapped={st.swapped} lazy_tiles={st.lazy_tiles} ms={st.device_ms:.2f}", flush=True)
    dist.barrier()
    world.dev.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
