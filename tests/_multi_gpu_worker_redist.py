"""Worker of tests/test_gpu_multi.py::test_contraction_with_redistribution: run under torchrun.

The operands are created with process maps that have nothing to do with SUMMA's cyclic maps (round robin
by tile ordinal, and "everything on the last rank"); the engine redistributes their tiles over NCCL
send/recv before the SUMMA (what the reference's ArrayEvalImpl does tile by tile, array_eval.h:170).
Cases: a matrix product and the permuted 4-index contraction of BASELINE config 5 in miniature.
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
from tiledarray_b200.tiledarray import ContEngine, DistArray, TiledRange, TiledRange1, World  # noqa: E402


def host_full(tr, seed):
    full = np.zeros(tr.elements_shape)
    for o in range(tr.ntiles):
        idx = tr.tile_index(o)
        ext = tr.tile_extent(idx)
        full[tr.tile_slices(idx)] = util_rng.tile_fill(o, int(np.prod(ext)), seed).reshape(ext)
    return full


def main():
    rank, size, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world = World(device=Device(local), rank=rank, size=size)
    s, b = TiledRange1(0, 16, 48), TiledRange1(0, 32, 64, 80)
    cases = [("m,n", "m,k", "k,n", [b, s], [s, b], "mk,kn->mn"),
             ("i,a,j,b", "i,k,a,c", "j,c,k,b", [s, s, b, b], [s, b, s, b], "ikac,jckb->iajb")]
    worst, ok = 0.0, True
    comm_ready = False
    for target, lidx, ridx, dl, dr, spec in cases:
        trL, trR = TiledRange(dl), TiledRange(dr)
        A, B = host_full(trL, 11), host_full(trR, 12)
        ref = np.einsum(spec, A, B)
        # the fused result grid decides the process grid (ProcGrid on (M tiles, N tiles))
        names = dict(zip(lidx.split(","), dl))
        names.update(dict(zip(ridx.split(","), dr)))
        inner = [x for x in lidx.split(",") if x in ridx.split(",")]
        mo = [x for x in lidx.split(",") if x not in inner]
        no = [x for x in ridx.split(",") if x not in inner]
        Mt, Nt = int(np.prod([names[x].ntiles for x in mo])), int(np.prod([names[x].ntiles for x in no]))
        Me, Ne = int(np.prod([names[x].extent for x in mo])), int(np.prod([names[x].extent for x in no]))
        g = world.proc_grid(Mt, Nt, Me, Ne)
        if not comm_ready:
            world.init_comm(g.proc_rows, g.proc_cols)
            comm_ready = True
        elif (g.proc_rows, g.proc_cols) != world.grid:
            continue  # one communicator set per process in this worker
        a = DistArray(world, trL, owner=lambda o: o % size).fill_random(11)          # round robin
        bb = DistArray(world, trR, owner=lambda o: size - 1).fill_random(12)        # everything on the last rank
        trC = TiledRange([names[x] for x in target.split(",")])
        c = DistArray(world, trC)
        c[target] = a[lidx] * bb[ridx]
        for o in sorted(c.tiles):
            sl = trC.tile_slices(trC.tile_index(o))
            worst = max(worst, O.rel_frobenius(c.find(o), ref[sl]))
        n = torch.tensor([len(c.tiles)], device="cuda")
        dist.all_reduce(n)
        ok = ok and int(n.item()) == trC.ntiles
        for x in (a, bb, c):
            x.release()
    w = torch.tensor([worst], device="cuda")
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"MULTI_GPU_RESULT ok={ok and w.item() < 1e-12} worst={w.item():.3e}", flush=True)
    dist.barrier()
    world.dev.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
