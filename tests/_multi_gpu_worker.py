"""Worker of tests/test_gpu_multi.py: run under torchrun, one rank per GPU.

Contracts C[m,n] = A[m,k] * B[k,n] with the SUMMA driver over NCCL and checks every local result
tile against the oracle (inputs are regenerated on the host from the counter RNG).
usage: _multi_gpu_worker.py <Mt> <Kt> <Nt> <tile> <density> <steps_per_launch> [device|host]
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
from tiledarray_b200.tiledarray import ContEngine, DistArray, SparseShape, TiledRange, TiledRange1, World, summa_arrays  # noqa: E402


def main():
    Mt, Kt, Nt, tile = (int(x) for x in sys.argv[1:5])
    density = float(sys.argv[5])
    ContEngine.steps_per_launch = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    memory = sys.argv[7] if len(sys.argv) > 7 else "device"
    rank, size, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world = World(device=Device(local), rank=rank, size=size)
    trA = TiledRange([TiledRange1.make_uniform(Mt * tile, tile), TiledRange1.make_uniform(Kt * tile, tile)])
    trB = TiledRange([TiledRange1.make_uniform(Kt * tile, tile), TiledRange1.make_uniform(Nt * tile, tile)])
    trC = TiledRange([trA.dims[0], trB.dims[1]])
    g = world.proc_grid(Mt, Nt, Mt * tile, Nt * tile)
    world.init_comm(g.proc_rows, g.proc_cols)
    shA = shB = None
    rng = np.random.default_rng(5)
    if density < 1.0:
        nA = np.where(rng.random((Mt, Kt)) < density, float(tile), 0.0).astype(np.float32)
        nB = np.where(rng.random((Kt, Nt)) < density, float(tile), 0.0).astype(np.float32)
        shA, shB = SparseShape(world, nA, trA), SparseShape(world, nB, trB)
    a, b = summa_arrays(world, trA, trB, shA, shB, memory)
    a.fill_random(101)
    b.fill_random(202)
    c = DistArray(world, trC, memory=memory)
    c["m,n"] = a["m,k"] * b["k,n"]
    st = ContEngine.last_stats
    # oracle check of every local result tile
    worst = 0.0
    for o in sorted(c.tiles):
        i, j = trC.tile_index(o)
        ref = np.zeros((tile, tile))
        for k in range(Kt):
            if a.is_zero(i * Kt + k) or b.is_zero(k * Nt + j):
                continue
            At = util_rng.tile_fill(i * Kt + k, tile * tile, 101).reshape(tile, tile)
            Bt = util_rng.tile_fill(k * Nt + j, tile * tile, 202).reshape(tile, tile)
            ref += At @ Bt
        worst = max(worst, O.rel_frobenius(c.find(o), ref))
    ntiles = torch.tensor([len(c.tiles), st.npairs], device="cuda")
    dist.all_reduce(ntiles)
    w = torch.tensor([worst], device="cuda")
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
    if rank == 0:
        if density < 1.0:
            nzc = int((c.shape.norms >= np.float32(SparseShape.threshold())).sum())
            want_pairs = int(((shA.norms > 0).astype(int) @ (shB.norms > 0).astype(int)).sum())
        else:
            nzc, want_pairs = Mt * Nt, Mt * Nt * Kt
        ok = w.item() < 1e-12 and int(ntiles[0]) == nzc and int(ntiles[1]) == want_pairs
        print(f"MULTI_GPU_RESULT ok={ok} worst={w.item():.3e} tiles={int(ntiles[0])}/{nzc} pairs={int(ntiles[1])}/{want_pairs} "
              f"grid={g.proc_rows}x{g.proc_cols} ms={st.device_ms:.2f} bcast_bytes={st.bcast_bytes}", flush=True)
    dist.barrier()
    world.dev.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
