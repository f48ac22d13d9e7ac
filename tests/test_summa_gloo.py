"""World-size-2 test of the multi-rank host logic on CPU (gloo).

Each rank asks libtadev for ITS communication + contraction schedule (tadev_proc_grid_make,
tadev_cyclic_owner, tadev_summa_steps, tadev_summa_schedule — the exact lists the CUDA driver
executes, summa.cu) and then plays the driver's role with gloo broadcasts over row/column
groups and the oracle's tile GEMM: if ownership, roots, panel contents or the group-consistent
broadcast decisions were wrong, a rank would hang, crash on a missing tile, or produce a wrong
block. The gathered result is compared with the dense product exactly (integer tiles), the way
tests/dist_eval_contraction_eval.cpp:293-472 checks the reference's Summa.
"""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, case, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import ta_oracle as O
    from tiledarray_b200 import _lib

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    lib = _lib.load()
    m_ext, k_ext, n_ext, density, seed = case
    Mt, Kt, Nt = len(m_ext), len(k_ext), len(n_ext)
    g = _lib.ProcGridC()
    _lib.check(lib.tadev_proc_grid_make(rank, world_size, Mt, Nt, sum(m_ext), sum(n_ext), C.byref(g)))
    Pr, Pc, r, c = g.proc_rows, g.proc_cols, g.rank_row, g.rank_col
    # every rank builds the same global data from the seed but KEEPS only the tiles it owns
    rng = np.random.default_rng(seed)
    thr = np.float32(0.5)
    a_n = np.where(rng.random((Mt, Kt)) < density, 1.0, 0.0).astype(np.float32)
    b_n = np.where(rng.random((Kt, Nt)) < density, 1.0, 0.0).astype(np.float32)
    c_n = ((a_n @ b_n) > 0).astype(np.float32)
    A = {(i, k): rng.integers(0, 101, (m_ext[i], k_ext[k])).astype(np.float64) for i in range(Mt) for k in range(Kt)}
    B = {(k, j): rng.integers(0, 101, (k_ext[k], n_ext[j])).astype(np.float64) for k in range(Kt) for j in range(Nt)}
    own = C.c_int()

    def owner(tile, cols):
        _lib.check(lib.tadev_cyclic_owner(tile, cols, Pr, Pc, C.byref(own)))
        return own.value

    myA = {key: t for key, t in A.items() if a_n[key] >= thr and owner(key[0] * Kt + key[1], Kt) == rank}
    myB = {key: t for key, t in B.items() if b_n[key] >= thr and owner(key[0] * Nt + key[1], Nt) == rank}

    # process groups: my grid row (A panels) and my grid column (B panels); every rank must
    # create every group (torch.distributed contract)
    row_groups = [dist.new_group([rr * Pc + cc for cc in range(Pc)]) for rr in range(Pr)]
    col_groups = [dist.new_group([rr * Pc + cc for rr in range(Pr)]) for cc in range(Pc)]

    fp = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
    cap = Kt + 1
    step_k, flags = (C.c_int32 * cap)(), (C.c_int32 * cap)()
    a_begin, b_begin = (C.c_int32 * (cap + 1))(), (C.c_int32 * (cap + 1))()
    a_rows, b_cols = (C.c_int32 * (Mt * Kt + 1))(), (C.c_int32 * (Kt * Nt + 1))()
    ns = C.c_int32()
    out = {}
    npairs = 0
    if r >= 0:
        _lib.check(lib.tadev_summa_steps(Pr, Pc, r, c, Mt, Nt, Kt, fp(a_n), fp(b_n), fp(c_n), float(thr), step_k, flags,
                                         a_begin, a_rows, b_begin, b_cols, C.byref(ns)))
        for s in range(ns.value):
            k = step_k[s]
            rows = [a_rows[q] for q in range(a_begin[s], a_begin[s + 1])]
            cols = [b_cols[q] for q in range(b_begin[s], b_begin[s + 1])]
            a_panel, b_panel = {}, {}
            if flags[s] & 2:  # A panel travels along my grid row from column k % Pc
                root = r * Pc + (k % Pc)
                for i in rows:
                    t = torch.from_numpy(myA[(i, k)].copy()) if rank == root else torch.empty(m_ext[i], k_ext[k], dtype=torch.float64)
                    dist.broadcast(t, root, group=row_groups[r])
                    a_panel[i] = t.numpy()
            elif flags[s] & 1:
                a_panel = {i: myA[(i, k)] for i in rows}
            if flags[s] & 4:  # B panel travels along my grid column from row k % Pr
                root = (k % Pr) * Pc + c
                for j in cols:
                    t = torch.from_numpy(myB[(k, j)].copy()) if rank == root else torch.empty(k_ext[k], n_ext[j], dtype=torch.float64)
                    dist.broadcast(t, root, group=col_groups[c])
                    b_panel[j] = t.numpy()
            elif flags[s] & 1:
                b_panel = {j: myB[(k, j)] for j in cols}
            if flags[s] & 1:
                helper = O.GemmHelper(O.NoTrans, O.NoTrans, 2, 2, 2)
                for i in rows:
                    for j in cols:
                        if c_n[i, j] < thr:
                            continue
                        assert owner(i * Nt + j, Nt) == rank  # result stays with the grid owner
                        out[(i, j)] = O.tile_gemm(a_panel[i], b_panel[j], 1.0, helper, result=out.get((i, j)))
                        npairs += 1
    gathered = [None] * world_size
    dist.all_gather_object(gathered, (out, npairs))
    if rank == 0:
        mo, ko, no = (np.concatenate([[0], np.cumsum(x)]) for x in (m_ext, k_ext, n_ext))
        Ad, Bd = np.zeros((mo[-1], ko[-1])), np.zeros((ko[-1], no[-1]))
        for (i, k), t in A.items():
            if a_n[i, k] >= thr:
                Ad[mo[i]:mo[i + 1], ko[k]:ko[k + 1]] = t
        for (k, j), t in B.items():
            if b_n[k, j] >= thr:
                Bd[ko[k]:ko[k + 1], no[j]:no[j + 1]] = t
        Cref = Ad @ Bd
        Cgot = np.zeros_like(Cref)
        seen = set()
        for o, _ in gathered:
            for (i, j), t in o.items():
                assert (i, j) not in seen  # each result tile is accumulated entirely by one rank
                seen.add((i, j))
                Cgot[mo[i]:mo[i + 1], no[j]:no[j + 1]] = t
        ok = bool(np.array_equal(Cgot, Cref))
        want_pairs = int(((a_n >= thr).astype(int) @ (b_n >= thr).astype(int))[c_n >= thr].sum())
        ret["ok"] = ok and sum(n for _, n in gathered) == want_pairs
        ret["grid"] = (Pr, Pc)
    dist.barrier()
    dist.destroy_process_group()


CASES = [
    # (m_ext, k_ext, n_ext, density, seed)  -> grid chosen by ProcGrid for 2 ranks
    ([3, 4, 2, 5], [2, 3, 4, 2, 3], [4, 2, 3, 5], 1.0, 1),      # square-ish dense   -> 1 x 2
    ([3, 4, 2, 5, 3, 2], [2, 3, 4, 2, 3], [4, 2, 3], 0.5, 2),   # sparse
    ([2] * 12, [3, 2, 4], [5], 0.7, 3),                          # tall: one tile column   -> 2 x 1
    ([3, 4, 2, 5], [2, 3, 4, 2, 3, 2, 2], [4, 2, 3, 5, 2], 0.25, 4),  # very sparse: skipped steps
]


@pytest.mark.parametrize("case", CASES)
def test_summa_two_ranks_gloo(case):
    import torch.multiprocessing as mp
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, case, ret), nprocs=2, join=True)
    assert ret.get("ok") is True, dict(ret)
    from oracle import ta_oracle as O
    g = O.proc_grid(0, 2, len(case[0]), len(case[2]), sum(case[0]), sum(case[2]))
    assert ret["grid"] == (g.proc_rows, g.proc_cols)
    if len(case[2]) == 1:
        assert ret["grid"] == (2, 1)  # both grid orientations are exercised by CASES


def _plan_worker(rank, world_size, port, case, ret):
    """Play the EXACT broadcast plan tadev_summa_f64 would issue (tadev_summa_comm_trace: one broadcast per window,
    communicator and root, in issue order) with gloo broadcasts on a 2x2 grid. A rank-dependent plan (different call
    order, different byte counts, a missing call) hangs or fails the size check; receivers verify the root's signature."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from tiledarray_b200 import _lib

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    lib = _lib.load()
    m_ext, k_ext, n_ext, density, seed, flags, row_blocks = case
    Mt, Kt, Nt = len(m_ext), len(k_ext), len(n_ext)
    Pr, Pc = 2, 2
    r, c = rank // Pc, rank % Pc
    rng = np.random.default_rng(seed)
    if density < 1.0:
        a_n = (rng.random((Mt, Kt)) < density).astype(np.float32)
        b_n = (rng.random((Kt, Nt)) < density).astype(np.float32)
        c_n = ((a_n @ b_n) > 0).astype(np.float32)
    else:
        a_n = b_n = c_n = None
    sp = _lib.SummaPlanC()
    sp.Mt, sp.Nt, sp.Kt = Mt, Nt, Kt
    keep = [np.asarray(x, dtype=np.int64) for x in (m_ext, n_ext, k_ext)]
    sp.m_ext, sp.n_ext, sp.k_ext = (x.ctypes.data_as(C.POINTER(C.c_int64)) for x in keep)
    if a_n is not None:
        sp.a_norms, sp.b_norms, sp.c_norms = (x.ctypes.data_as(C.POINTER(C.c_float)) for x in (a_n, b_n, c_n))
    sp.threshold = 0.5
    sp.flags, sp.row_blocks = flags, row_blocks
    n = C.c_int64()
    _lib.check(lib.tadev_summa_comm_trace(Pr, Pc, r, c, C.byref(sp), None, None, None, None, None, 0, C.byref(n)))
    cap = max(n.value, 1)
    comm, group, kk, root, nbytes = ((C.c_int32 * cap)(), (C.c_int32 * cap)(), (C.c_int32 * cap)(), (C.c_int32 * cap)(), (C.c_int64 * cap)())
    _lib.check(lib.tadev_summa_comm_trace(Pr, Pc, r, c, C.byref(sp), comm, group, kk, root, nbytes, cap, C.byref(n)))
    row_groups = [dist.new_group([rr * Pc + cc for cc in range(Pc)]) for rr in range(Pr)]
    col_groups = [dist.new_group([rr * Pc + cc for rr in range(Pr)]) for cc in range(Pc)]
    ncalls, moved = 0, 0
    for x in range(n.value):
        which, k, rt, nb = comm[x], kk[x], root[x], nbytes[x]
        grp = row_groups[r] if which == 0 else col_groups[c]
        src = r * Pc + rt if which == 0 else rt * Pc + c   # world rank of the root inside my row / column
        words = max(1, min(int(nb) // 8, 4096))             # a scaled-down payload of the same call
        sig = float(which * 1000003 + k * 1009 + rt * 31) + float(nb % 9973)
        t = torch.full((words,), sig, dtype=torch.float64) if rank == src else torch.zeros(words, dtype=torch.float64)
        dist.broadcast(t, src, group=grp)
        assert bool((t == sig).all()), (rank, x, which, k, rt, nb)
        ncalls += 1
        moved += int(nb)
    gathered = [None] * world_size
    dist.all_gather_object(gathered, (ncalls, moved))
    if rank == 0:
        ret["ok"] = True
        ret["calls"] = [g[0] for g in gathered]
    dist.barrier()
    dist.destroy_process_group()


PLAN_CASES = [
    # (m_ext, k_ext, n_ext, density, seed, plan flags, row_blocks)
    ([64] * 6, [64] * 9, [64] * 5, 1.0, 1, 0, 0),                     # dense, device-resident
    ([64] * 7, [32] * 40, [64] * 6, 0.3, 2, 0, 0),                    # block-sparse, many steps: ramped windows of merged panels
    ([64] * 8, [64] * 10, [64] * 6, 1.0, 3, 1 | 2 | 4, 0),            # host-resident operands and result: row blocks, B cache, depth 3
    ([64] * 8, [64] * 10, [64] * 6, 0.5, 4, 8, 3),                    # lazy left operand, three row blocks, block-sparse
]


@pytest.mark.parametrize("case", PLAN_CASES)
def test_summa_broadcast_plan_four_ranks_gloo(case):
    import torch.multiprocessing as mp
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_plan_worker, args=(4, port, case, ret), nprocs=4, join=True)
    assert ret.get("ok") is True, dict(ret)
    assert all(n > 0 for n in ret["calls"]), ret["calls"]
