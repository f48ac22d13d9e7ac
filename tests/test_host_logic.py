"""Host-side planning logic of libtadev (no GPU needed) against the oracle, bit for bit."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import ta_oracle as O
from tiledarray_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib):
    """The C-ABI library loads without a GPU and exports every entry include/tadev.h declares."""
    hdr = open(os.path.join(ROOT, "include", "tadev.h")).read()
    declared = set(re.findall(r"\b(tadev_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"tadev_ctx", "tadev_stream"}
    assert len(declared) >= 40
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, missing
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert b"sm_100a" in lib.tadev_version()


def test_no_cpu_fallback_without_device(lib):
    """On a box without a GPU a context cannot be created: the product has no CPU path."""
    n = C.c_int()
    _lib.check(lib.tadev_device_count(C.byref(n)))
    if n.value > 0:
        pytest.skip("a CUDA device is present")
    ctx = C.c_void_p()
    rc = lib.tadev_init(0, 0, C.byref(ctx))
    assert rc == _lib.ENODEVICE
    assert b"no CPU path" in lib.tadev_last_error()


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tiledarray_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def _grid(lib, rank, nprocs, rows, cols, rs, cs):
    g = _lib.ProcGridC()
    _lib.check(lib.tadev_proc_grid_make(rank, nprocs, rows, cols, rs, cs, C.byref(g)))
    return g


def test_proc_grid_matches_oracle(lib):
    """proc_grid.h:97-260 restated twice (C++ product, Python oracle): identical on random cases
    in the style of tests/proc_grid.cpp:42-155, for every rank of small grids."""
    rng = np.random.default_rng(0)
    for trial in range(300):
        nprocs = int(rng.integers(1, 4096)) if trial % 3 else int(rng.integers(1, 17))
        rows, cols = int(rng.integers(1, 1024)), int(rng.integers(1, 1024))
        rs, cs = rows * int(rng.integers(1, 512)), cols * int(rng.integers(1, 513))
        ranks = range(nprocs) if nprocs <= 16 else (0, nprocs // 2, nprocs - 1)
        for rank in ranks:
            g, o = _grid(lib, rank, nprocs, rows, cols, rs, cs), O.proc_grid(rank, nprocs, rows, cols, rs, cs)
            assert (g.proc_rows, g.proc_cols, g.proc_size, g.rank_row, g.rank_col, g.local_rows, g.local_cols,
                    g.local_size) == (o.proc_rows, o.proc_cols, o.proc_size, o.rank_row, o.rank_col, o.local_rows,
                                      o.local_cols, o.local_size)


def test_cyclic_owner_matches_oracle(lib):
    own = C.c_int()
    for rows, cols, pr, pc in ((5, 7, 2, 3), (32, 32, 4, 2), (128, 128, 2, 2), (4, 169, 1, 8)):
        for t in range(rows * cols):
            _lib.check(lib.tadev_cyclic_owner(t, cols, pr, pc, C.byref(own)))
            assert own.value == O.cyclic_owner(t, cols, pr, pc)


CONTRACTIONS = [
    ("m,n", "m,k", "k,n"), ("n,m", "m,k", "k,n"), ("m,n", "k,m", "k,n"), ("m,n", "m,k", "n,k"), ("m,n", "k,m", "n,k"),
    ("a,b,i,j", "c,d,i,j", "a,b,c,d"), ("i,a,j,b", "i,k,a,c", "j,c,k,b"), ("i,j,a,b", "i,j,c,d", "a,b,c,d"),
    ("a,b,c", "a,x,y", "y,x,b,c"), ("a,c,b", "a,x,y", "x,y,b,c"), ("i,j", "i,a,b", "b,a,j"), ("i,j", "a,i,b", "j,b,a"),
    ("a,b", "a", "b"), ("a,b,c,d", "a,b", "c,d"), ("i", "i,k", "k"), ("b,a", "a,k,l", "l,b,k"),
    ("x,y,z", "x,k", "y,k,z"), ("p,q", "k1,p,k2", "k2,q,k1"),
]


@pytest.mark.parametrize("target,left,right", CONTRACTIONS)
def test_plan_contraction_matches_oracle(lib, target, left, right):
    """GEMMPermutationOptimizer (permopt.h:254-376) + result permutation, C++ vs oracle."""
    p = _lib.ContractionPlanC()
    _lib.check(lib.tadev_plan_contraction(target.encode(), left.encode(), right.encode(), C.byref(p)))
    o = O.plan_contraction(target, left, right)

    def perm(arr, rank):
        return None if arr[0] < 0 else [int(x) for x in arr[:rank]]

    assert p.left_target.decode() == ",".join(o.left_target)
    assert p.right_target.decode() == ",".join(o.right_target)
    assert p.result_gemm.decode() == ",".join(o.result_gemm)
    assert (p.left_permtype, p.right_permtype, p.opA, p.opB) == (o.left_permtype, o.right_permtype, o.opA, o.opB)
    assert perm(p.perm_left, p.left_rank) == o.perm_left
    assert perm(p.perm_right, p.right_rank) == o.perm_right
    assert perm(p.perm_result, p.result_rank) == o.perm_result
    assert p.inner_rank == o.helper.num_contract_ranks


def test_plan_contraction_rejects_bad_input(lib):
    p = _lib.ContractionPlanC()
    assert lib.tadev_plan_contraction(b"i,j", b"i,i", b"i,j", C.byref(p)) == _lib.EINVAL
    assert lib.tadev_plan_contraction(b"i,z", b"i,k", b"k,j", C.byref(p)) == _lib.EINVAL
    assert lib.tadev_plan_contraction(b"i", b"i,k", b"k,j", C.byref(p)) == _lib.EINVAL


def _schedule(lib, Pr, Pc, r, c, Mt, Nt, Kt, a, b, cn, thr):
    steps_k = (C.c_int32 * (Kt + 1))()
    begin = (C.c_int32 * (Kt + 2))()
    ns, npairs = C.c_int32(), C.c_int64()
    cap = Mt * Nt * max(Kt, 1)
    pi, pj = (C.c_int32 * max(cap, 1))(), (C.c_int32 * max(cap, 1))()

    def ptr(x):
        return x.ctypes.data_as(C.c_void_p) if x is not None else None

    _lib.check(lib.tadev_summa_schedule(Pr, Pc, r, c, Mt, Nt, Kt, ptr(a), ptr(b), ptr(cn), thr, steps_k, begin,
                                        C.byref(ns), pi, pj, cap, C.byref(npairs)))
    out = []
    for s in range(ns.value):
        out.append((steps_k[s], [(pi[q], pj[q]) for q in range(begin[s], begin[s + 1])]))
    return out


@pytest.mark.parametrize("grid", [(1, 1), (1, 2), (2, 2), (4, 2), (3, 2)])
@pytest.mark.parametrize("density", [1.0, 0.5, 0.1])
def test_summa_schedule_matches_oracle(lib, grid, density):
    """Steps (iterate_sparse, contraction_eval.h:974-1001) and tile-pair lists (contract,
    :1311-1384) of every rank: product (C++) vs oracle, exact, including order."""
    rng = np.random.default_rng(int(density * 100) + grid[0] * 7 + grid[1])
    Mt, Nt, Kt = 11, 9, 13
    thr = np.float32(0.5)
    if density == 1.0:
        a = b = cn = None
        az = bz = cz = None
    else:
        a = np.where(rng.random((Mt, Kt)) < density, 1.0, 0.0).astype(np.float32)
        b = np.where(rng.random((Kt, Nt)) < density, 1.0, 0.0).astype(np.float32)
        cn = ((a @ b) > 0).astype(np.float32)
        az, bz, cz = a < thr, b < thr, cn < thr
    Pr, Pc = grid
    total = 0
    for r in range(Pr):
        for c in range(Pc):
            got = _schedule(lib, Pr, Pc, r, c, Mt, Nt, Kt, a, b, cn, float(thr))
            want = O.summa_rank_schedule(Pr, Pc, r, c, Mt, Nt, Kt, az, bz, cz)
            want = [(k, pairs) for (k, pairs) in want]
            assert got == want
            total += sum(len(p) for _, p in got)
    if density == 1.0:
        assert total == Mt * Nt * Kt
    else:
        assert total == int(((~az).astype(int) @ (~bz).astype(int)).sum())


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors in tiledarray_b200/_lib.py must have the C layout of include/tadev.h
    (compiled here with gcc and compared field by field)."""
    import ctypes as C
    import subprocess
    from tiledarray_b200 import _lib as L
    structs = {"tadev_array_desc": L.ArrayDescC, "tadev_contract_options": L.ContractOptionsC,
               "tadev_contraction_info": L.ContractionInfoC, "tadev_contract_stats": L.ContractStatsC,
               "tadev_summa_plan": L.SummaPlanC, "tadev_summa_stats": L.SummaStatsC, "tadev_permute_source": L.PermuteSourceC,
               "tadev_uniform_source": L.UniformSourceC, "tadev_gemm_group": L.GemmGroup, "tadev_gemm_task": L.GemmTask,
               "tadev_proc_grid": L.ProcGridC, "tadev_contraction_plan": L.ContractionPlanC,
               "tadev_contraction_layout_info": L.LayoutInfoC}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "tadev.h"', "int main(void){"]
    for cname, py in structs.items():
        lines.append(f'printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in py._fields_:
            lines.append(f'printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = {}
    for ln in subprocess.check_output([str(exe)], text=True).splitlines():
        cname, fname, val = ln.split()
        got[(cname, fname)] = int(val)
    for cname, py in structs.items():
        assert got[(cname, "size")] == C.sizeof(py), cname
        for fname, _ in py._fields_:
            assert got[(cname, fname)] == getattr(py, fname).offset, (cname, fname)


def _comm_trace(lib, Pr, Pc, r, c, m_ext, n_ext, k_ext, a_n, b_n, c_n, flags, spl, row_blocks):
    import ctypes as C
    from tiledarray_b200 import _lib as L
    sp = L.SummaPlanC()
    sp.Mt, sp.Nt, sp.Kt = len(m_ext), len(n_ext), len(k_ext)
    keep = [np.asarray(x, dtype=np.int64) for x in (m_ext, n_ext, k_ext)]
    sp.m_ext, sp.n_ext, sp.k_ext = (x.ctypes.data_as(C.POINTER(C.c_int64)) for x in keep)
    fp = C.POINTER(C.c_float)
    if a_n is not None:
        norms = [np.ascontiguousarray(x, dtype=np.float32) for x in (a_n, b_n, c_n)]
        keep += norms
        sp.a_norms, sp.b_norms, sp.c_norms = (x.ctypes.data_as(fp) for x in norms)
    sp.threshold = 1e-6
    sp.flags, sp.steps_per_launch, sp.row_blocks = flags, spl, row_blocks
    n = C.c_int64()
    L.check(lib.tadev_summa_comm_trace(Pr, Pc, r, c, C.byref(sp), None, None, None, None, None, 0, C.byref(n)))
    cap = max(n.value, 1)
    comm, group, k, root = (np.zeros(cap, dtype=np.int32) for _ in range(4))
    nbytes = np.zeros(cap, dtype=np.int64)
    L.check(lib.tadev_summa_comm_trace(Pr, Pc, r, c, C.byref(sp), comm.ctypes.data, group.ctypes.data, k.ctypes.data,
                                       root.ctypes.data, nbytes.ctypes.data, cap, C.byref(n)))
    return [(int(comm[i]), int(group[i]), int(k[i]), int(root[i]), int(nbytes[i])) for i in range(n.value)]


@pytest.mark.parametrize("grid", [(1, 2), (2, 1), (2, 2), (4, 2), (2, 3)])
@pytest.mark.parametrize("flags", [0, 1 | 2 | 4, 1, 2 | 4, 8])
@pytest.mark.parametrize("density", [1.0, 0.4, 0.15])
def test_summa_nccl_groups_agree_across_ranks(lib, grid, flags, density):
    """Deadlock-freedom of the panel broadcasts, on the CPU: every rank of a row (column) communicator
    must issue the same broadcasts, in the same order, partitioned into the same NCCL groups (NCCL may
    reorder the calls of one group), whatever its own sparsity pattern, memory placement of the arrays
    (flags: host-resident / lazy operands, host-resident result => row blocks) and window size. The
    first version of the driver cut windows by the rank's own compute steps and hung on 2 GPUs with
    block-sparse host-resident arrays."""
    Pr, Pc = grid
    rng = np.random.default_rng(Pr * 100 + Pc * 10 + flags + int(density * 100))
    for trial in range(4):
        Mt, Nt, Kt = (int(x) for x in rng.integers(3, 10, 3))
        m_ext, n_ext, k_ext = (rng.integers(8, 70, n) * 2 for n in (Mt, Nt, Kt))
        if density < 1.0:
            a_n = (rng.random((Mt, Kt)) < density).astype(np.float32)
            b_n = (rng.random((Kt, Nt)) < density).astype(np.float32)
            c_n = ((a_n @ b_n) > 0).astype(np.float32)
        else:
            a_n = b_n = c_n = None
        for spl, rb in ((0, 0), (1, 0), (2, 3), (3, 2)):
            traces = {(r, c): _comm_trace(lib, Pr, Pc, r, c, m_ext, n_ext, k_ext, a_n, b_n, c_n, flags, spl, rb)
                      for r in range(Pr) for c in range(Pc)}

            def per_comm(t, which):
                calls = [(k, root, nb, g) for (cm, g, k, root, nb) in t if cm == which]
                groups, first = [], {}
                for k, root, nb, g in calls:  # renumber the groups of this communicator 0, 1, 2, ...
                    groups.append((k, root, nb, first.setdefault(g, len(first))))
                return groups

            for r in range(Pr):  # row communicator: ranks (r, *)
                ref = per_comm(traces[(r, 0)], 0)
                for c in range(1, Pc):
                    assert per_comm(traces[(r, c)], 0) == ref, (grid, flags, density, trial, spl, rb, "row", r, c)
            for c in range(Pc):  # column communicator: ranks (*, c)
                ref = per_comm(traces[(0, c)], 1)
                for r in range(1, Pr):
                    assert per_comm(traces[(r, c)], 1) == ref, (grid, flags, density, trial, spl, rb, "col", r, c)
            # roots follow the cyclic maps and every call moves data
            for t in traces.values():
                for cm, g, k, root, nb in t:
                    assert root == (k % Pc if cm == 0 else k % Pr) and nb > 0


def test_general_product_planner_reference_known_answers(lib):
    """tests/general_product.cpp:62-135 (GeneralPermutationOptimizer): fused / contracted / external classes,
    canonical layouts (fused..., external..., contracted...) with the class order following the target, and the
    two error cases (implicit reduction; pure Hadamard belongs to the element-wise engine)."""
    def plan(target, left, right):
        p, nf = _lib.ContractionPlanC(), C.c_int32(-1)
        rc = lib.tadev_plan_general_product(target.encode(), left.encode(), right.encode(), C.byref(p), C.byref(nf))
        return rc, p, nf.value

    rc, p, nf = plan("c,b,i,k", "b,c,i,j", "c,j,b,k")  # two fused indices; class order follows the target
    assert rc == 0 and nf == 2
    assert (p.left_target, p.right_target, p.result_gemm) == (b"c,b,i,j", b"c,b,j,k", b"c,b,i,k")
    assert p.perm_result[0] == -1 and list(p.perm_left[:4]) == [1, 0, 2, 3] and list(p.perm_right[:4]) == [0, 2, 1, 3]
    assert p.inner_rank == 1 and (p.opA, p.opB) == (0, 0)
    rc, p, nf = plan("i", "i,j", "i,j")  # Hadamard reduction: i fused, j contracted, no externals
    assert rc == 0 and nf == 1 and (p.left_target, p.right_target, p.result_gemm) == (b"i,j", b"i,j", b"i")
    assert p.perm_left[0] == -1 and p.perm_right[0] == -1 and p.perm_result[0] == -1 and p.inner_rank == 1
    rc, p, nf = plan("b,i,k", "i,j,b", "k,j,b")  # non-canonical arguments are permuted
    assert rc == 0 and nf == 1 and (p.left_target, p.right_target, p.result_gemm) == (b"b,i,j", b"b,j,k", b"b,i,k")
    rc, p, nf = plan("p,q,r1", "p,r1", "q,r1")  # non-canonical target: evaluated as (r1,p,q), then permuted
    assert rc == 0 and nf == 1 and p.result_gemm == b"r1,p,q" and list(p.perm_result[:3]) == [2, 0, 1]
    rc, p, nf = plan("b,k", "b", "b,k")  # fused broadcast: the left argument is entirely fused
    assert rc == 0 and nf == 1 and (p.left_target, p.right_target, p.result_gemm) == (b"b", b"b,k", b"b,k") and p.inner_rank == 0
    assert plan("i,k", "b,i,j", "b,j,k")[2] == 0  # pure contraction: not a general product
    assert plan("b,i", "b,i,j", "b,i")[0] == _lib.EINVAL  # implicit reduction over j is rejected
    assert plan("b,i,j", "b,i,j", "b,i,j")[0] == _lib.EINVAL  # pure Hadamard: element-wise engine


def test_operand_exchange_planner(lib):
    """tadev_plan_contraction_opt: the exchanged product C^T = B^T A^T is chosen iff it needs fewer explicit
    tile permutations (BASELINE config 4 becomes a plain NN product; config 5 stays as written)."""
    def plan(target, left, right):
        p, sw = _lib.ContractionPlanC(), C.c_int32(-1)
        _lib.check(lib.tadev_plan_contraction_opt(target.encode(), left.encode(), right.encode(), C.byref(p), C.byref(sw)))
        return p, sw.value

    p, sw = plan("a,b,i,j", "c,d,i,j", "a,b,c,d")
    assert sw == 1 and (p.opA, p.opB) == (0, 0) and p.perm_left[0] == p.perm_right[0] == p.perm_result[0] == -1
    assert p.result_gemm == b"a,b,i,j"
    p, sw = plan("i,a,j,b", "i,k,a,c", "j,c,k,b")
    assert sw == 0 and p.perm_left[0] >= 0 and p.perm_right[0] >= 0 and p.perm_result[0] == -1
    p, sw = plan("m,n", "m,k", "k,n")
    assert sw == 0
    p, sw = plan("n,m", "m,k", "k,n")
    assert sw == 1 and (p.opA, p.opB) == (1, 1) and p.perm_result[0] == -1


def test_make_uniform_reference_known_answers():
    """tests/tiled_range1.cpp:352-393 (TiledRange1::make_uniform) on the Python front-end and on the oracle."""
    from tiledarray_b200.tiledarray import TiledRange1
    cases = [((3, 10, 0), (0, 3)), ((50, 10, 0), (0, 10, 20, 30, 40, 50)), ((55, 10, 0), (0, 10, 19, 28, 37, 46, 55)),
             ((59, 10, 0), (0, 10, 20, 30, 40, 50, 59)), ((3, 10, 3), (3, 6)), ((50, 10, 10), (10, 20, 30, 40, 50, 60)),
             ((55, 10, 10), (10, 20, 29, 38, 47, 56, 65)), ((59, 10, 10), (10, 20, 30, 40, 50, 60, 69)),
             ((50, 30, 0), (0, 25, 50))]
    for (extent, tile, lo), want in cases:
        assert TiledRange1.make_uniform(extent, tile, lo).bounds == want
        assert O.TiledRange1.uniform(extent, tile, lo).bounds == want
    assert TiledRange1.make_uniform(32768, 1024).extents == [1024] * 32  # the benchmark tilings are unaffected


def test_cpp_host_api(lib):
    """include/tiledarray.hpp metadata classes + host-only planners, in C++, without a GPU
    (tests/cpp/test_host_api.cpp restates tests/tiled_range1.cpp, tiled_range.cpp, general_product.cpp:96-135)."""
    import subprocess
    exe = os.path.join(ROOT, "tests", "cpp", "build", "test_host_api")
    if not os.path.exists(exe):
        import __graft_entry__
        __graft_entry__.build()
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "tiledarray_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and "HOST API TESTS PASSED" in out.stdout, out.stdout + out.stderr


def test_contraction_layout_named_configs(lib):
    """tadev_contraction_layout: where SUMMA wants every operand tile (fused GEMM-side position, ProcGrid) for the
    BASELINE expressions: plain matrix product, config 4 (operands exchanged: R[ab,ij] = V[ab,cd] T[cd,ij]) and
    config 5 (both operands explicitly permuted: A -> (i,a | c,k), B -> (c,k | j,b))."""
    from tiledarray_b200.tiledarray import TiledRange, TiledRange1, contraction_layout
    t = TiledRange1.make_uniform(40, 10)
    u = TiledRange1.make_uniform(30, 10)
    tr_mk, tr_kn = TiledRange([t, u]), TiledRange([u, t])
    info, (lr, lc), (rr, rc) = contraction_layout(8, "m,n", "m,k", tr_mk, "k,n", tr_kn)
    assert (info.swapped, info.Mt, info.Nt, info.Kt, info.left_role, info.right_role) == (0, 4, 4, 3, 0, 1)
    assert (info.Pr, info.Pc) == (4, 2)
    for o in range(tr_mk.ntiles):
        assert (lr[o], lc[o]) == divmod(o, 3)
    for o in range(tr_kn.ntiles):
        assert (rr[o], rc[o]) == divmod(o, 4)
    # config 4 miniature: o tiles (2), v tiles (3)
    o1, v1 = TiledRange1(0, 4, 6), TiledRange1(0, 3, 6, 8)
    trT, trV = TiledRange([v1, v1, o1, o1]), TiledRange([v1, v1, v1, v1])
    info, (lr, lc), (rr, rc) = contraction_layout(8, "a,b,i,j", "c,d,i,j", trT, "a,b,c,d", trV)
    assert info.swapped == 1 and (info.opA, info.opB) == (0, 0) and (info.left_role, info.right_role) == (1, 0)
    assert (info.Mt, info.Nt, info.Kt) == (9, 4, 9)
    for o in range(trT.ntiles):  # T(c,d,i,j) is the GEMM's right operand B(k = cd, j = ij)
        c, d, i, j = trT.tile_index(o)
        assert (lr[o], lc[o]) == (c * 3 + d, i * 2 + j)
    for o in range(trV.ntiles):  # V(a,b,c,d) is the GEMM's left operand A(i = ab, k = cd)
        a, b, c, d = trV.tile_index(o)
        assert (rr[o], rc[o]) == (a * 3 + b, c * 3 + d)
    # without the exchange: as written, T is the left operand with opA = T (stored [k][m])
    info2, (lr2, lc2), _ = contraction_layout(8, "a,b,i,j", "c,d,i,j", trT, "a,b,c,d", trV, exchange_operands=False)
    assert info2.swapped == 0 and (info2.opA, info2.opB) == (1, 1) and info2.left_role == 0
    for o in range(trT.ntiles):
        c, d, i, j = trT.tile_index(o)
        assert (lr2[o], lc2[o]) == (i * 2 + j, c * 3 + d)
    # config 5 miniature
    s1, b1 = TiledRange1(0, 2, 4), TiledRange1(0, 3, 6, 9)
    trA, trB = TiledRange([s1, s1, b1, b1]), TiledRange([s1, b1, s1, b1])
    info, (lr, lc), (rr, rc) = contraction_layout(4, "i,a,j,b", "i,k,a,c", trA, "j,c,k,b", trB)
    assert info.swapped == 0 and (info.Mt, info.Nt, info.Kt) == (6, 6, 6) and (info.Pr, info.Pc) == (2, 2)
    lt = plan_target = None
    from tiledarray_b200 import _lib as L
    import ctypes as C
    P = L.ContractionPlanC()
    L.check(L.load().tadev_plan_contraction(b"i,a,j,b", b"i,k,a,c", b"j,c,k,b", C.byref(P)))
    lt, rt = P.left_target.decode().split(","), P.right_target.decode().split(",")
    assert set(lt[:2]) == {"i", "a"} and set(rt[2:]) == {"j", "b"} and lt[2:] == rt[:2]
    ntl = {"i": 2, "k": 2, "a": 3, "c": 3, "j": 2, "b": 3}
    for o in range(trA.ntiles):
        idx = dict(zip("ikac", trA.tile_index(o)))
        row = idx[lt[0]] * ntl[lt[1]] + idx[lt[1]]
        col = idx[lt[2]] * ntl[lt[3]] + idx[lt[3]]
        assert (lr[o], lc[o]) == (row, col)
    for o in range(trB.ntiles):
        idx = dict(zip("jckb", trB.tile_index(o)))
        row = idx[rt[0]] * ntl[rt[1]] + idx[rt[1]]
        col = idx[rt[2]] * ntl[rt[3]] + idx[rt[3]]
        assert (rr[o], rc[o]) == (row, col)


def test_tiledarray_shim_type_checks():
    """include/tadev_tiledarray_shim.hpp (TA::Tile<tadevTensor>, is_device_tile, madness::archive load/store, array
    conversions, SummaTadev : DistEvalImpl) is written against TiledArray's own headers; here it is type-checked with
    every template instantiated against the facsimile declarations of the reference interfaces
    (tests/cpp/ta_facsimile/README.md)."""
    import subprocess
    out = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"),
                          "-I", os.path.join(ROOT, "tests", "cpp", "ta_facsimile"),
                          os.path.join(ROOT, "tests", "cpp", "test_shim_syntax.cpp")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]


def test_summa_windows_config4_row_blocks_and_ramp(lib):
    """Window policy of the driver on BASELINE config 4's structure (lazy left operand => result in row blocks, B
    cached after the first block, 8x1 grid): windows ramp up 1, 2, 2, ... steps, and they are cut by the ROW BLOCK's
    share of an A panel (22.7 GB per step for the whole panel, 2.3 GB per block), not by the whole panel — with the
    whole-panel bound every window degenerated to a single step (1690 instead of 850 launches on one GPU)."""
    from tiledarray_b200 import _lib as L
    v = [64] * 12 + [32]
    o = [64, 36]
    m_ext = [x * y for x in v for y in v]      # fused (a,b): 169 tile rows of V
    k_ext = list(m_ext)                        # fused (c,d)
    n_ext = [x * y for x in o for y in o]      # fused (i,j): 4 tile columns
    Pr, Pc = 8, 1
    t = _comm_trace(lib, Pr, Pc, 3, 0, m_ext, n_ext, k_ext, None, None, None, L.SUMMA_A_LAZY, 0, 0)
    cols = [(g, k) for (cm, g, k, root, nb) in t if cm == 1]   # column-communicator broadcasts (B panels), block 0 only
    groups = sorted({g for g, _ in cols})
    # one broadcast per (window, root row): 169 steps in windows of 1, 2, 2, 2, ... => 85 windows
    assert len(groups) == 85, len(groups)
    assert sum(1 for g, _ in cols if g == groups[0]) == 1          # first window: a single step, one root
    assert sum(1 for g, _ in cols if g == groups[1]) == 2          # then two steps = two roots per window
    assert not [x for x in t if x[0] == 0]                          # Pc == 1: A panels never travel
