"""Pin the CPU oracle against the known answers of the reference's own tests (no GPU).

Each test cites the reference test (file:line under /root/reference) whose closed form it
restates; see tests/known_answers.py.
"""
import numpy as np
import pytest

from oracle import cpu as ocpu
from oracle import ta_oracle as O
from tests import known_answers as KA


# ---- permutation ----------------------------------------------------------------------------
@pytest.mark.parametrize("extent,perm,out_order", KA.LIBRETT_CASES)
def test_permute_librett_known_answers(extent, perm, out_order):
    """tests/librett.cpp:475-745."""
    a = KA.librett_input(extent)
    for impl in (O.tile_permute, O.tile_permute_loops, ocpu.permute):
        b = impl(a, perm)
        assert tuple(b.shape) == tuple(O.permute_array(perm, extent))
        assert KA.librett_check(b, extent, out_order), impl.__name__


def test_permute_array_image_convention():
    """permutation.h:69-79 / tests/permutation.cpp: result[perm[i]] = arg[i]."""
    assert O.permute_array([2, 0, 1], ["a", "b", "c"]) == ["b", "c", "a"]
    p = [3, 1, 0, 2]
    assert O.permute_array(O.perm_inverse(p), O.permute_array(p, list("wxyz"))) == list("wxyz")


# ---- tile GEMM --------------------------------------------------------------------------------
@pytest.mark.parametrize("opA", [0, 1])
@pytest.mark.parametrize("opB", [0, 1])
def test_gemm_vs_triple_loop(opA, opB):
    """tests/math_blas.cpp:156-250: BLAS wrapper vs naive triple loop, tol 1e-3 %."""
    rng = np.random.default_rng(1)
    m, n, k = 7, 5, 9
    A = rng.uniform(-1, 1, (m, k) if opA == 0 else (k, m))
    B = rng.uniform(-1, 1, (k, n) if opB == 0 else (n, k))
    C0 = rng.uniform(-1, 1, (m, n))
    ref = O.tile_gemm_loops(opA, opB, m, n, k, 3.0, A, B, 2.0, C0)
    np.testing.assert_allclose(ocpu.gemm(opA, opB, m, n, k, 3.0, A, B, 2.0, C0), ref, rtol=1e-5)
    np.testing.assert_allclose(ocpu.gemm(opA, opB, m, n, k, 3.0, A, B, 2.0, C0, naive=True), ref, rtol=1e-12)
    helper = O.GemmHelper(opA, opB, 2, 2, 2)
    np.testing.assert_allclose(O.tile_gemm(A, B, 3.0, helper) + 2.0 * C0, ref, rtol=1e-12)


@pytest.mark.parametrize("opA", [0, 1])
@pytest.mark.parametrize("opB", [0, 1])
def test_contract_reduce_matrix_multiply_exact(opA, opB):
    """tests/tile_op_contract_reduce.cpp:109-220: integer tiles, factor 3, NN/TN/NT/TT, then a
    second application accumulates (result == 2 * 3 * A * B)."""
    rng = np.random.default_rng(2)
    m, n, k = KA.CONTRACT_REDUCE_MNK
    A = KA.int_tile(rng, (m, k))
    B = KA.int_tile(rng, (k, n))
    left = A if opA == 0 else np.ascontiguousarray(A.T)
    right = B if opB == 0 else np.ascontiguousarray(B.T)
    helper = O.GemmHelper(opA, opB, 2, 2, 2)
    res = O.tile_gemm(left, right, KA.CONTRACT_REDUCE_FACTOR, helper)
    assert np.array_equal(res, 3 * (A @ B))
    res = O.tile_gemm(left, right, KA.CONTRACT_REDUCE_FACTOR, helper, result=res)
    assert np.array_equal(res, 6 * (A @ B))


def test_gemm_helper_rank_fusion():
    """math/gemm_helper.h:62-98,255-274 and tests/tile_op_contract_reduce.cpp:222-345 (rank-3/4)."""
    h = O.GemmHelper(O.NoTrans, O.NoTrans, 4, 3, 3)  # (a,b,k) x (k,c,d) -> (a,b,c,d)
    assert h.num_contract_ranks == 1
    assert h.compute_matrix_sizes((2, 3, 5), (5, 7, 11)) == (6, 77, 5)
    assert h.make_result_extent((2, 3, 5), (5, 7, 11)) == (2, 3, 7, 11)
    h = O.GemmHelper(O.Trans, O.Trans, 4, 4, 4)  # (c,d,i,j)^T x (a,b,c,d)^T -> (i,j,a,b)
    assert (h.left_inner, h.left_outer, h.right_outer, h.right_inner) == ((0, 2), (2, 4), (0, 2), (2, 4))
    assert h.compute_matrix_sizes((3, 4, 5, 6), (7, 8, 3, 4)) == (30, 56, 12)


# ---- SparseShape --------------------------------------------------------------------------------
def _fixture_trange():
    d = O.TiledRange1(KA.FIXTURE_BOUNDS)
    return O.TiledRange((d,) * KA.FIXTURE_RANK)


def test_sparse_shape_ctor_scaling_and_screening():
    """tests/sparse_shape.cpp:82-154: norm/volume, hard zero below threshold, zero count."""
    tr = _fixture_trange()
    rng = np.random.default_rng(42)
    ext = [d.extents for d in tr.dims]
    norms = KA.fixture_norms(rng, tr.tiles_shape, ext, 0.5, KA.SPARSE_FIXTURE_THRESHOLD)
    sh = O.SparseShape.from_tile_norms(norms, tr, KA.SPARSE_FIXTURE_THRESHOLD)
    nz = 0
    for idx in np.ndindex(*tr.tiles_shape):
        vol = float(np.prod(tr.tile_extent(idx)))
        expected = norms[idx] / vol
        if expected < KA.SPARSE_FIXTURE_THRESHOLD:
            expected = 0.0
            nz += 1
            assert sh.norms[idx] == 0.0 and sh.is_zero(idx)
        else:
            assert not sh.is_zero(idx)
            np.testing.assert_allclose(sh.norms[idx], expected, rtol=1e-6)  # BOOST_CHECK_CLOSE 1e-4 %
    assert sh.zero_tile_count == nz
    assert abs(sh.sparsity() - nz / norms.size) < 1e-6


@pytest.mark.parametrize("seed", [23, 82, 7])
def test_sparse_shape_gemm_reference_formula(seed):
    """tests/sparse_shape.cpp:1478-1541: result == (left*vol).gemm(right*vol, 7.2)/(size_0*size_1),
    thresholded; tolerance 1e-4 % — the reference's own exactness for this function."""
    tr = _fixture_trange()
    rng = np.random.default_rng(seed)
    ext = [d.extents for d in tr.dims]
    thr0 = KA.SPARSE_FIXTURE_THRESHOLD
    left = O.SparseShape.from_tile_norms(KA.fixture_norms(rng, tr.tiles_shape, ext, 0.1, thr0), tr, thr0)
    right = O.SparseShape.from_tile_norms(KA.fixture_norms(rng, tr.tiles_shape, ext, 0.1, thr0), tr, thr0)
    thr = 10 * thr0  # tweak_threshold(): the result inherits the (changed) default threshold
    helper = O.GemmHelper(O.NoTrans, O.NoTrans, 2, 3, 3)
    res = left.gemm(right, -7.2, helper, threshold=thr)
    vol = np.empty(tr.tiles_shape, dtype=np.float64)
    for idx in np.ndindex(*tr.tiles_shape):
        vol[idx] = np.prod(tr.tile_extent(idx))
    m = tr.tiles_shape[0]
    n = tr.tiles_shape[-1]
    L = (left.norms.astype(np.float64) * vol).reshape(m, -1)
    R = (right.norms.astype(np.float64) * vol).reshape(-1, n)
    result_norms = 7.2 * (L @ R)
    nz = 0
    for i in range(m):
        for j in range(n):
            expected = result_norms[i, j] / (tr.dims[0].extents[i] * tr.dims[2].extents[j])
            if expected < thr:
                expected = 0.0
            got = float(res.norms[i, j])
            if got < thr:
                assert res.is_zero((i, j))
                nz += 1
            assert got == pytest.approx(expected, rel=1e-5, abs=1e-12)
    assert abs(res.sparsity() - nz / float(m * n)) < 1e-6


@pytest.mark.parametrize("seed", [5, 61])
def test_sparse_shape_scale_add_mult_reference_formulas(seed):
    """tests/sparse_shape.cpp:763-834 (scale), :836-915 (add, add_scale; subt == add, sparse_shape.h:1499),
    :1320-1395 (mult, mult_scale): expected values are the reference tests' own formulas, tolerance
    1e-4 % (BOOST_CHECK_CLOSE), zero decisions and zero counts exact."""
    tr = _fixture_trange()
    rng = np.random.default_rng(seed)
    ext = [d.extents for d in tr.dims]
    thr = KA.SPARSE_FIXTURE_THRESHOLD
    left = O.SparseShape.from_tile_norms(KA.fixture_norms(rng, tr.tiles_shape, ext, 0.3, thr), tr, thr)
    right = O.SparseShape.from_tile_norms(KA.fixture_norms(rng, tr.tiles_shape, ext, 0.3, thr), tr, thr)
    vol = np.empty(tr.tiles_shape, dtype=np.float64)
    for idx in np.ndindex(*tr.tiles_shape):
        vol[idx] = np.prod(tr.tile_extent(idx))
    L, R = left.norms.astype(np.float64), right.norms.astype(np.float64)
    cases = [(left.scale(-4.1), L * 4.1), (left.add(right), L + R), (left.add(right, -7.2), (L + R) * 7.2),
             (left.mult(right), L * R * vol), (left.mult(right, 2.5), L * R * 2.5 * vol)]
    for res, expected in cases:
        expected = np.where(expected < thr, 0.0, expected)
        nz = 0
        for idx in np.ndindex(*tr.tiles_shape):
            got = float(res.norms[idx])
            assert got == pytest.approx(expected[idx], rel=1e-5, abs=1e-12)
            if got < thr:
                assert res.is_zero(idx)
                nz += 1
            else:
                assert not res.is_zero(idx)
        assert res.zero_tile_count == nz


def test_sparse_shape_gemm_c_and_numpy_oracles_bit_identical():
    """The oracle's fixed fp32 order is stated twice (numpy, C with -ffp-contract=off); the two
    statements must agree bit for bit — this is the spec the CUDA screening kernel is held to."""
    rng = np.random.default_rng(5)
    a = rng.uniform(0, 3, (37, 29)).astype(np.float32)
    b = rng.uniform(0, 3, (29, 41)).astype(np.float32)
    ksz = rng.integers(1, 12, 29).astype(np.float32)
    ref = O.shape_gemm_kernel(a, b, ksz, np.float32(7.2))
    thr = np.float32(np.median(ref))
    got, nz = ocpu.shape_gemm(a, b, ksz, 7.2, float(thr))
    refz = np.where(ref < thr, np.float32(0), ref)
    assert np.array_equal(got.view(np.uint32), refz.view(np.uint32))
    assert nz == int((ref < thr).sum())


def test_sparse_shape_perm_and_mask():
    """sparse_shape.h:1222 (perm), :653-676 (mask); tests/sparse_shape.cpp perm/mask cases."""
    tr = _fixture_trange()
    rng = np.random.default_rng(3)
    ext = [d.extents for d in tr.dims]
    thr = KA.SPARSE_FIXTURE_THRESHOLD
    sh = O.SparseShape.from_tile_norms(KA.fixture_norms(rng, tr.tiles_shape, ext, 0.5, thr), tr, thr)
    perm = [1, 2, 0]  # make_perm(): i -> i+1, last -> 0 (tests/sparse_shape_fixture.h:84-91)
    p = sh.perm(perm)
    for idx in np.ndindex(*tr.tiles_shape):
        assert p.norms[tuple(O.permute_array(perm, idx))] == sh.norms[idx]
    assert p.zero_tile_count == sh.zero_tile_count
    other = O.SparseShape.from_tile_norms(KA.fixture_norms(rng, tr.tiles_shape, ext, 0.5, thr), tr, thr)
    m = sh.mask(other)
    for idx in np.ndindex(*tr.tiles_shape):
        if other.is_zero(idx):
            assert m.is_zero(idx)
        else:
            assert m.norms[idx] == sh.norms[idx]
    assert m.zero_tile_count == int((m.norms < np.float32(thr)).sum())


# ---- ProcGrid / CyclicPmap --------------------------------------------------------------------------
def test_proc_grid_invariants_random():
    """tests/proc_grid.cpp:42-155 random_constructor_test invariants."""
    rng = np.random.default_rng(11)
    for _ in range(100):
        nprocs = int(rng.integers(1, 4096))
        rows = int(rng.integers(1, 1024))
        cols = int(rng.integers(1, 1024))
        row_size = rows * int(rng.integers(1, 512))
        col_size = cols * int(rng.integers(1, 513))
        g0 = O.proc_grid(0, nprocs, rows, cols, row_size, col_size)
        assert g0.proc_size <= nprocs and g0.proc_size == g0.proc_rows * g0.proc_cols
        assert 1 <= g0.proc_rows <= min(rows, nprocs) and 1 <= g0.proc_cols <= min(cols, nprocs)
        assert (g0.rank_row, g0.rank_col) == (0, 0)
        # sum of local sizes over the grid covers the matrix exactly (sampled along one row/col)
        lr = sum(O.proc_grid(rr * g0.proc_cols, nprocs, rows, cols, row_size, col_size).local_rows
                 for rr in range(g0.proc_rows))
        lc = sum(O.proc_grid(rc, nprocs, rows, cols, row_size, col_size).local_cols for rc in range(g0.proc_cols))
        assert (lr, lc) == (rows, cols)
        if g0.proc_size < nprocs:
            gx = O.proc_grid(g0.proc_size, nprocs, rows, cols, row_size, col_size)
            assert (gx.rank_row, gx.rank_col, gx.local_size) == (-1, -1, 0)


def test_proc_grid_named_configs():
    """grids for the BASELINE configs, traced from proc_grid.h:97-178,226-245."""
    for P, want in ((1, (1, 1)), (2, (1, 2)), (4, (2, 2)), (8, (4, 2))):
        g = O.proc_grid(0, P, 32, 32, 32768, 32768)
        assert (g.proc_rows, g.proc_cols) == want
    g = O.proc_grid(0, 8, 4, 169, 100 * 100, 800 * 800)  # C4 as written: 4 x 169 fused tile grid
    assert (g.proc_rows, g.proc_cols) == (1, 8)


def test_cyclic_owner():
    """tests/cyclic_pmap.cpp:82-118 owner: (row % Pr) * Pc + col % Pc."""
    for rows, cols, pr, pc in ((5, 7, 2, 3), (10, 10, 4, 2), (3, 3, 1, 1)):
        for t in range(rows * cols):
            assert O.cyclic_owner(t, cols, pr, pc) == ((t // cols) % pr) * pc + (t % cols) % pc


# ---- GEMM permutation optimizer ----------------------------------------------------------------------
def test_permopt_named_configs():
    """SURVEY.md §8(a7) worked results of permopt.h:254-376 for the BASELINE configs."""
    p = O.plan_contraction("m,n", "m,k", "k,n")
    assert (p.opA, p.opB, p.perm_left, p.perm_right, p.perm_result) == (0, 0, None, None, None)
    p = O.plan_contraction("a,b,i,j", "c,d,i,j", "a,b,c,d")  # C4
    assert (p.left_permtype, p.right_permtype) == (O.PT_TRANSPOSE, O.PT_TRANSPOSE)
    assert p.result_gemm == list("ijab") and p.perm_result == [2, 3, 0, 1]
    p = O.plan_contraction("i,a,j,b", "i,k,a,c", "j,c,k,b")  # C5
    assert (p.left_permtype, p.right_permtype) == (O.PT_GENERAL, O.PT_GENERAL)
    assert p.left_target == list("iack") and p.right_target == list("ckjb") and p.perm_result is None
    assert p.perm_left == [0, 3, 1, 2] and p.perm_right == [2, 0, 1, 3]


# ---- SUMMA vs dense -------------------------------------------------------------------------------------
@pytest.mark.parametrize("grid", [(1, 1), (2, 2), (2, 3)])
def test_summa_contract_exact_int_sparse(grid):
    """tests/dist_eval_contraction_eval.cpp:293-472 (eval / sparse_eval): integer tiles, result
    gathered and compared to l*r exactly; zero tiles => all-zero reference block (:434-445)."""
    rng = np.random.default_rng(9)
    m_ext, k_ext, n_ext = [2, 3, 5, 7], [3, 2, 5], [7, 2, 3, 5, 2]
    Mt, Kt, Nt = len(m_ext), len(k_ext), len(n_ext)
    a_zero = rng.random((Mt, Kt)) < 0.4
    b_zero = rng.random((Kt, Nt)) < 0.4
    a_tiles = {(i, k): KA.int_tile(rng, (m_ext[i], k_ext[k])) for i in range(Mt) for k in range(Kt) if not a_zero[i, k]}
    b_tiles = {(k, j): KA.int_tile(rng, (k_ext[k], n_ext[j])) for k in range(Kt) for j in range(Nt) if not b_zero[k, j]}
    c_zero = ~((~a_zero).astype(int) @ (~b_zero).astype(int) > 0)
    out, npairs = O.summa_contract(a_tiles, b_tiles, Mt, Nt, Kt, 0, 0, 1.0, a_zero, b_zero, c_zero, *grid)
    mo, ko, no = np.concatenate([[0], np.cumsum(m_ext)]), np.concatenate([[0], np.cumsum(k_ext)]), np.concatenate([[0], np.cumsum(n_ext)])
    A = np.zeros((mo[-1], ko[-1]))
    B = np.zeros((ko[-1], no[-1]))
    for (i, k), t in a_tiles.items():
        A[mo[i]:mo[i + 1], ko[k]:ko[k + 1]] = t
    for (k, j), t in b_tiles.items():
        B[ko[k]:ko[k + 1], no[j]:no[j + 1]] = t
    Cref = A @ B
    for i in range(Mt):
        for j in range(Nt):
            blk = Cref[mo[i]:mo[i + 1], no[j]:no[j + 1]]
            if c_zero[i, j]:
                assert (i, j) not in out and not blk.any()
            else:
                assert np.array_equal(out[(i, j)], blk)
    assert npairs == int(((~a_zero).astype(int) @ (~b_zero).astype(int)).sum())
    # the C restatement (threads + vendor DGEMM per pair) agrees
    out_c, _, npairs_c = ocpu.cpu_contract(a_tiles, b_tiles, m_ext, n_ext, k_ext, 0, 0, 1.0, c_zero, nthreads=3)
    assert npairs_c == npairs
    for key, t in out.items():
        assert np.array_equal(out_c[key], t)


def test_screening_zero_pattern_is_order_insensitive_away_from_threshold():
    """The result shape of a contraction is an fp32 norm product (sparse_shape.h:1589-1663). The engine and this
    oracle sum k sequentially with individually rounded operations; the reference calls a vendor SGEMM whose
    summation order is unspecified, so values agree only to rounding (the reference's own test pins 1e-4 %,
    tests/sparse_shape.cpp:1478-1611). What matters downstream is the ZERO PATTERN (which result tiles exist): for
    shapes whose non-zero products are well separated from the threshold — any realistic tile norm — every summation
    order, and exact arithmetic, gives the same pattern. (INTEGRATION.md §5 documents the remaining difference:
    products within a few ulp of the threshold.)"""
    rng = np.random.default_rng(12)
    thr = np.float32(O.FLT_EPSILON)
    for trial in range(20):
        Mt, Nt, Kt = (int(x) for x in rng.integers(3, 40, 3))
        a = np.where(rng.random((Mt, Kt)) < 0.3, rng.uniform(1e-3, 10.0, (Mt, Kt)), 0.0).astype(np.float32)
        b = np.where(rng.random((Kt, Nt)) < 0.3, rng.uniform(1e-3, 10.0, (Kt, Nt)), 0.0).astype(np.float32)
        ksz = rng.integers(1, 600, Kt).astype(np.float32)
        seq = O.shape_gemm_kernel(a, b, ksz, np.float32(1.0))                       # the engine's order
        exact = (a.astype(np.float64) * ksz) @ (b.astype(np.float64) * ksz[:, None])  # any order, no rounding
        rev = O.shape_gemm_kernel(a[:, ::-1].copy(), b[::-1].copy(), ksz[::-1].copy(), np.float32(1.0))  # reversed k
        pat = seq >= thr
        assert np.array_equal(pat, exact >= float(thr)) and np.array_equal(pat, rev >= thr)
        assert np.allclose(seq, exact, rtol=1e-5)
