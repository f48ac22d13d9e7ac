"""Parity of the CUDA kernels (through the C ABI) against the CPU oracle.

Bars (BASELINE.json north_star): FP64 values within 1e-12 relative Frobenius error; screening,
tile lists and permutations bit-exact. Integer-valued tiles must be reproduced EXACTLY (the
reference's own integration tests use integer tiles for that reason).
"""
import numpy as np
import pytest

from oracle import cpu as ocpu
from oracle import ta_oracle as O
from tests import known_answers as KA
from tiledarray_b200 import OP_N, OP_T, TadevError

pytestmark = pytest.mark.gpu
TOL = 1e-12  # relative Frobenius norm (north_star)


def _gemm(dev, opA, opB, m, n, k, alpha, A, B, beta, C0):
    dA, dB, dC = dev.upload(A), dev.upload(B), dev.upload(C0)
    dev.gemm(opA, opB, m, n, k, alpha, dA, dB, beta, dC)
    out = dev.download(dC, np.float64, (m, n))
    for b in (dA, dB, dC):
        b.free()
    return out


# ---- tile GEMM --------------------------------------------------------------------------------
SHAPES = [(256, 256, 256), (128, 128, 16), (1, 1, 1), (3, 5, 7), (37, 53, 29), (130, 257, 100), (64, 36, 4096),
          (36, 36, 1296), (100, 100, 1000), (512, 384, 200), (2, 2, 2), (129, 127, 17), (16, 2048, 24), (1024, 8, 8),
          (18, 36, 27), (130, 258, 18), (126, 130, 22)]


@pytest.mark.parametrize("opA", [OP_N, OP_T])
@pytest.mark.parametrize("opB", [OP_N, OP_T])
@pytest.mark.parametrize("beta", [0.0, 1.0])
def test_tile_gemm_parity(dev, opA, opB, beta):
    """tadev_gemm_f64 vs the oracle's vendor-DGEMM path over ragged, even/odd (TMA fast path and
    cp.async generic path), tiny and K-tail shapes; tolerance 1e-12 rel. Frobenius."""
    rng = np.random.default_rng(1234 + 2 * opA + opB)
    for (m, n, k) in SHAPES:
        A = rng.uniform(-1, 1, (m, k) if opA == OP_N else (k, m))
        B = rng.uniform(-1, 1, (k, n) if opB == OP_N else (n, k))
        C0 = rng.uniform(-1, 1, (m, n))
        got = _gemm(dev, opA, opB, m, n, k, 1.5, A, B, beta, C0)
        ref = ocpu.gemm(opA, opB, m, n, k, 1.5, A, B, beta, C0)
        assert O.rel_frobenius(got, ref) < TOL, (m, n, k, opA, opB, beta)


@pytest.mark.parametrize("opA", [OP_N, OP_T])
@pytest.mark.parametrize("opB", [OP_N, OP_T])
def test_contract_reduce_exact_int(dev, opA, opB):
    """tests/tile_op_contract_reduce.cpp:109-220 on the GPU: integer tiles, factor 3, second
    application accumulates — results must be EXACT."""
    rng = np.random.default_rng(2)
    m, n, k = KA.CONTRACT_REDUCE_MNK
    A, B = KA.int_tile(rng, (m, k)), KA.int_tile(rng, (k, n))
    left = A if opA == OP_N else np.ascontiguousarray(A.T)
    right = B if opB == OP_N else np.ascontiguousarray(B.T)
    dA, dB, dC = dev.upload(left), dev.upload(right), dev.alloc(m * n * 8)
    dev.gemm(opA, opB, m, n, k, KA.CONTRACT_REDUCE_FACTOR, dA, dB, 0.0, dC)
    assert np.array_equal(dev.download(dC, np.float64, (m, n)), 3 * (A @ B))
    dev.gemm(opA, opB, m, n, k, KA.CONTRACT_REDUCE_FACTOR, dA, dB, 1.0, dC)
    assert np.array_equal(dev.download(dC, np.float64, (m, n)), 6 * (A @ B))


@pytest.mark.parametrize("even", [True, False])
def test_grouped_gemm_chained_accumulation(dev, even):
    """One launch, several result tiles, each the in-register sum of several (left,right) pairs
    with different K (ContractReduce pair-accumulate + add_to merge, contract_reduce.h:386-453),
    mixed beta flags, an empty group and a k == 0 pair."""
    rng = np.random.default_rng(7)
    dims = [(256, 128), (130, 66), (64, 258), (2, 2)] if even else [(37, 53), (129, 5), (3, 200), (1, 1)]
    ks = [64, 128, 16, 0, 36, 250] if even else [7, 64, 33, 0, 1, 250]
    groups, refs, bufs = [], [], []
    for gi, (m, n) in enumerate(dims):
        kk = ks[gi:] + ks[:gi]
        As = [rng.uniform(-1, 1, (m, k)) for k in kk]
        Bs = [rng.uniform(-1, 1, (k, n)) for k in kk]
        C0 = rng.uniform(-1, 1, (m, n))
        acc = gi % 2
        dAs, dBs, dC = [dev.upload(a) for a in As], [dev.upload(b) for b in Bs], dev.upload(C0)
        bufs += dAs + dBs + [dC]
        groups.append((dC.ptr, m, n, acc, [(a.ptr, b.ptr, k) for a, b, k in zip(dAs, dBs, kk)]))
        refs.append((dC, m, n, 0.5 * sum(a @ b for a, b in zip(As, Bs)) + (C0 if acc else 0.0)))
    # an empty group: C = 0 (beta 0) — the k_ == 0 branch of Summa::initialize (contraction_eval.h:1023)
    dZ = dev.upload(rng.uniform(-1, 1, (16, 16)))
    groups.append((dZ.ptr, 16, 16, 0, []))
    refs.append((dZ, 16, 16, np.zeros((16, 16))))
    dev.gemm_grouped(OP_N, OP_N, 0.5, groups)
    for (dC, m, n, ref) in refs:
        got = dev.download(dC, np.float64, (m, n))
        assert O.rel_frobenius(got, ref) < TOL
    for b in bufs + [dZ]:
        b.free()


@pytest.mark.parametrize("tile", [1024, 512])
@pytest.mark.parametrize("opA", [OP_N, OP_T])
@pytest.mark.parametrize("opB", [OP_N, OP_T])
def test_ws_kernel_random_baseline_tiles(dev, tile, opA, opB):
    """The warp-specialised TMA kernel on RANDOM data at BASELINE's tile sizes (1024^3: configs 2 and 5; 512^3:
    config 3), all four op combinations, a chain of three tile pairs per result tile (ContractReduce over K) with
    beta = 0 and beta = 1 groups in one launch, against the oracle's vendor DGEMM, tile by tile, <= 1e-12 rel.
    Frobenius. (Linearity / fill(1) properties cannot see a consistent index error; random data does.)"""
    rng = np.random.default_rng(tile + 2 * opA + opB)
    m = n = k = tile
    nchain = 3
    groups, refs, bufs = [], [], []
    for gi in range(2):
        As = [rng.uniform(-1, 1, (m, k) if opA == OP_N else (k, m)) for _ in range(nchain)]
        Bs = [rng.uniform(-1, 1, (k, n) if opB == OP_N else (n, k)) for _ in range(nchain)]
        C0 = rng.uniform(-1, 1, (m, n))
        dAs, dBs, dC = [dev.upload(a) for a in As], [dev.upload(b) for b in Bs], dev.upload(C0)
        bufs += dAs + dBs + [dC]
        groups.append((dC.ptr, m, n, gi, [(a.ptr, b.ptr, k) for a, b in zip(dAs, dBs)]))
        ref = C0.copy() if gi else np.zeros((m, n))
        for a, b in zip(As, Bs):
            ref = ocpu.gemm(opA, opB, m, n, k, 0.75, a, b, 1.0, ref)
        refs.append((dC, ref))
    dev.gemm_grouped(opA, opB, 0.75, groups)
    for dC, ref in refs:
        got = dev.download(dC, np.float64, (m, n))
        assert O.rel_frobenius(got, ref) < TOL
        # block by block: a misplaced 128 x 128 block must not hide in the global norm
        for bi in range(0, m, 128):
            for bj in range(0, n, 128):
                assert O.rel_frobenius(got[bi:bi + 128, bj:bj + 128], ref[bi:bi + 128, bj:bj + 128]) < 1e-11
    for b in bufs:
        b.free()


def test_gemm_argument_errors(dev):
    """Error behaviour of the boundary: bad arguments return TADEV_EINVAL (TA_ASSERT analogue)."""
    d = dev.alloc(64)
    with pytest.raises(TadevError):
        dev.gemm(OP_N, OP_N, 2, 2, 2, 1.0, d, d, 0.5, d)  # beta must be 0 or 1
    with pytest.raises(TadevError):
        dev.gemm(5, OP_N, 2, 2, 2, 1.0, d, d, 0.0, d)
    with pytest.raises(TadevError):
        dev.gemm(OP_N, OP_N, -1, 2, 2, 1.0, d, d, 0.0, d)
    d.free()


def test_gemm_linearity_large(dev):
    """Size-independent property at a BASELINE-size tile pair (1024^3): (2A)(B) == 2 (A B) exactly
    (power-of-two scaling commutes with rounding), and fill(1) inputs give exactly k everywhere
    (the verification of examples/device/ta_dense_device.cpp)."""
    n = 1024
    dA, dB, dC, dC2 = dev.alloc(n * n * 8), dev.alloc(n * n * 8), dev.alloc(n * n * 8), dev.alloc(n * n * 8)
    dev.fill_uniform(dA, n * n, 11)
    dev.fill_uniform(dB, n * n, 12)
    dev.gemm(OP_N, OP_N, n, n, n, 1.0, dA, dB, 0.0, dC)
    dev.scale(n * n, dA, 2.0)
    dev.gemm(OP_N, OP_N, n, n, n, 1.0, dA, dB, 0.0, dC2)
    c1, c2 = dev.download(dC, np.float64, (n, n)), dev.download(dC2, np.float64, (n, n))
    assert np.array_equal(2.0 * c1, c2)
    ones = np.ones((n, n))
    dev.upload_into(dA, ones)
    dev.upload_into(dB, ones)
    dev.gemm(OP_T, OP_N, n, n, n, 1.0, dA, dB, 0.0, dC)
    assert np.array_equal(dev.download(dC, np.float64, (n, n)), np.full((n, n), float(n)))
    for b in (dA, dB, dC, dC2):
        b.free()


# ---- permutation ------------------------------------------------------------------------------------
@pytest.mark.parametrize("extent,perm,out_order", KA.LIBRETT_CASES)
def test_permute_librett_known_answers(dev, extent, perm, out_order):
    """tests/librett.cpp:475-745 on the new kernel (bit-exact)."""
    a = KA.librett_input(extent)
    dx, dy = dev.upload(a), dev.alloc(a.nbytes)
    dev.permute(extent, perm, 8, dx, dy)
    b = dev.download(dy, np.float64, O.permute_array(perm, extent))
    assert KA.librett_check(b, extent, out_order)
    dx.free(), dy.free()


def test_permute_random_vs_oracle(dev):
    """Random ranks 1..6, unit extents, odd extents, all three kernels (memcpy, row copy, tiled
    transpose), 4/8/16-byte elements; bit-exact vs the oracle."""
    rng = np.random.default_rng(99)
    cases = [((7,), (0,)), ((4, 5, 6), (0, 1, 2)), ((1, 9, 1, 5), (3, 2, 1, 0)), ((33, 65), (1, 0)),
             ((16, 16, 64, 64), (0, 3, 1, 2)), ((16, 64, 16, 64), (2, 0, 1, 3)), ((3, 4, 5, 6, 7, 2), (5, 3, 1, 0, 2, 4)),
             ((64, 36, 64, 36), (2, 3, 0, 1)), ((5, 1, 7), (2, 1, 0)), ((2, 3, 4, 5), (1, 0, 2, 3)), ((130, 3, 70), (2, 1, 0))]
    for _ in range(25):
        rank = int(rng.integers(2, 7))
        cases.append((tuple(int(x) for x in rng.integers(1, 9, rank)), tuple(int(x) for x in rng.permutation(rank))))
    for ext, perm in cases:
        x = rng.uniform(-1, 1, ext)
        dx, dy = dev.upload(x), dev.alloc(x.nbytes)
        dev.permute(ext, perm, 8, dx, dy)
        ref = O.tile_permute(x, perm)
        assert np.array_equal(dev.download(dy, np.float64, ref.shape), ref), (ext, perm)
        dx.free(), dy.free()
    for dtype, nbytes in ((np.float32, 4), (np.complex128, 16)):
        x = (rng.uniform(-1, 1, (6, 10, 14)) + 0).astype(dtype)
        dx, dy = dev.upload(x), dev.alloc(x.nbytes)
        dev.permute(x.shape, (2, 0, 1), nbytes, dx, dy)
        ref = O.tile_permute(x, (2, 0, 1))
        assert np.array_equal(dev.download(dy, dtype, ref.shape), ref)
        dx.free(), dy.free()


@pytest.mark.parametrize("rows,cols", [(1000, 1000), (999, 1001), (1000, 16), (1000, 8), (1000, 32), (16, 1000), (8, 1000),
                                       (6, 1001), (1001, 6), (257, 130)])
def test_permute_fast_transpose_shapes(dev, rows, cols):
    """Every tile shape of the fast tiled-transpose kernel (8x256 ... 256x8), full and ragged tiles,
    16-byte vector path (aligned, even extents) and scalar path (pointer offset by 8 bytes / odd extents),
    single and batched launches; bit-exact vs numpy."""
    rng = np.random.default_rng(rows * 7 + cols)
    nb = 3
    x = rng.uniform(-1, 1, (nb, 2, rows, cols))
    ref = np.ascontiguousarray(np.transpose(x, (0, 1, 3, 2)))
    for misalign in (0, 8):
        dx = dev.alloc(x.nbytes + 16)
        dy = dev.alloc(x.nbytes + 16)
        src, dst = dx.view(misalign, x.nbytes), dy.view(misalign, x.nbytes)
        dev.upload_into(src, x)
        dev.memset(dy, 0)
        tile_bytes = 2 * rows * cols * 8
        if misalign == 0 or tile_bytes % 16 == 0:
            srcs = [src.view(i * tile_bytes, tile_bytes) for i in range(nb)]
            dsts = [dst.view(i * tile_bytes, tile_bytes) for i in range(nb)]
            dev.permute_batched((2, rows, cols), (0, 2, 1), 8, srcs, dsts)
            assert np.array_equal(dev.download(dst, np.float64, ref.shape), ref), ("batched", misalign)
            dev.memset(dy, 0)
        for i in range(nb):
            dev.permute((2, rows, cols), (0, 2, 1), 8, src.view(i * tile_bytes, tile_bytes), dst.view(i * tile_bytes, tile_bytes))
        assert np.array_equal(dev.download(dst, np.float64, ref.shape), ref), ("single", misalign)
        dx.free(), dy.free()


def test_permute_round_trip_large(dev):
    """Size-independent property on a C5-size tile (16,16,64,64 -> 8 MiB): permute then inverse
    permute is the identity, bit for bit."""
    ext, perm = (16, 16, 64, 64), (0, 3, 1, 2)
    n = int(np.prod(ext))
    dx, dy, dz = dev.alloc(n * 8), dev.alloc(n * 8), dev.alloc(n * 8)
    dev.fill_uniform(dx, n, 5)
    dev.permute(ext, perm, 8, dx, dy)
    dev.permute(O.permute_array(perm, ext), O.perm_inverse(perm), 8, dy, dz)
    assert np.array_equal(dev.download(dx, np.float64, ext), dev.download(dz, np.float64, ext))
    for b in (dx, dy, dz):
        b.free()


def test_permute_errors(dev):
    d = dev.alloc(64)
    e = dev.alloc(64)
    with pytest.raises(TadevError):
        dev.permute((2, 2), (0, 0), 8, d, e)  # not a permutation
    with pytest.raises(TadevError):
        dev.permute((2, 2), (1, 0), 8, d, d)  # in place
    with pytest.raises(TadevError):
        dev.permute((2, 2), (1, 0), 3, d, e)  # element size
    d.free(), e.free()


def test_add_to_and_scale(dev):
    """ContractReduce's partial-result merge (contract_reduce.h:397-398) and scale."""
    rng = np.random.default_rng(4)
    x, y = rng.uniform(-1, 1, 10007), rng.uniform(-1, 1, 10007)
    dx, dy = dev.upload(x), dev.upload(y)
    dev.add_to(x.size, dx, dy)
    assert np.array_equal(dev.download(dx, np.float64, x.shape), x + y)
    dev.scale(x.size, dx, -0.75)
    assert np.array_equal(dev.download(dx, np.float64, x.shape), (x + y) * -0.75)
    dx.free(), dy.free()


# ---- SparseShape screening + tile lists -------------------------------------------------------------------
def _fixture_shapes(seed, fill):
    d = O.TiledRange1(KA.FIXTURE_BOUNDS)
    tr = O.TiledRange((d,) * KA.FIXTURE_RANK)
    rng = np.random.default_rng(seed)
    norms = KA.fixture_norms(rng, tr.tiles_shape, [x.extents for x in tr.dims], fill, KA.SPARSE_FIXTURE_THRESHOLD)
    return tr, norms


@pytest.mark.parametrize("rank", [1, 2, 3, 4])
def test_shape_scale_bit_exact(dev, rank):
    """scale_tile_norms<InverseVolume> (sparse_shape.h:149-217) incl. the rank-1 divide branch."""
    rng = np.random.default_rng(rank)
    ext = [rng.integers(1, 40, int(rng.integers(2, 7))).astype(np.float32) for _ in range(rank)]
    norms = rng.uniform(0, 50, [len(e) for e in ext]).astype(np.float32)
    thr = 0.05
    ref, nz_ref = O.shape_scale_norms(norms, ext, thr)
    left, right = O.shape_scale_factors(ext)
    got, nz = dev.shape_scale(norms, left, right, thr)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)) and nz == nz_ref


@pytest.mark.parametrize("seed", [23, 82, 7])
def test_shape_gemm_bit_exact(dev, seed):
    """SparseShape::gemm (sparse_shape.h:1589-1681) on the reference's own fixture shapes
    (tests/sparse_shape.cpp:1478-1541): norms and zero count bit-exact vs the oracle."""
    tr, nl = _fixture_shapes(seed, 0.1)
    _, nr = _fixture_shapes(seed + 1, 0.1)
    thr0 = KA.SPARSE_FIXTURE_THRESHOLD
    left = O.SparseShape.from_tile_norms(nl, tr, thr0)
    right = O.SparseShape.from_tile_norms(nr, tr, thr0)
    helper = O.GemmHelper(O.NoTrans, O.NoTrans, 2, 3, 3)
    ref = left.gemm(right, -7.2, helper, threshold=10 * thr0)
    ksz = O._recursive_outer_product(left.size_vectors[1:3], False)
    got, nz = dev.shape_gemm(left.norms.reshape(5, 25), right.norms.reshape(25, 5), ksz, 7.2, 10 * thr0)
    assert np.array_equal(got.view(np.uint32), ref.norms.view(np.uint32))
    assert nz == ref.zero_tile_count


def test_shape_gemm_bit_exact_c3_size_and_outer(dev):
    """Config-3-sized shapes (128 x 128 tile grids, 10 % density) and the outer-product branch."""
    rng = np.random.default_rng(6)
    a = np.where(rng.random((128, 128)) < 0.1, rng.uniform(0.5, 2, (128, 128)), 0).astype(np.float32)
    b = np.where(rng.random((128, 128)) < 0.1, rng.uniform(0.5, 2, (128, 128)), 0).astype(np.float32)
    ksz = np.full(128, 512, dtype=np.float32)
    ref = O.shape_gemm_kernel(a, b, ksz, np.float32(1.0))
    thr = np.float32(O.FLT_EPSILON)
    refz = np.where(ref < thr, np.float32(0), ref)
    got, nz = dev.shape_gemm(a, b, ksz, 1.0, float(thr))
    assert np.array_equal(got.view(np.uint32), refz.view(np.uint32)) and nz == int((ref < thr).sum())
    u, v = rng.uniform(0, 2, 17).astype(np.float32), rng.uniform(0, 2, 9).astype(np.float32)
    ref = (np.multiply.outer(u, v).astype(np.float32) * np.float32(3.0)).astype(np.float32)
    refz = np.where(ref < np.float32(1.0), np.float32(0), ref)
    got, nz = dev.shape_gemm(u, v, np.zeros(0, np.float32), 3.0, 1.0)
    assert np.array_equal(got.view(np.uint32), refz.view(np.uint32)) and nz == int((ref < 1.0).sum())


@pytest.mark.parametrize("grid", [(1, 1), (2, 2), (4, 2)])
def test_pairlist_bit_exact(dev, grid):
    """tadev_build_pairlist vs Summa::contract restated (contraction_eval.h:1311-1384): same
    pairs in the same order for every rank and step, sparse and dense."""
    rng = np.random.default_rng(3)
    Mt, Nt, Kt = 37, 29, 11
    thr = np.float32(0.5)
    a = np.where(rng.random((Mt, Kt)) < 0.3, 1.0, 0.0).astype(np.float32)
    b = np.where(rng.random((Kt, Nt)) < 0.3, 1.0, 0.0).astype(np.float32)
    cn = ((a @ b) > 0).astype(np.float32)
    cn[rng.random((Mt, Nt)) < 0.2] = 0  # extra result-shape zeros (user mask, cont_engine.h:526-528)
    Pr, Pc = grid
    for (an, bn, cnn) in ((a, b, cn), (None, None, None)):
        for r in range(Pr):
            for c in range(Pc):
                want = {k: p for k, p in O.summa_rank_schedule(Pr, Pc, r, c, Mt, Nt, Kt,
                                                               None if an is None else an < thr,
                                                               None if bn is None else bn < thr,
                                                               None if cnn is None else cnn < thr)}
                for k in range(Kt):
                    pi, pj = dev.build_pairlist(k, Pr, Pc, r, c, an, bn, cnn, Mt, Nt, Kt, float(thr))
                    assert list(zip(pi.tolist(), pj.tolist())) == want.get(k, [])


@pytest.mark.parametrize("grid", [(1, 1), (2, 2), (4, 2), (1, 8)])
def test_window_tile_lists_bit_exact(dev, grid):
    """tadev_build_tile_lists — the kernel the SUMMA driver runs per window (csrc/tilelist.cu) — against
    Summa::contract restated (contraction_eval.h:1311-1384): for a window of steps, the pairs the oracle schedules
    step by step, regrouped per local result tile in step order, must be EXACTLY the device's chains (every rank,
    sparse with extra result zeros and dense, whole-contraction windows and partial windows)."""
    rng = np.random.default_rng(17)
    Mt, Nt, Kt = 37, 29, 11
    thr = np.float32(0.5)
    a = np.where(rng.random((Mt, Kt)) < 0.3, 1.0, 0.0).astype(np.float32)
    b = np.where(rng.random((Kt, Nt)) < 0.3, 1.0, 0.0).astype(np.float32)
    cn = ((a @ b) > 0).astype(np.float32)
    cn[rng.random((Mt, Nt)) < 0.2] = 0
    Pr, Pc = grid
    for (an, bn, cnn) in ((a, b, cn), (None, None, None)):
        for r in range(Pr):
            for c in range(Pc):
                sched = dict(O.summa_rank_schedule(Pr, Pc, r, c, Mt, Nt, Kt, None if an is None else an < thr,
                                                   None if bn is None else bn < thr, None if cnn is None else cnn < thr))
                for window in (list(range(Kt)), [2, 3, 4, 5], [7], []):
                    chains = {}
                    for k in window:
                        for (i, j) in sched.get(k, []):
                            chains.setdefault((i, j), []).append(k)
                    gb, tk = dev.build_tile_lists(window, Pr, Pc, r, c, an, bn, cnn, Mt, Nt, Kt, float(thr))
                    ncl = (Nt - c + Pc - 1) // Pc if c < Nt else 0
                    nrl = (Mt - r + Pr - 1) // Pr if r < Mt else 0
                    assert len(gb) == nrl * ncl + 1 and gb[-1] == len(tk) == sum(len(v) for v in chains.values())
                    for li in range(nrl):
                        for lj in range(ncl):
                            g = li * ncl + lj
                            assert tk[gb[g]:gb[g + 1]].tolist() == chains.get((r + li * Pr, c + lj * Pc), [])
