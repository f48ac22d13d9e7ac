"""End-to-end contraction through the TiledArray-mirroring API (DistArray / expressions ->
ContEngine -> tadev_plan_contraction / tadev_shape_* / tadev_permute / tadev_summa_f64) against
the CPU oracle. These read like tests/expressions_impl.h:1808-2675 (cont, cont_permute,
scale_cont, cont_non_uniform, outer_product) and tests/dist_eval_contraction_eval.cpp:293-472.
"""
import numpy as np
import pytest

from oracle import ta_oracle as O
from tests import known_answers as KA
from tiledarray_b200.tiledarray import ContEngine, DistArray, SparseShape, TiledRange, TiledRange1, TiledArrayException

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _tr(*dims):
    return TiledRange([TiledRange1(*d) if not isinstance(d, TiledRange1) else d for d in dims])


def _uniform(extent, tile):
    return TiledRange1.make_uniform(extent, tile)


def _dense_array(world, tr, rng, integer=False):
    full = KA.int_tile(rng, tr.elements_shape) if integer else rng.uniform(-1, 1, tr.elements_shape)
    return DistArray(world, tr).init_from_numpy(full), full


def _check(world, target, lidx, ridx, trL, trR, trC, factor=1.0, integer=True, seed=0):
    rng = np.random.default_rng(seed)
    a, A = _dense_array(world, trL, rng, integer)
    b, B = _dense_array(world, trR, rng, integer)
    c = DistArray(world, trC)
    expr = a[lidx] * b[ridx]
    c[target] = expr if factor == 1.0 else factor * expr
    ref = factor * np.einsum(f"{lidx.replace(',', '')},{ridx.replace(',', '')}->{target.replace(',', '')}", A, B)
    got = c.to_numpy()
    if integer:
        assert np.array_equal(got, ref)
    else:
        assert O.rel_frobenius(got, ref) < TOL
    for x in (a, b, c):
        x.release()
    return ContEngine.last_stats


def test_cont_dense_matrix(world):
    """c("m,n") = a("m,k") * b("k,n") — examples/gemm/ta_dense.cpp:174; config 1 in miniature."""
    t = _uniform(96, 32)
    st = _check(world, "m,n", "m,k", "k,n", _tr(t, t), _tr(t, t), _tr(t, t))
    assert st.npairs == 27 and st.nlaunches == 1
    assert st.flops == 2.0 * 96 ** 3


def test_cont_fixture_prime_tiles(world):
    """Prime tile widths (tests/range_fixture.h:67-132): ragged, odd extents -> generic kernel."""
    d = TiledRange1(*KA.FIXTURE_BOUNDS)
    _check(world, "m,n", "m,k", "k,n", _tr(d, d), _tr(d, d), _tr(d, d))
    _check(world, "m,n", "m,k", "k,n", _tr(d, d), _tr(d, d), _tr(d, d), integer=False, factor=-1.5)


@pytest.mark.parametrize("target,lidx,ridx", [
    ("m,n", "k,m", "k,n"), ("m,n", "m,k", "n,k"), ("m,n", "k,m", "n,k"),  # BLAS-transpose forms (implicit permute)
    ("n,m", "m,k", "k,n"),                                                # cont_permute: result permutation
])
def test_cont_transposes_and_result_permute(world, target, lidx, ridx):
    dm, dk, dn = TiledRange1(0, 4, 10, 16), TiledRange1(0, 6, 8), TiledRange1(0, 2, 12, 20, 24)
    dims = {"m": dm, "k": dk, "n": dn}
    trL = _tr(*[dims[x] for x in lidx.split(",")])
    trR = _tr(*[dims[x] for x in ridx.split(",")])
    trC = _tr(*[dims[x] for x in target.split(",")])
    _check(world, target, lidx, ridx, trL, trR, trC)


def test_cont_ccsd_ppl_shape(world):
    """Config 4 in miniature: R("a,b,i,j") = T("c,d,i,j") * V("a,b,c,d") (examples/gemm/
    ta_cc_abcd.cpp:240) — both operands matrix_transpose, result permute; ragged
    occupied/virtual tilings like o=100/64, v=800/64."""
    o, v = TiledRange1(0, 6, 10), TiledRange1(0, 8, 16, 20)
    old = ContEngine.exchange_operands
    try:
        ContEngine.exchange_operands = False  # the reference's plan: opA = opB = T + result permutation
        st = _check(world, "a,b,i,j", "c,d,i,j", "a,b,c,d", _tr(v, v, o, o), _tr(v, v, v, v), _tr(v, v, o, o))
        assert st.permute_ms > 0.0  # the result permutation ran
        _check(world, "a,b,i,j", "c,d,i,j", "a,b,c,d", _tr(v, v, o, o), _tr(v, v, v, v), _tr(v, v, o, o), integer=False, factor=0.5)
        ContEngine.exchange_operands = True  # R[ab,ij] = V[ab,cd] T[cd,ij]: plain NN, no permutation at all
        st = _check(world, "a,b,i,j", "c,d,i,j", "a,b,c,d", _tr(v, v, o, o), _tr(v, v, v, v), _tr(v, v, o, o))
        assert st.permute_ms == 0.0
        _check(world, "a,b,i,j", "c,d,i,j", "a,b,c,d", _tr(v, v, o, o), _tr(v, v, v, v), _tr(v, v, o, o), integer=False, factor=0.5)
    finally:
        ContEngine.exchange_operands = old


def test_cont_general_permutes(world):
    """Config 5 in miniature: C("i,a,j,b") = A("i,k,a,c") * B("j,c,k,b") — both operands need
    explicit tile permutations (permtype general), result needs none."""
    s, b = TiledRange1(0, 2, 6), TiledRange1(0, 4, 8, 10)
    _check(world, "i,a,j,b", "i,k,a,c", "j,c,k,b", _tr(s, s, b, b), _tr(s, b, s, b), _tr(s, b, s, b))
    _check(world, "i,a,j,b", "i,k,a,c", "j,c,k,b", _tr(s, s, b, b), _tr(s, b, s, b), _tr(s, b, s, b), integer=False)


@pytest.mark.parametrize("spl,rb", [(0, 0), (1, 0), (2, 2), (1, 3)])
def test_cont_streamed_permutes(world, spl, rb):
    """Argument permutations performed just in time per SUMMA window by the permute provider (lazy
    tiles, dist_eval/array_eval.h:42,170) instead of up front: same exact result, no permuted copy."""
    s, b = TiledRange1(0, 2, 6), TiledRange1(0, 4, 8, 10)
    old = (ContEngine.stream_permutes, ContEngine.steps_per_launch, ContEngine.row_blocks)
    try:
        ContEngine.stream_permutes, ContEngine.steps_per_launch, ContEngine.row_blocks = True, spl, rb
        st = _check(world, "i,a,j,b", "i,k,a,c", "j,c,k,b", _tr(s, s, b, b), _tr(s, b, s, b), _tr(s, b, s, b))
        assert st.lazy_tiles >= 2 * 2 * 2 * 3 * 3 and st.permute_ms == 0.0
        _check(world, "a,i,b,j", "i,k,a,c", "j,c,k,b", _tr(s, s, b, b), _tr(s, b, s, b), _tr(b, s, b, s), integer=False)
    finally:
        ContEngine.stream_permutes, ContEngine.steps_per_launch, ContEngine.row_blocks = old


@pytest.mark.parametrize("lazy_side", ["left", "right", "both"])
def test_cont_lazy_operands(world, lazy_side):
    """Operands that are never stored: tiles are generated on the device from the counter RNG when the
    SUMMA window needs them (BASELINE config 4's V = 3.28 TB). Same values as the materialised array
    (fill_random uses the same generator), checked against host-regenerated tiles too."""
    from tests import util_rng
    dm, dk, dn = TiledRange1(0, 4, 10, 16, 18), TiledRange1(0, 6, 8, 20), TiledRange1(0, 2, 12, 20, 24)
    trL, trR, trC = _tr(dm, dk), _tr(dk, dn), _tr(dm, dn)

    def make(tr, seed, lazy):
        if lazy:
            return DistArray(world, tr, memory="lazy", lazy_seed=seed)
        return DistArray(world, tr).fill_random(seed)

    def host(tr, seed):
        full = np.zeros(tr.elements_shape)
        for o in range(tr.ntiles):
            idx = tr.tile_index(o)
            ext = tr.tile_extent(idx)
            full[tr.tile_slices(idx)] = util_rng.tile_fill(o, int(np.prod(ext)), seed).reshape(ext)
        return full

    a = make(trL, 11, lazy_side in ("left", "both"))
    b = make(trR, 12, lazy_side in ("right", "both"))
    ref = host(trL, 11) @ host(trR, 12)
    c = DistArray(world, trC)
    old = (ContEngine.steps_per_launch, ContEngine.row_blocks)
    try:
        for spl, rb in ((0, 0), (1, 2), (2, 4)):
            ContEngine.steps_per_launch, ContEngine.row_blocks = spl, rb
            c["m,n"] = a["m,k"] * b["k,n"]
            assert O.rel_frobenius(c.to_numpy(), ref) < TOL, (spl, rb)
            assert ContEngine.last_stats.lazy_tiles > 0
        ContEngine.steps_per_launch, ContEngine.row_blocks = 0, 0
        c["n,m"] = a["m,k"] * b["k,n"]  # exchanged operands: the lazy array changes sides
        assert O.rel_frobenius(c.to_numpy(), ref.T) < TOL
    finally:
        ContEngine.steps_per_launch, ContEngine.row_blocks = old
    assert O.rel_frobenius(a.find(1), host(trL, 11)[trL.tile_slices(trL.tile_index(1))]) < 1e-15
    for x in (a, b, c):
        x.release()


def test_cont_rank3_and_outer_product(world):
    d3, d2 = TiledRange1(0, 3, 5), TiledRange1(0, 2, 6, 7)
    _check(world, "a,b,c", "a,x,y", "y,x,b,c", _tr(d3, d2, d3), _tr(d3, d2, d2, d3), _tr(d3, d2, d3))
    _check(world, "a,b", "a", "b", _tr(d3), _tr(d2), _tr(d3, d2))  # outer product (k rank 0)
    _check(world, "i", "i,k", "k", _tr(d2, d3), _tr(d3), _tr(d2))  # matrix-vector


def _sparse_pair(world, trL, trR, density, rng, integer=True):
    def make(tr):
        full = KA.int_tile(rng, tr.elements_shape) if integer else rng.uniform(-1, 1, tr.elements_shape)
        norms = np.zeros(tr.tiles_shape, dtype=np.float32)
        for o in range(tr.ntiles):
            idx = tr.tile_index(o)
            if rng.random() < density:
                blk = full[tr.tile_slices(idx)]
                norms[idx] = np.float32(np.linalg.norm(blk))  # true Frobenius norm of the tile
            else:
                full[tr.tile_slices(idx)] = 0.0
        sh = SparseShape(world, norms, tr)
        return DistArray(world, tr, sh).init_from_numpy(full), full, norms
    return make(trL), make(trR)


@pytest.mark.parametrize("density", [0.6, 0.2])
def test_cont_sparse(world, density):
    """tests/dist_eval_contraction_eval.cpp:375-472 sparse_eval: block-sparse operands with true
    tile norms; result shape by device screening == oracle's, bit for bit; values exact; zero
    result tiles are absent and the dense reference block is all-zero (:434-445)."""
    rng = np.random.default_rng(int(density * 10))
    dm, dk, dn = _uniform(40, 8), _uniform(48, 8), _uniform(56, 8)
    trL, trR = _tr(dm, dk), _tr(dk, dn)
    (a, A, nA), (b, B, nB) = _sparse_pair(world, trL, trR, density, rng)
    c = DistArray(world, _tr(dm, dn))
    c["m,n"] = a["m,k"] * b["k,n"]
    # shapes: product vs oracle (bit-exact)
    oa = O.SparseShape.from_tile_norms(nA, O.TiledRange((O.TiledRange1(dm.bounds), O.TiledRange1(dk.bounds))))
    ob = O.SparseShape.from_tile_norms(nB, O.TiledRange((O.TiledRange1(dk.bounds), O.TiledRange1(dn.bounds))))
    assert np.array_equal(a.shape.norms.view(np.uint32), oa.norms.view(np.uint32))
    oc = oa.gemm(ob, 1.0, O.GemmHelper(0, 0, 2, 2, 2))
    assert np.array_equal(c.shape.norms.view(np.uint32), oc.norms.view(np.uint32))
    assert c.shape.zero_tile_count == oc.zero_tile_count
    ref = A @ B
    assert np.array_equal(c.to_numpy(), ref)
    for o in range(c.trange.ntiles):
        if c.is_zero(o):
            assert o not in c.tiles and not ref[c.trange.tile_slices(c.trange.tile_index(o))].any()
    st = ContEngine.last_stats
    want_pairs = int(((oa.norms >= np.float32(oa.threshold)).astype(int) @ (ob.norms >= np.float32(ob.threshold)).astype(int)).sum())
    assert st.npairs == want_pairs
    for x in (a, b, c):
        x.release()


def test_cont_sparse_permuted_4index(world):
    """Sparse + general permutes + result permute: shapes are permuted with the tiles
    (SparseShape::perm, sparse_shape.h:1222) and the tile list still matches the data."""
    rng = np.random.default_rng(31)
    s, b = TiledRange1(0, 2, 6), TiledRange1(0, 4, 8, 10)
    trL, trR = _tr(s, s, b, b), _tr(s, b, s, b)
    (a, A, _), (bb, B, _) = _sparse_pair(world, trL, trR, 0.5, rng, integer=False)
    c = DistArray(world, _tr(b, s, b, s))
    c["a,i,b,j"] = 2.0 * (a["i,k,a,c"] * bb["j,c,k,b"])
    ref = 2.0 * np.einsum("ikac,jckb->aibj", A, B)
    assert O.rel_frobenius(c.to_numpy(), ref) < TOL
    for x in (a, bb, c):
        x.release()


def test_cont_reference_style_dense_fill(world):
    """examples/device/ta_dense_device.cpp verification: a.fill(va), b.fill(vb) => every element of
    c equals Nk*va*vb; config-1 tiling (tile 256) at N=1024."""
    t = _uniform(1024, 256)
    a = DistArray(world, _tr(t, t)).fill(1.0)
    b = DistArray(world, _tr(t, t)).fill(0.5)
    c = DistArray(world, _tr(t, t))
    c["m,n"] = a["m,k"] * b["k,n"]
    assert np.array_equal(c.to_numpy(), np.full((1024, 1024), 512.0))
    for x in (a, b, c):
        x.release()


def test_cont_accumulate(world):
    """c("m,n") += a("m,k") * b("k,n") (expressions_impl.h cont_plus_reduce / scale_cont forms): the product
    is accumulated into the existing result tiles (beta = 1 in the GEMM epilogue), dense and sparse."""
    rng = np.random.default_rng(5)
    t = TiledRange1(0, 8, 20, 32)
    a, A = _dense_array(world, _tr(t, t), rng, True)
    b, B = _dense_array(world, _tr(t, t), rng, True)
    c = DistArray(world, _tr(t, t))
    c["m,n"] = a["m,k"] * b["k,n"]
    c["m,n"] += 2.0 * (a["m,k"] * b["k,n"])
    c["m,n"] += a["m,k"] * b["k,n"]
    assert np.array_equal(c.to_numpy(), 4.0 * (A @ B))
    (sa, SA, _), (sb, SB, _) = _sparse_pair(world, _tr(t, t), _tr(t, t), 0.5, rng)
    sc = DistArray(world, _tr(t, t))
    sc["m,n"] = sa["m,k"] * sb["k,n"]
    sc["m,n"] += sa["m,k"] * sb["k,n"]
    assert np.array_equal(sc.to_numpy(), 2.0 * (SA @ SB))
    for x in (a, b, c, sa, sb, sc):
        x.release()


def test_elementwise_expressions_dense(world):
    """AddEngine / SubtEngine / ScalEngine / Hadamard MultEngine on device tiles (tests/expressions_impl.h
    add, subt, scale, mult, permute blocks): one batched launch per expression; operands in a different
    index order are permuted first. Exact on integer data."""
    rng = np.random.default_rng(77)
    d1, d2 = TiledRange1(0, 3, 10, 16), TiledRange1(0, 5, 8, 21)
    a, A = _dense_array(world, _tr(d1, d2), rng, True)
    b, B = _dense_array(world, _tr(d1, d2), rng, True)
    bt, BT = _dense_array(world, _tr(d2, d1), rng, True)
    c = DistArray(world, _tr(d1, d2))
    c["i,j"] = a["i,j"] + b["i,j"]
    assert np.array_equal(c.to_numpy(), A + B)
    c["i,j"] = a["i,j"] - b["i,j"]
    assert np.array_equal(c.to_numpy(), A - B)
    c["i,j"] = 2.0 * a["i,j"] + 3.0 * bt["j,i"]
    assert np.array_equal(c.to_numpy(), 2 * A + 3 * BT.T)
    c["i,j"] = -(a["i,j"] - 2.0 * b["i,j"])
    assert np.array_equal(c.to_numpy(), -(A - 2 * B))
    c["i,j"] = a["i,j"] * b["i,j"]  # Hadamard: every index shared and kept
    assert np.array_equal(c.to_numpy(), A * B)
    c["i,j"] = 0.5 * (a["i,j"] * bt["j,i"])
    assert np.array_equal(c.to_numpy(), 0.5 * A * BT.T)
    c["i,j"] = 4.0 * a["i,j"]
    assert np.array_equal(c.to_numpy(), 4 * A)
    ct = DistArray(world, _tr(d2, d1))
    ct["j,i"] = a["i,j"]  # pure permutation
    assert np.array_equal(ct.to_numpy(), A.T)
    c["i,j"] = c["i,j"] + a["i,j"]  # the result is also an operand
    assert np.array_equal(c.to_numpy(), 5 * A)
    x, X = _dense_array(world, _tr(d1, d2, d1), rng, True)
    y, Y = _dense_array(world, _tr(d1, d1, d2), rng, True)
    z = DistArray(world, _tr(d2, d1, d1))
    z["b,a,c"] = x["a,b,c"] - 2.0 * y["c,a,b"]
    assert np.array_equal(z.to_numpy(), np.einsum("abc->bac", X) - 2 * np.einsum("cab->bac", Y))
    for t in (a, b, bt, c, ct, x, y, z):
        t.release()


def test_elementwise_expressions_sparse_and_truncate(world):
    """Block-sparse add / subt / Hadamard: result shapes by SparseShape::add / mult (bit-exact vs the oracle),
    absent tiles act as zeros, result tiles exist exactly where the result shape is non-zero; truncate()
    recomputes the shape from the true norms (dist_array.h:1553) and drops tiles that cancelled."""
    rng = np.random.default_rng(13)
    d = TiledRange1(0, 4, 8, 14, 20)
    tr = _tr(d, d)
    (a, A, nA), (b, B, nB) = _sparse_pair(world, tr, tr, 0.5, rng)
    otr = O.TiledRange((O.TiledRange1(d.bounds), O.TiledRange1(d.bounds)))
    oa, ob = O.SparseShape.from_tile_norms(nA, otr), O.SparseShape.from_tile_norms(nB, otr)
    c = DistArray(world, tr)
    c["i,j"] = a["i,j"] + b["i,j"]
    assert np.array_equal(c.shape.norms.view(np.uint32), oa.add(ob).norms.view(np.uint32))
    assert np.array_equal(c.to_numpy(), A + B)
    assert sorted(c.tiles) == [o for o in range(tr.ntiles) if not c.shape.is_zero(o)]
    c["i,j"] = 2.0 * a["i,j"] - 0.5 * b["i,j"]
    assert np.array_equal(c.shape.norms.view(np.uint32), oa.scale(2.0).add(ob.scale(0.5)).norms.view(np.uint32))
    assert np.array_equal(c.to_numpy(), 2 * A - 0.5 * B)
    c["i,j"] = 3.0 * (a["i,j"] * b["i,j"])
    assert np.array_equal(c.shape.norms.view(np.uint32), oa.mult(ob, 3.0).norms.view(np.uint32))
    assert np.array_equal(c.to_numpy(), 3 * A * B)
    assert c.shape.zero_tile_count == oa.mult(ob, 3.0).zero_tile_count
    # a - a: the estimated shape keeps every tile of a, truncate() finds that they are all zero
    c["i,j"] = a["i,j"] - a["i,j"]
    assert len(c.tiles) == len(a.tiles) and not c.to_numpy().any()
    c.truncate()
    assert len(c.tiles) == 0 and c.shape.zero_tile_count == tr.ntiles
    # truncate on real data reproduces the shape built from the true norms
    c["i,j"] = a["i,j"] + b["i,j"]
    c.truncate()
    want = O.SparseShape.from_tile_norms(np.array([[np.linalg.norm((A + B)[tr.tile_slices((i, j))]) for j in range(4)]
                                                   for i in range(4)], dtype=np.float32), otr)
    np.testing.assert_allclose(c.shape.norms, want.norms, rtol=1e-6)
    assert (c.shape.norms == 0).tolist() == (want.norms == 0).tolist()
    for t in (a, b, c):
        t.release()


@pytest.mark.parametrize("target,lidx,ridx", [
    ("b,i,j", "b,i,k", "b,k,j"),          # canonical batched GEMM
    ("i,b,j", "i,k,b", "k,b,j"),          # every array needs a permutation
    ("b,i,j", "b,i", "b,j"),              # batched outer product (nothing contracted)
    ("b", "b,k", "b,k"),                  # no external index: row-wise dot products
    ("b,c,i", "c,b,i,k", "b,k,c"),        # two fused indices, right operand has no external index
])
def test_general_products_dense(world, target, lidx, ridx):
    """General products (fused + contracted + free indices; TensorProduct::General, cont_engine.h:679-1100,
    BatchedContractReduce): the fused indices are batch indices of one grouped-GEMM launch. Exact (int)."""
    rng = np.random.default_rng(len(target) * 13 + len(lidx))
    dims = {"b": TiledRange1(0, 3, 5), "c": TiledRange1(0, 2, 6), "i": TiledRange1(0, 4, 10), "j": TiledRange1(0, 6, 8, 14),
            "k": TiledRange1(0, 2, 8, 12)}
    trL = _tr(*[dims[x] for x in lidx.split(",")])
    trR = _tr(*[dims[x] for x in ridx.split(",")])
    trC = _tr(*[dims[x] for x in target.split(",")])
    st = _check(world, target, lidx, ridx, trL, trR, trC)
    assert st.nlaunches == 1
    _check(world, target, lidx, ridx, trL, trR, trC, integer=False, factor=-0.75)


def _patterned(world, tr, seed):
    """make_patterned_array of tests/general_product.cpp:149-168: v = seed + sum_d 0.1^d * (x_d + 1)."""
    idx = np.indices(tr.elements_shape).astype(np.float64)
    full = np.full(tr.elements_shape, float(seed))
    scale = 1.0
    for d in range(tr.rank):
        full = full + scale * (idx[d] + 1.0)
        scale *= 0.1
    return DistArray(world, tr).init_from_numpy(full), full


def test_general_products_reference_cases(world):
    """tests/general_product.cpp:183-294 on the device engine, same tilings and data pattern, against einsum:
    batched GEMM, permuted arguments, batched outer product, fused broadcast on either side, non-canonical
    target; plus the neighbouring pure contraction and pure Hadamard forms of the same arrays."""
    a, A = _patterned(world, _tr((0, 2, 5), (0, 3, 4), (0, 2, 6, 7)), 1.0)      # b, i, j
    b, B = _patterned(world, _tr((0, 2, 5), (0, 2, 6, 7), (0, 4, 5)), 2.0)      # b, j, k
    c = DistArray(world, _tr((0, 2, 5), (0, 3, 4), (0, 4, 5)))
    c["b,i,k"] = a["b,i,j"] * b["b,j,k"]
    assert O.rel_frobenius(c.to_numpy(), np.einsum("bij,bjk->bik", A, B)) < TOL
    d = DistArray(world, _tr((0, 3, 4), (0, 4, 5)))
    d["i,k"] = a["b,i,j"] * b["b,j,k"]                                          # pure contraction
    assert O.rel_frobenius(d.to_numpy(), np.einsum("bij,bjk->ik", A, B)) < TOL
    e = DistArray(world, a.trange)
    e["b,i,j"] = a["b,i,j"] * a["b,i,j"]                                        # pure Hadamard
    assert O.rel_frobenius(e.to_numpy(), A * A) < TOL
    a2, A2 = _patterned(world, _tr((0, 3, 4), (0, 2, 6, 7), (0, 2, 5)), 1.0)    # i, j, b
    b2, B2 = _patterned(world, _tr((0, 4, 5), (0, 2, 6, 7), (0, 2, 5)), 2.0)    # k, j, b
    c["b,i,k"] = a2["i,j,b"] * b2["k,j,b"]
    assert O.rel_frobenius(c.to_numpy(), np.einsum("ijb,kjb->bik", A2, B2)) < TOL
    a3, A3 = _patterned(world, _tr((0, 2, 5), (0, 3, 4)), 1.0)                  # b, i
    b3, B3 = _patterned(world, _tr((0, 2, 5), (0, 4, 5)), 2.0)                  # b, k
    c["b,i,k"] = a3["b,i"] * b3["b,k"]                                          # batched outer product
    assert O.rel_frobenius(c.to_numpy(), np.einsum("bi,bk->bik", A3, B3)) < TOL
    s1, S1 = _patterned(world, _tr((0, 2, 5)), 1.0)                             # b
    cc = DistArray(world, b3.trange)
    cc["b,k"] = s1["b"] * b3["b,k"]                                             # fused broadcast, left fully fused
    assert O.rel_frobenius(cc.to_numpy(), S1[:, None] * B3) < TOL
    cc["b,k"] = b3["b,k"] * s1["b"]                                             # ... and on the right
    assert O.rel_frobenius(cc.to_numpy(), S1[:, None] * B3) < TOL
    x, X = _patterned(world, _tr((0, 2, 4), (0, 3, 5)), 1.0)                    # orbital x auxiliary
    w = DistArray(world, _tr((0, 2, 4), (0, 2, 4), (0, 3, 5)))
    w["p,q,r1"] = x["p,r1"] * x["q,r1"]                                         # non-canonical target
    assert O.rel_frobenius(w.to_numpy(), np.einsum("pr,qr->pqr", X, X)) < TOL
    for t in (a, b, c, d, e, a2, b2, a3, b3, s1, cc, x, w):
        t.release()


def test_general_product_sparse_and_accumulate(world):
    """Block-sparse general product: result shape = SparseShape::gemm_batched (sparse_shape.h:1707-1900), i.e. the
    oracle's norm GEMM slab by slab, bit-exact; zero result tiles absent; c += accumulates."""
    rng = np.random.default_rng(3)
    db, di, dk, dj = TiledRange1(0, 3, 5, 9), TiledRange1(0, 4, 10), TiledRange1(0, 2, 8, 12), TiledRange1(0, 6, 8, 14)
    trL, trR, trC = _tr(db, di, dk), _tr(db, dk, dj), _tr(db, di, dj)
    (a, A, nA), (b, B, nB) = _sparse_pair(world, trL, trR, 0.5, rng)
    c = DistArray(world, trC)
    c["b,i,j"] = a["b,i,k"] * b["b,k,j"]
    ref = np.einsum("bik,bkj->bij", A, B)
    assert np.array_equal(c.to_numpy(), ref)
    otrL = O.TiledRange(tuple(O.TiledRange1(d.bounds) for d in (db, di, dk)))
    otrR = O.TiledRange(tuple(O.TiledRange1(d.bounds) for d in (db, dk, dj)))
    oa, ob = O.SparseShape.from_tile_norms(nA, otrL), O.SparseShape.from_tile_norms(nB, otrR)
    for h in range(db.ntiles):  # slab by slab: the folded (fused-mode-free) norm GEMM
        sa = O.SparseShape(oa.norms[h], oa.size_vectors[1:], 0)
        sb = O.SparseShape(ob.norms[h], ob.size_vectors[1:], 0)
        want = sa.gemm(sb, 1.0, O.GemmHelper(0, 0, 2, 2, 2))
        assert np.array_equal(c.shape.norms[h].view(np.uint32), want.norms.view(np.uint32)), h
    assert sorted(c.tiles) == [o for o in range(trC.ntiles) if not c.shape.is_zero(o)]
    c["b,i,j"] += 2.0 * (a["b,i,k"] * b["b,k,j"])
    assert np.array_equal(c.to_numpy(), 3.0 * ref)
    for x in (a, b, c):
        x.release()


def test_cont_errors(world):
    t, u = _uniform(8, 4), _uniform(8, 2)
    a = DistArray(world, _tr(t, t)).fill(1.0)
    b = DistArray(world, _tr(u, t)).fill(1.0)
    c = DistArray(world, _tr(t, t))
    with pytest.raises(TiledArrayException):  # inner tilings not congruent (TA_ASSERT in kernels.h:117-131)
        c["m,n"] = a["m,k"] * b["k,n"]
    with pytest.raises(TiledArrayException):
        c["m,n"] = a["m,k,z"] * b["k,n"]
    for x in (a, b, c):
        x.release()


@pytest.mark.parametrize("where", [("host", "host", "host"), ("host", "device", "device"), ("device", "host", "host"),
                                   ("device", "device", "host")])
@pytest.mark.parametrize("sparse", [False, True])
def test_cont_host_resident_arrays(world, where, sparse):
    """Host-resident (pinned) operands/result: the SUMMA driver streams operand panels to the
    device window by window and returns the result in row blocks (tadev.h TADEV_SUMMA_*_ON_HOST);
    small windows / several row blocks / ragged tiles exercise every pipeline edge. Exact (int)."""
    rng = np.random.default_rng(17 + sparse)
    dm, dk, dn = TiledRange1(0, 4, 10, 16, 18, 30, 32), TiledRange1(0, 6, 8, 20, 22, 30), TiledRange1(0, 2, 12, 20, 24)
    trL, trR, trC = _tr(dm, dk), _tr(dk, dn), _tr(dm, dn)

    def make(tr, mem):
        full = KA.int_tile(rng, tr.elements_shape)
        sh = None
        if sparse:
            norms = np.zeros(tr.tiles_shape, dtype=np.float32)
            for o in range(tr.ntiles):
                idx = tr.tile_index(o)
                if rng.random() < 0.6:
                    norms[idx] = np.float32(np.linalg.norm(full[tr.tile_slices(idx)]))
                else:
                    full[tr.tile_slices(idx)] = 0.0
            sh = SparseShape(world, norms, tr)
        return DistArray(world, tr, sh, memory=mem).init_from_numpy(full), full

    a, A = make(trL, where[0])
    b, B = make(trR, where[1])
    c = DistArray(world, trC, memory=where[2])
    old = (ContEngine.steps_per_launch, ContEngine.row_blocks)
    try:
        for spl, rb in ((0, 0), (2, 3), (1, 6)):
            ContEngine.steps_per_launch, ContEngine.row_blocks = spl, rb
            c["m,n"] = a["m,k"] * b["k,n"]
            assert np.array_equal(c.to_numpy(), A @ B), (spl, rb)
            st = ContEngine.last_stats
            if where[2] == "host":
                assert st.d2h_bytes == sum(t.nbytes for t in c.tiles.values())
            if where[0] == "host" and not sparse and where[2] != "host":
                assert st.h2d_bytes >= A.nbytes  # every A tile uploaded exactly once
    finally:
        ContEngine.steps_per_launch, ContEngine.row_blocks = old
    for x in (a, b, c):
        x.release()


def _otr(*dims):
    return O.TiledRange(tuple(O.TiledRange1(d.bounds) for d in dims))


@pytest.mark.parametrize("permuted", [False, True])
def test_cont_set_shape_mask(world, permuted):
    """c("m,n") = (a("m,k") * b("k,n")).set_shape(mask) — the reference's own block-sparse example expression
    (examples/gemm/ta_sparse.cpp:190; ContEngine::init_struct applies SparseShape::mask after make_shape,
    cont_engine.h:526-528, sparse_shape.h:653-676). Result shape bit-exact vs the oracle's gemm(...).mask(...),
    masked tiles absent, values exact, and only pairs that feed surviving result tiles are executed."""
    rng = np.random.default_rng(77)
    dm, dk, dn = _uniform(40, 8), _uniform(48, 8), _uniform(56, 8)
    trL, trR = _tr(dm, dk), _tr(dk, dn)
    (a, A, nA), (b, B, nB) = _sparse_pair(world, trL, trR, 0.6, rng)
    trC = _tr(dn, dm) if permuted else _tr(dm, dn)
    mnorms = np.where(rng.random(trC.tiles_shape) < 0.5, np.float32(3.0), np.float32(0.0)).astype(np.float32)
    mask = SparseShape(world, mnorms, trC)
    c = DistArray(world, trC)
    if permuted:
        c["n,m"] = (a["m,k"] * b["k,n"]).set_shape(mask)
    else:
        c["m,n"] = (a["m,k"] * b["k,n"]).set_shape(mask)
    oa = O.SparseShape.from_tile_norms(nA, _otr(dm, dk))
    ob = O.SparseShape.from_tile_norms(nB, _otr(dk, dn))
    om = O.SparseShape.from_tile_norms(mnorms, _otr(dn, dm) if permuted else _otr(dm, dn))
    oc = oa.gemm(ob, 1.0, O.GemmHelper(0, 0, 2, 2, 2), perm=[1, 0] if permuted else None).mask(om)
    assert np.array_equal(c.shape.norms.view(np.uint32), oc.norms.view(np.uint32))
    assert c.shape.zero_tile_count == oc.zero_tile_count
    ref = (A @ B).T if permuted else A @ B
    keep = np.zeros_like(ref)
    for o in range(trC.ntiles):
        sl = trC.tile_slices(trC.tile_index(o))
        if oc.is_zero(o):
            assert o not in c.tiles
        else:
            keep[sl] = ref[sl]
    assert np.array_equal(c.to_numpy(), keep)
    nzc = (oc.norms >= np.float32(oc.threshold))
    nzc = nzc.T if permuted else nzc
    pa, pb = (oa.norms >= np.float32(oa.threshold)).astype(int), (ob.norms >= np.float32(ob.threshold)).astype(int)
    assert ContEngine.last_stats.npairs == int(((pa @ pb) * nzc).sum())
    for x in (a, b, c):
        x.release()
    # dense policy has no shape override
    t = _uniform(16, 8)
    da, _ = _dense_array(world, _tr(t, t), rng)
    with pytest.raises(TiledArrayException):
        DistArray(world, _tr(t, t))["m,n"] = (da["m,k"] * da["k,n"]).set_shape(SparseShape(world, np.ones((2, 2), np.float32), _tr(t, t)))
    da.release()


def test_cont_accumulate_into_foreign_layouts(world):
    """c += a*b must use the EXISTING array's own tile pointers: arrays whose arena order differs from the engine's
    layout, results of an earlier product with a result permutation, sparse results with a superset shape, products
    that need a result permutation, and products with tiles c does not have yet (shape union)."""
    rng = np.random.default_rng(11)
    t = TiledRange1(0, 8, 20, 32)
    tr = _tr(t, t)
    a, A = _dense_array(world, tr, rng, True)
    b, B = _dense_array(world, tr, rng, True)
    # (a) c created with a reversed arena order
    c0full = KA.int_tile(rng, tr.elements_shape)
    c = DistArray(world, tr, arena_order=list(range(tr.ntiles))[::-1]).init_from_numpy(c0full)
    c["m,n"] += a["m,k"] * b["k,n"]
    assert np.array_equal(c.to_numpy(), c0full + A @ B)
    # (b) c produced by a product that needed a result permutation (its offsets follow the GEMM key order)
    u = TiledRange1(0, 6, 16)
    x, X = _dense_array(world, _tr(t, u), rng, True)
    y, Y = _dense_array(world, _tr(u, t), rng, True)
    d = DistArray(world, _tr(t, t))
    ContEngine.exchange_operands = False
    try:
        d["n,m"] = x["m,k"] * y["k,n"]
        assert ContEngine.last_stats.swapped is False
        d["m,n"] += a["m,k"] * b["k,n"]                      # plain product into the permuted-layout array
        assert np.array_equal(d.to_numpy(), (X @ Y).T + A @ B)
        d["n,m"] += x["m,k"] * y["k,n"]                      # (d) += of a product that needs a result permutation
        assert np.array_equal(d.to_numpy(), 2 * (X @ Y).T + A @ B)
    finally:
        ContEngine.exchange_operands = True
    # (c) sparse c with a superset shape, then (e) a product with tiles c lacks (union)
    (sa, SA, nA), (sb, SB, nB) = _sparse_pair(world, tr, tr, 0.4, rng)
    full = KA.int_tile(rng, tr.elements_shape)
    sup = DistArray(world, tr, SparseShape(world, np.full(tr.tiles_shape, 50.0, np.float32), tr)).init_from_numpy(full)
    sup["m,n"] += sa["m,k"] * sb["k,n"]
    assert np.array_equal(sup.to_numpy(), full + SA @ SB)
    few = np.zeros(tr.tiles_shape, np.float32)
    few[0, 0] = 40.0
    sub_full = np.zeros(tr.elements_shape)
    sub_full[tr.tile_slices((0, 0))] = KA.int_tile(rng, tr.tile_extent((0, 0)))
    sub = DistArray(world, tr, SparseShape(world, few, tr)).init_from_numpy(sub_full)
    sub["m,n"] += sa["m,k"] * sb["k,n"]
    assert np.array_equal(sub.to_numpy(), sub_full + SA @ SB)
    prod_nz = ((nA > 0).astype(int) @ (nB > 0).astype(int)) > 0
    for o in range(tr.ntiles):
        i, j = tr.tile_index(o)
        assert (o in sub.tiles) == bool(prod_nz[i, j] or (i, j) == (0, 0))
        assert sub.is_zero(o) == (o not in sub.tiles)
    for z in (a, b, c, x, y, d, sa, sb, sup, sub):
        z.release()


def test_c3_shaped_every_tile(world):
    """BASELINE config 3's structure at a reduced tile size: 128 x 128 tile grid, 10 % random tile density, tile 8
    (N = 1024). EVERY result tile is compared with the oracle (dense numpy product of the block-sparse operands),
    the result shape with the oracle's SparseShape::gemm bit for bit, and the executed tile-pair count with the
    tile lists (ta_sparse.cpp:193 flop convention)."""
    nt, T = 128, 8
    t = _uniform(nt * T, T)
    tr = _tr(t, t)
    rng = np.random.default_rng(3)

    def make(seed):
        r = np.random.default_rng(seed)
        nz = r.permutation(nt * nt)[: int(0.10 * nt * nt)]
        full = np.zeros((nt * T, nt * T))
        norms = np.zeros((nt, nt), np.float32)
        for o in nz:
            i, j = divmod(int(o), nt)
            blk = KA.int_tile(rng, (T, T))
            full[i * T:(i + 1) * T, j * T:(j + 1) * T] = blk
            norms[i, j] = np.float32(np.linalg.norm(blk))
        return DistArray(world, tr, SparseShape(world, norms, tr)).init_from_numpy(full), full, norms

    a, A, nA = make(5)
    b, B, nB = make(6)
    c = DistArray(world, tr)
    c["m,n"] = a["m,k"] * b["k,n"]
    otr = _otr(t, t)
    oa, ob = O.SparseShape.from_tile_norms(nA, otr), O.SparseShape.from_tile_norms(nB, otr)
    oc = oa.gemm(ob, 1.0, O.GemmHelper(0, 0, 2, 2, 2))
    assert np.array_equal(c.shape.norms.view(np.uint32), oc.norms.view(np.uint32))
    ref = A @ B
    got = c.to_numpy()
    assert np.array_equal(got, ref)  # every tile (absent tiles are zero blocks of to_numpy)
    nz_tiles = int((oc.norms >= np.float32(oc.threshold)).sum())
    assert len(c.tiles) == nz_tiles
    pairs = int((((oa.norms >= np.float32(oa.threshold)).astype(int)) @ ((ob.norms >= np.float32(ob.threshold)).astype(int))).sum())
    assert ContEngine.last_stats.npairs == pairs
    assert ContEngine.last_stats.flops == 2.0 * pairs * T ** 3
    for x in (a, b, c):
        x.release()


@pytest.mark.parametrize("case", ["dense_ragged", "sparse", "permuted", "host"])
def test_device_lists_match_host_lists(world, case, monkeypatch):
    """The SUMMA driver's device-built tile lists (default) and the host-built lists (TADEV_HOST_LISTS=1) must give
    the same launches: identical results bit for bit, identical pair counts and flop counts."""
    rng = np.random.default_rng(23)
    if case == "dense_ragged":
        d = TiledRange1(*KA.FIXTURE_BOUNDS)
        a, _ = _dense_array(world, _tr(d, d), rng)
        b, _ = _dense_array(world, _tr(d, d), rng)
        args = ("m,n", "m,k", "k,n", _tr(d, d))
    elif case == "sparse":
        dm, dk, dn = _uniform(96, 16), _uniform(128, 16), _uniform(80, 16)
        (a, _, _), (b, _, _) = _sparse_pair(world, _tr(dm, dk), _tr(dk, dn), 0.4, rng, integer=False)
        args = ("m,n", "m,k", "k,n", _tr(dm, dn))
    elif case == "permuted":
        s, v = TiledRange1(0, 4, 8), TiledRange1(0, 16, 32, 40)
        a, _ = _dense_array(world, _tr(s, s, v, v), rng)
        b, _ = _dense_array(world, _tr(s, v, s, v), rng)
        args = ("i,a,j,b", "i,k,a,c", "j,c,k,b", _tr(s, v, s, v))
    else:
        t = _uniform(256, 64)
        full_a, full_b = rng.uniform(-1, 1, (256, 256)), rng.uniform(-1, 1, (256, 256))
        a = DistArray(world, _tr(t, t), memory="host").init_from_numpy(full_a)
        b = DistArray(world, _tr(t, t), memory="host").init_from_numpy(full_b)
        args = ("m,n", "m,k", "k,n", _tr(t, t))
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("TADEV_HOST_LISTS", mode)
        c = DistArray(world, args[3], memory="host" if case == "host" else "device")
        c[args[0]] = a[args[1]] * b[args[2]]
        st = ContEngine.last_stats
        out[mode] = (c.to_numpy(), st.npairs, st.flops, st.nlaunches, sorted(c.tiles))
        c.release()
    assert np.array_equal(out["0"][0], out["1"][0])
    assert out["0"][1:] == out["1"][1:]
    for x in (a, b):
        x.release()


@pytest.mark.parametrize("where", ["host", "lazy", "mixed"])
def test_cont_permuted_host_and_lazy_operands(world, where):
    """Operands that are NOT device-resident and need an explicit argument permutation (BASELINE config 5's index
    pattern): host-resident tiles are uploaded and permuted, lazy tiles generated and permuted, tile by tile when a
    SUMMA window asks for them (what ArrayEvalImpl does per tile, dist_eval/array_eval.h:170,330); also a
    host-resident result that needs a result permutation."""
    from tests import util_rng
    s, v = TiledRange1(0, 2, 6), TiledRange1(0, 8, 16, 20)
    trA, trB = _tr(s, s, v, v), _tr(s, v, s, v)

    def host_full(tr, seed):
        full = np.zeros(tr.elements_shape)
        for o in range(tr.ntiles):
            idx = tr.tile_index(o)
            ext = tr.tile_extent(idx)
            full[tr.tile_slices(idx)] = util_rng.tile_fill(o, int(np.prod(ext)), seed).reshape(ext)
        return full

    A, B = host_full(trA, 41), host_full(trB, 42)
    if where == "host":
        a = DistArray(world, trA, memory="host").fill_random(41)
        b = DistArray(world, trB, memory="host").fill_random(42)
    elif where == "lazy":
        a = DistArray(world, trA, memory="lazy", lazy_seed=41)
        b = DistArray(world, trB, memory="lazy", lazy_seed=42)
    else:
        a = DistArray(world, trA, memory="host").fill_random(41)
        b = DistArray(world, trB, memory="lazy", lazy_seed=42)
    ref = np.einsum("ikac,jckb->iajb", A, B)
    for spl in (0, 1):
        ContEngine.steps_per_launch = spl
        try:
            c = DistArray(world, _tr(s, v, s, v))
            c["i,a,j,b"] = a["i,k,a,c"] * b["j,c,k,b"]
            assert O.rel_frobenius(c.to_numpy(), ref) < TOL, (where, spl)
            assert ContEngine.last_stats.lazy_tiles > 0
            c.release()
            # host-resident result in another index order: the product is permuted on the device, then copied back
            ch = DistArray(world, _tr(v, s, v, s), memory="host")
            ch["a,i,b,j"] = a["i,k,a,c"] * b["j,c,k,b"]
            assert O.rel_frobenius(ch.to_numpy(), ref.transpose(1, 0, 3, 2)) < TOL, (where, spl)
            ch.release()
        finally:
            ContEngine.steps_per_launch = 0
    for x in (a, b):
        x.release()
