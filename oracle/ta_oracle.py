"""CPU oracle for the TiledArray contraction path — TEST INFRASTRUCTURE ONLY.

This module restates, in numpy / plain Python, the reference algorithm of the contraction hot
path of ValeevGroup/tiledarray. It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it. Nothing in the
product package ``tiledarray_b200`` imports it, and there is no code path by which a product
call can be served by this module.

Parity pinning: the reference itself cannot be built here (needs MADNESS, Boost,
Eigen, BTAS, BLAS++, range-v3, Umpire; none present, no network), so the oracle is pinned
against the known answers the reference's own tests hold for this path — restated in
``tests/test_oracle_*.py`` with file:line citations (tests/librett.cpp, tests/proc_grid.cpp,
tests/cyclic_pmap.cpp, tests/sparse_shape.cpp, tests/permutation.cpp, tests/math_blas.cpp,
tests/tile_op_contract_reduce.cpp, tests/dist_eval_contraction_eval.cpp).
The FP64 arithmetic of the path lives in a third-party dependency that is absent from the
reference tree: BLAS++ ``::blas::gemm`` (reached at src/TiledArray/math/blas.h:171-177; arrives
un-pinned through BTAS @ 245e49f117981d6124e0f1aa0d1ae72f1c16318b) -> vendor DGEMM. Its published
contract, C := alpha*op(A)*op(B) + beta*C, is restated with numpy's bundled OpenBLAS 0.3.30.

All citations are file:line under /root/reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

FLT_EPSILON = float(np.finfo(np.float32).eps)  # SparseShape default threshold (sparse_shape.h:1941)

# --------------------------------------------------------------------------------------------
# Permutation (src/TiledArray/permutation.h:69-79): image form, result[perm[i]] = arg[i]


def permute_array(perm: Sequence[int], arg: Sequence) -> list:
    """permute_array (permutation.h:67-79)."""
    assert sorted(perm) == list(range(len(perm))) and len(perm) == len(arg)
    out = [None] * len(arg)
    for i, p in enumerate(perm):
        out[p] = arg[i]
    return out


def perm_inverse(perm: Sequence[int]) -> List[int]:
    inv = [0] * len(perm)
    for i, p in enumerate(perm):
        inv[p] = i
    return inv


def tile_permute(x: np.ndarray, perm: Sequence[int]) -> np.ndarray:
    """Out-of-place tile permutation with TiledArray semantics (tensor/permute.h:118-209):
    result.extent[perm[i]] = arg.extent[i] and result[permute(idx)] = arg[idx]."""
    if len(perm) == 0:
        return x.copy()
    return np.ascontiguousarray(np.transpose(x, perm_inverse(perm)))


def tile_permute_loops(x: np.ndarray, perm: Sequence[int]) -> np.ndarray:
    """Same as tile_permute but by explicit index loops (small cases; independent check)."""
    out = np.empty(permute_array(perm, x.shape), dtype=x.dtype)
    for idx in np.ndindex(*x.shape):
        out[tuple(permute_array(perm, idx))] = x[idx]
    return out


# --------------------------------------------------------------------------------------------
# TiledRange1 / TiledRange (src/TiledArray/tiled_range1.h:47, tiled_range.h)


@dataclass(frozen=True)
class TiledRange1:
    bounds: Tuple[int, ...]  # tile boundaries, len = ntiles + 1

    @staticmethod
    def uniform(extent: int, tile: int, lo: int = 0) -> "TiledRange1":
        """TiledRange1::make_uniform (tiled_range1.h:289-313): ceil(extent/tile) tiles, the first
        (extent + ntiles - 1) % ntiles + 1 of them one element larger than the rest."""
        ntiles = (extent + tile - 1) // tile
        quot, rem = divmod(extent + ntiles - 1, ntiles)
        avg, nplus = quot - 1, rem + 1
        b, e = [], lo
        for i in range(ntiles):
            b.append(e)
            e += avg + 1 if i < nplus else avg
        b.append(lo + extent)
        return TiledRange1(tuple(b))

    @property
    def ntiles(self) -> int:
        return len(self.bounds) - 1

    @property
    def extents(self) -> List[int]:
        return [self.bounds[i + 1] - self.bounds[i] for i in range(self.ntiles)]

    @property
    def extent(self) -> int:
        return self.bounds[-1] - self.bounds[0]


@dataclass(frozen=True)
class TiledRange:
    dims: Tuple[TiledRange1, ...]

    @property
    def rank(self) -> int:
        return len(self.dims)

    @property
    def tiles_shape(self) -> Tuple[int, ...]:
        return tuple(d.ntiles for d in self.dims)

    def tile_extent(self, tidx: Sequence[int]) -> Tuple[int, ...]:
        return tuple(d.extents[t] for d, t in zip(self.dims, tidx))

    def tile_slices(self, tidx: Sequence[int]):
        return tuple(slice(d.bounds[t] - d.bounds[0], d.bounds[t + 1] - d.bounds[0]) for d, t in zip(self.dims, tidx))

    @property
    def elements_shape(self) -> Tuple[int, ...]:
        return tuple(d.extent for d in self.dims)


# --------------------------------------------------------------------------------------------
# GemmHelper (src/TiledArray/math/gemm_helper.h:41-278)

NoTrans, Trans = 0, 1


@dataclass
class GemmHelper:
    left_op: int
    right_op: int
    result_rank: int
    left_rank: int
    right_rank: int
    left_inner: Tuple[int, int] = field(init=False)
    left_outer: Tuple[int, int] = field(init=False)
    right_inner: Tuple[int, int] = field(init=False)
    right_outer: Tuple[int, int] = field(init=False)

    def __post_init__(self):
        assert (self.left_rank + self.right_rank - self.result_rank) % 2 == 0  # gemm_helper.h:68
        c = self.num_contract_ranks
        if self.left_op == NoTrans:  # :75-83
            self.left_outer, self.left_inner = (0, self.left_rank - c), (self.left_rank - c, self.left_rank)
        else:
            self.left_inner, self.left_outer = (0, c), (c, self.left_rank)
        if self.right_op == NoTrans:  # :86-94
            self.right_inner, self.right_outer = (0, c), (c, self.right_rank)
        else:
            self.right_outer, self.right_inner = (0, self.right_rank - c), (self.right_rank - c, self.right_rank)

    @property
    def num_contract_ranks(self) -> int:
        return (self.left_rank + self.right_rank - self.result_rank) >> 1

    def compute_matrix_sizes(self, left_extent: Sequence[int], right_extent: Sequence[int]) -> Tuple[int, int, int]:
        """gemm_helper.h:255-274."""
        assert len(left_extent) == self.left_rank and len(right_extent) == self.right_rank
        m = int(np.prod(left_extent[self.left_outer[0]:self.left_outer[1]], dtype=np.int64))
        k = int(np.prod(left_extent[self.left_inner[0]:self.left_inner[1]], dtype=np.int64))
        n = int(np.prod(right_extent[self.right_outer[0]:self.right_outer[1]], dtype=np.int64))
        return m, n, k

    def make_result_extent(self, left_extent: Sequence[int], right_extent: Sequence[int]) -> Tuple[int, ...]:
        """gemm_helper.h:166-192."""
        return tuple(left_extent[self.left_outer[0]:self.left_outer[1]]) + tuple(
            right_extent[self.right_outer[0]:self.right_outer[1]])

    def left_right_congruent(self, left_extent, right_extent) -> bool:
        return list(left_extent[self.left_inner[0]:self.left_inner[1]]) == list(
            right_extent[self.right_inner[0]:self.right_inner[1]])


def tile_gemm(left: np.ndarray, right: np.ndarray, factor: float, helper: GemmHelper,
              result: Optional[np.ndarray] = None) -> np.ndarray:
    """Tensor::gemm (tensor/tensor.h:3132-3219) -> detail::gemm (tensor/kernels.h:92-231) ->
    math::blas::gemm (math/blas.h:171-177): C = factor*op(A)*op(B) (+ C), row-major, natural lds.
    beta = 0 when ``result`` is None (tensor.h:3135-3140) else 1."""
    assert helper.left_right_congruent(left.shape, right.shape)
    m, n, k = helper.compute_matrix_sizes(left.shape, right.shape)
    A = left.reshape((m, k) if helper.left_op == NoTrans else (k, m))
    B = right.reshape((k, n) if helper.right_op == NoTrans else (n, k))
    opA = A if helper.left_op == NoTrans else A.T
    opB = B if helper.right_op == NoTrans else B.T
    prod = factor * (opA @ opB)
    ext = helper.make_result_extent(left.shape, right.shape)
    if result is None:
        return prod.reshape(ext)
    assert tuple(result.shape) == tuple(ext)
    result += prod.reshape(ext)
    return result


def tile_gemm_loops(opA: int, opB: int, m: int, n: int, k: int, alpha, A, B, beta, C):
    """Naive triple loop with the row-major conventions of tests/math_blas.cpp:156-250 (small cases)."""
    A = np.asarray(A).reshape((m, k) if opA == NoTrans else (k, m))
    B = np.asarray(B).reshape((k, n) if opB == NoTrans else (n, k))
    out = np.array(C, dtype=np.result_type(A, B, C)).reshape(m, n) * beta
    for i in range(m):
        for j in range(n):
            s = 0
            for x in range(k):
                a = A[i, x] if opA == NoTrans else A[x, i]
                b = B[x, j] if opB == NoTrans else B[j, x]
                s += a * b
            out[i, j] += alpha * s
    return out


# --------------------------------------------------------------------------------------------
# SparseShape<float> (src/TiledArray/sparse_shape.h)

f32 = np.float32


def _recursive_outer_product(size_vectors: Sequence[np.ndarray], inverse: bool) -> np.ndarray:
    """recursive_outer_product (sparse_shape.h:105-132); op = identity or 1/size (:184-189)."""
    dim = len(size_vectors)
    if dim == 1:
        v = np.asarray(size_vectors[0], dtype=f32)
        return (f32(1) / v).astype(f32) if inverse else v.copy()
    middle = (dim >> 1) + (dim & 1)
    left = _recursive_outer_product(size_vectors[:middle], inverse)
    right = _recursive_outer_product(size_vectors[middle:], inverse)
    return np.multiply.outer(left, right).astype(f32).ravel()


def shape_scale_factors(size_vectors: Sequence[np.ndarray]) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """The (left, right) vectors of scale_tile_norms<InverseVolume> (sparse_shape.h:191-201);
    rank 1 returns (sizes, None) — the divide branch (:158-173)."""
    dim = len(size_vectors)
    if dim == 1:
        return np.asarray(size_vectors[0], dtype=f32), None
    middle = (dim >> 1) + (dim & 1)
    return (_recursive_outer_product(size_vectors[:middle], True),
            _recursive_outer_product(size_vectors[middle:], True))


def shape_scale_norms(tile_norms: np.ndarray, size_vectors: Sequence[np.ndarray],
                      threshold: float = FLT_EPSILON) -> Tuple[np.ndarray, int]:
    """scale_tile_norms<ScaleBy::InverseVolume, Screen=true> (sparse_shape.h:149-217)."""
    norms = np.array(tile_norms, dtype=f32)
    thr = f32(threshold)
    left, right = shape_scale_factors(size_vectors)
    if right is None:
        out = (norms.ravel() / left).astype(f32)  # norm /= size
    else:
        xy = np.multiply.outer(left, right).astype(f32).ravel()  # x * y
        out = (norms.ravel() * xy).astype(f32)  # norm *= x*y
    zero = out < thr
    out[zero] = f32(0)
    return out.reshape(norms.shape), int(zero.sum())


@dataclass
class SparseShape:
    """SparseShape<float> restated (sparse_shape.h:77). ``norms`` are the *scaled* norms."""
    norms: np.ndarray  # float32, shape = tiles range
    size_vectors: List[np.ndarray]  # float32 tile extents per mode
    zero_tile_count: int
    threshold: float = FLT_EPSILON

    @staticmethod
    def from_tile_norms(tile_norms: np.ndarray, trange: TiledRange, threshold: float = FLT_EPSILON,
                        do_not_scale: bool = False) -> "SparseShape":
        """SparseShape(Tensor<float> tile_norms, trange, do_not_scale) (sparse_shape.h:335-349)."""
        sv = [np.asarray(d.extents, dtype=f32) for d in trange.dims]
        assert tuple(tile_norms.shape) == trange.tiles_shape
        if not do_not_scale:
            norms, nz = shape_scale_norms(tile_norms, sv, threshold)
        else:
            norms = np.array(tile_norms, dtype=f32)
            z = norms < f32(threshold)  # screen_out_zero_tiles (:261-270)
            norms[z] = 0
            nz = int(z.sum())
        return SparseShape(norms, sv, nz, threshold)

    def is_zero(self, ordinal_or_index) -> bool:
        """is_zero (sparse_shape.h:495-498): norm < my_threshold."""
        v = self.norms.ravel()[ordinal_or_index] if np.isscalar(ordinal_or_index) else self.norms[tuple(ordinal_or_index)]
        return bool(v < f32(self.threshold))

    def sparsity(self) -> float:
        """sparsity (sparse_shape.h:533-538)."""
        return self.zero_tile_count / float(self.norms.size)

    def perm(self, perm: Sequence[int]) -> "SparseShape":
        """perm (sparse_shape.h:1222-1225): permuted norms + permuted size vectors (:241-257)."""
        return SparseShape(tile_permute(self.norms, perm), permute_array(perm, self.size_vectors),
                           self.zero_tile_count, self.threshold)

    def mask(self, mask_shape: "SparseShape") -> "SparseShape":
        """mask (sparse_shape.h:653-676)."""
        assert self.norms.shape == mask_shape.norms.shape
        hit = (self.norms >= f32(self.threshold)) & (mask_shape.norms < f32(mask_shape.threshold))
        out = self.norms.copy()
        out[hit] = 0
        return SparseShape(out, self.size_vectors, self.zero_tile_count + int(hit.sum()), self.threshold)

    def scale(self, factor: float) -> "SparseShape":
        """scale (sparse_shape.h:1243-1262): value *= |factor|; hard zero below my_threshold."""
        out = (self.norms * f32(abs(factor))).astype(f32)
        z = out < f32(self.threshold)
        out[z] = 0
        return SparseShape(out, self.size_vectors, int(z.sum()), self.threshold)

    def add(self, other: "SparseShape", factor: Optional[float] = None) -> "SparseShape":
        """add / subt (sparse_shape.h:1309-1331, 1370-1391, 1499): left += right [; left *= |factor|];
        hard zero below the global threshold."""
        assert self.norms.shape == other.norms.shape
        out = (self.norms + other.norms).astype(f32)
        if factor is not None:
            out = (out * f32(abs(factor))).astype(f32)
        z = out < f32(self.threshold)
        out[z] = 0
        return SparseShape(out, self.size_vectors, int(z.sum()), self.threshold)

    def mult(self, other: "SparseShape", factor: Optional[float] = None) -> "SparseShape":
        """mult (sparse_shape.h:1522-1563): tile_norms.mult(other[, |factor|]) then
        scale_tile_norms<ScaleBy::Volume> (:149-217: rank 1 norm *= size; else norm *= x*y with the
        non-inverted recursive outer products), screened."""
        assert self.norms.shape == other.norms.shape
        out = (self.norms * other.norms).astype(f32)
        if factor is not None:
            out = (out * f32(abs(factor))).astype(f32)
        dim = len(self.size_vectors)
        if dim == 1:
            out = (out.ravel() * np.asarray(self.size_vectors[0], dtype=f32)).astype(f32)
        else:
            middle = (dim >> 1) + (dim & 1)
            xy = np.multiply.outer(_recursive_outer_product(self.size_vectors[:middle], False),
                                   _recursive_outer_product(self.size_vectors[middle:], False)).astype(f32).ravel()
            out = (out.ravel() * xy).astype(f32)
        z = out < f32(self.threshold)
        out[z] = 0
        return SparseShape(out.reshape(self.norms.shape), self.size_vectors, int(z.sum()), self.threshold)

    def gemm(self, other: "SparseShape", factor: float, helper: GemmHelper,
             perm: Optional[Sequence[int]] = None, threshold: Optional[float] = None) -> "SparseShape":
        """gemm (sparse_shape.h:1589-1681) [+ .perm(perm), :1687-1691].

        Fixed arithmetic order (the oracle's definition of bit-exactness): every fp32
        op individually rounded, k-sum sequential in k:
            la = a[m,k]*ksz[k]; rb = b[k,n]*ksz[k]; acc += la*rb; out = |factor|*acc.
        (The reference hands the k-sum to vendor SGEMM, :1654, whose internal order is not
        specified; its tests pin the result only to 1e-4 % — tests/sparse_shape.cpp:1478-1541.)"""
        thr = f32(self.threshold if threshold is None else threshold)
        abs_factor = f32(abs(factor))
        M, N, K = helper.compute_matrix_sizes(self.norms.shape, other.norms.shape)
        res_sv = list(self.size_vectors[helper.left_outer[0]:helper.left_outer[1]]) + list(
            other.size_vectors[helper.right_outer[0]:helper.right_outer[1]])
        res_ext = helper.make_result_extent(self.norms.shape, other.norms.shape)
        k_rank = helper.left_inner[1] - helper.left_inner[0]
        if k_rank > 0:
            ksz = _recursive_outer_product(self.size_vectors[helper.left_inner[0]:helper.left_inner[1]], False)
            a = self.norms.reshape((M, K) if helper.left_op == NoTrans else (K, M))
            b = other.norms.reshape((K, N) if helper.right_op == NoTrans else (N, K))
            a = a if helper.left_op == NoTrans else a.T
            b = b if helper.right_op == NoTrans else b.T
            out = shape_gemm_kernel(np.ascontiguousarray(a), np.ascontiguousarray(b), ksz, abs_factor)
        else:
            out = (np.multiply.outer(self.norms.ravel(), other.norms.ravel()).astype(f32) * abs_factor).astype(f32)
        zero = out < thr
        out[zero] = 0
        res = SparseShape(out.reshape(res_ext), res_sv, int(zero.sum()), float(thr))
        return res.perm(perm) if perm is not None else res


def shape_gemm_kernel(a: np.ndarray, b: np.ndarray, ksz: np.ndarray, abs_factor) -> np.ndarray:
    """The fp32 arithmetic of SparseShape::gemm in the oracle's fixed order (see SparseShape.gemm)."""
    a, b, ksz = a.astype(f32), b.astype(f32), ksz.astype(f32)
    la = (a * ksz[None, :]).astype(f32)  # :1637-1644
    rb = (b * ksz[:, None]).astype(f32)  # :1646-1652
    acc = np.zeros((a.shape[0], b.shape[1]), dtype=f32)
    for k in range(a.shape[1]):
        acc = (acc + np.multiply.outer(la[:, k], rb[k, :]).astype(f32)).astype(f32)
    return (f32(abs_factor) * acc).astype(f32)


# --------------------------------------------------------------------------------------------
# GEMMPermutationOptimizer (src/TiledArray/expressions/permopt.h:254-376)

PT_IDENTITY, PT_TRANSPOSE, PT_GENERAL = 1, 2, 3


def split_indices(s: str) -> List[str]:
    return [x.strip() for x in s.split(",")] if s.strip() else []


def gemm_permutation_optimizer(left: Sequence[str], right: Sequence[str], prefer_to_permute_left: bool = True):
    """compute_index_list_contraction (permopt.h:254-376).
    Returns (target_left, target_right, target_result, left_permtype, right_permtype)."""
    left, right = list(left), list(right)
    left_rank, right_rank = len(left), len(right)
    res_left: List[str] = []
    res_right: List[str] = []
    res: List[str] = []
    for var in left:  # :268-279
        if var not in right:
            res_left.append(var)
            res.append(var)
        else:
            res_right.append(var)
    inner_rank, left_outer_rank = len(res_right), len(res_left)
    right_outer_rank = right_rank - inner_rank
    if inner_rank == 0:  # outer product (:291-304)
        for var in right:
            res_right.append(var)
            res.append(var)
        return res_left, res_right, res, PT_GENERAL, PT_GENERAL
    ordered = l_nt = l_t = r_nt = r_t = True
    perm_left = (left_rank < right_rank) or (left_rank == right_rank and prefer_to_permute_left)  # :317-319
    for i, idx in enumerate(right):  # :324-360
        j = left.index(idx) if idx in left else left_rank
        if j == left_rank:
            res_right.append(idx)
            res.append(idx)
        else:
            x = len(res_left) - left_outer_rank
            ordered = ordered and (res_right[x] == idx)
            l_nt = l_nt and (j >= left_outer_rank)
            l_t = l_t and (j < inner_rank)
            r_nt = r_nt and (i < inner_rank)
            r_t = r_t and (i >= right_outer_rank)
            if ordered:
                res_left.append(idx)
            elif perm_left:
                res_left.append(idx)
                res_right[x] = idx
                l_nt = l_t = False
            else:
                res_left.append(res_right[x])
                r_nt = r_t = False

    def to_op(nt, t):
        return PT_IDENTITY if nt else (PT_TRANSPOSE if t else PT_GENERAL)

    return res_left, res_right, res, to_op(l_nt, l_t), to_op(r_nt, r_t)


def index_perm(frm: Sequence[str], to: Sequence[str]) -> Optional[List[int]]:
    """Image-form permutation taking index list ``frm`` to ``to``; None when identical."""
    if list(frm) == list(to):
        return None
    return [list(to).index(x) for x in frm]


@dataclass
class ContractionPlan:
    """What ContEngine decides for ``result(target) = left(l) * right(r)``
    (expressions/binary_engine.h:101-178, cont_engine.h:354-529)."""
    left_target: List[str]
    right_target: List[str]
    result_gemm: List[str]
    left_permtype: int
    right_permtype: int
    opA: int
    opB: int
    perm_left: Optional[List[int]]   # explicit argument permutation (general only)
    perm_right: Optional[List[int]]
    perm_result: Optional[List[int]]  # GEMM result order -> target order
    helper: GemmHelper


def plan_contraction(target: str, left: str, right: str) -> ContractionPlan:
    L, R, T = split_indices(left), split_indices(right), split_indices(target)
    tl, tr, res, lt, rt = gemm_permutation_optimizer(L, R, True)  # leaves equal -> prefer left
    opA = Trans if lt == PT_TRANSPOSE else NoTrans  # to_cblas_op (permopt.h:47-56)
    opB = Trans if rt == PT_TRANSPOSE else NoTrans
    helper = GemmHelper(opA, opB, len(res), len(L), len(R))
    return ContractionPlan(tl, tr, res, lt, rt, opA, opB,
                           index_perm(L, tl) if lt == PT_GENERAL else None,
                           index_perm(R, tr) if rt == PT_GENERAL else None,
                           index_perm(res, T) if T else None, helper)


# --------------------------------------------------------------------------------------------
# ProcGrid (src/TiledArray/proc_grid.h:97-260) and CyclicPmap (pmap/cyclic_pmap.h)


def _optimal_proc_row(nprocs: float, Mm: float, Nn: float) -> int:
    x = math.sqrt(nprocs)
    PMm, two_P = nprocs * Mm, nprocs + nprocs
    it = 0
    while True:
        x2 = x * x
        Nx2 = Nn * x2
        f = Nx2 * (2.0 * x2 - x) + PMm * (x - two_P)
        df = Nx2 * (8.0 * x - 3.0) + PMm
        xn = x - f / df
        r = abs(xn - x)
        x = xn
        it += 1
        if not (r > 0.1 and it < 21):
            break
    return int(x + 0.5)


def _minimize_unused_procs(x: int, y: int, nprocs: int, min_x: int, max_x: int) -> Tuple[int, int]:
    unused = x * y  # proc_grid.h:153 (as written in the reference)
    if unused == 0:
        return x, y
    delta = max(1, int(math.log2(nprocs)))
    optimal_x, diff = x, 0
    min_test_x = max(min_x, x - delta)
    test_x = min(x + delta, max_x)
    while test_x >= min_test_x:
        test_y = nprocs // test_x
        test_unused = nprocs - test_x * test_y
        test_diff = abs(optimal_x - test_x)
        if test_unused < unused or (test_unused == unused and test_diff < diff):
            x, y, unused, diff = test_x, test_y, test_unused, test_diff
        test_x -= 1
    return x, y


@dataclass
class ProcGrid:
    rows: int
    cols: int
    proc_rows: int
    proc_cols: int
    proc_size: int
    rank_row: int
    rank_col: int
    local_rows: int
    local_cols: int
    local_size: int


def proc_grid(rank: int, nprocs: int, rows: int, cols: int, row_size: int, col_size: int) -> ProcGrid:
    """ProcGrid::init (proc_grid.h:184-260)."""
    size = rows * cols
    rr = rc = -1
    lr = lc = ls = 0
    if nprocs == 1:
        pr = pc = ps = 1
        if rank < ps:
            rr = rc = 0
            lr, lc, ls = rows, cols, size
    elif size <= nprocs:
        pr, pc, ps = rows, cols, size
        if rank < ps:
            rr, rc = rank // pc, rank % pc
            lr = lc = ls = 1
    else:
        min_pr = max((nprocs + cols - 1) // cols, 1)
        max_pr = min(nprocs, rows)
        pr = max(min_pr, min(_optimal_proc_row(float(nprocs), float(row_size), float(col_size)), max_pr))
        pc = nprocs // pr
        if min_pr < pr < max_pr:
            pr, pc = _minimize_unused_procs(pr, pc, nprocs, min_pr, max_pr)
        ps = pr * pc
        if rank < ps:
            rr, rc = rank // pc, rank % pc
            lr = rows // pr + (1 if rr < rows % pr else 0)
            lc = cols // pc + (1 if rc < cols % pc else 0)
            ls = lr * lc
    return ProcGrid(rows, cols, pr, pc, ps, rr, rc, lr, lc, ls)


def cyclic_owner(tile: int, cols: int, proc_rows: int, proc_cols: int) -> int:
    """CyclicPmap::owner (pmap/cyclic_pmap.h:123-134)."""
    return ((tile // cols) % proc_rows) * proc_cols + (tile % cols) % proc_cols


# --------------------------------------------------------------------------------------------
# SUMMA (src/TiledArray/dist_eval/contraction_eval.h)


def summa_rank_schedule(Pr: int, Pc: int, r: int, c: int, Mt: int, Nt: int, Kt: int,
                        a_zero=None, b_zero=None, c_zero=None):
    """The contraction steps and tile-pair lists of grid rank (r,c).

    Steps: iterate_sparse (:974-1001) — the k where this rank's A column (rows i = r mod Pr) and
    B row (cols j = c mod Pc) both hold a non-zero tile. Pairs: contract (:1311-1384) — row-major
    double loop over col x row, skipping zero result tiles (:1370).
    ``*_zero`` are boolean arrays (True = zero tile) or None for dense.
    Returns [(k, [(i, j), ...]), ...]."""
    steps = []
    for k in range(Kt):
        col = [i for i in range(r, Mt, Pr) if a_zero is None or not a_zero[i, k]]  # get_col (:655-661)
        row = [j for j in range(c, Nt, Pc) if b_zero is None or not b_zero[k, j]]  # get_row (:667-676)
        if not col or not row:
            continue
        pairs = [(i, j) for i in col for j in row if c_zero is None or not c_zero[i, j]]
        steps.append((k, pairs))
    return steps


def summa_contract(a_tiles: Dict[Tuple[int, int], np.ndarray], b_tiles: Dict[Tuple[int, int], np.ndarray],
                   Mt: int, Nt: int, Kt: int, opA: int, opB: int, alpha: float,
                   a_zero=None, b_zero=None, c_zero=None, Pr: int = 1, Pc: int = 1):
    """Whole block-sparse contraction, executed as every rank of a Pr x Pc grid would: returns
    {(i,j): C tile} for all non-zero result tiles that received a contribution, and the total
    number of tile pairs. Tiles are 2-d (fused) matrices: A(i,k) is m x k (opA=N) or k x m."""
    out: Dict[Tuple[int, int], np.ndarray] = {}
    npairs = 0
    for r in range(Pr):
        for c in range(Pc):
            for k, pairs in summa_rank_schedule(Pr, Pc, r, c, Mt, Nt, Kt, a_zero, b_zero, c_zero):
                for (i, j) in pairs:
                    A, B = a_tiles[(i, k)], b_tiles[(k, j)]
                    prod = alpha * ((A if opA == NoTrans else A.T) @ (B if opB == NoTrans else B.T))
                    if (i, j) in out:
                        out[(i, j)] += prod  # ContractReduce accumulate (contract_reduce.h:409-453)
                    else:
                        out[(i, j)] = prod
                    npairs += 1
    return out, npairs


def rel_frobenius(x: np.ndarray, ref: np.ndarray) -> float:
    d = float(np.linalg.norm(np.asarray(x, dtype=np.float64).ravel() - np.asarray(ref, dtype=np.float64).ravel()))
    n = float(np.linalg.norm(np.asarray(ref, dtype=np.float64).ravel()))
    return d / n if n > 0 else d
