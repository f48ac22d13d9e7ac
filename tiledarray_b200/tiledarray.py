"""Host-side mirror of the TiledArray interface for the contraction path.

Names, argument meaning and error behaviour follow the reference so that tests read like
TiledArray's own (``c["m,n"] = a["m,k"] * b["k,n"]`` is ``c("m,n") = a("m,k") * b("k,n")``,
reference: src/TiledArray/dist_array.h:1311, expressions/tsr_expr.h:132, mult_expr.h:192,
expr.h:378). Only metadata lives here; every tile operation is a libtadev C-ABI call
(tadev_gemm_grouped_f64 / tadev_summa_f64 / tadev_permute / tadev_shape_*): there is no CPU
compute path in this package.

    World         one process <-> one GPU (+ NCCL communicators)          (MADWorld World analogue)
    TiledRange1/TiledRange                                                 (tiled_range1.h:47, tiled_range.h)
    SparseShape   SparseShape<float>; gemm/scale run on the device         (sparse_shape.h:77)
    DistArray     tiles in device memory, cyclic ownership on the grid     (dist_array.h:63)
    TsrExpr/MultExpr/ScalExpr + ContEngine                                  (expressions/cont_engine.h:354-677)
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import ContractionPlanC, ProcGridC, SummaPlanC, SummaStatsC, check
from .device import Device, DeviceBuffer, HostBuffer

FLT_EPSILON = float(np.finfo(np.float32).eps)
f32 = np.float32


class TiledArrayException(RuntimeError):
    """TiledArray::Exception analogue (error.h:39-83): raised where the reference TA_ASSERTs."""


def _ta_assert(cond: bool, msg: str) -> None:
    if not cond:
        raise TiledArrayException(msg)


# ---------------------------------------------------------------------------------------------
class TiledRange1:
    def __init__(self, *bounds: int):
        if len(bounds) == 1 and not isinstance(bounds[0], int):
            bounds = tuple(bounds[0])
        _ta_assert(len(bounds) >= 2 and all(b1 > b0 for b0, b1 in zip(bounds, bounds[1:])),
                   "TiledRange1: tile boundaries must be strictly increasing")
        self.bounds = tuple(int(b) for b in bounds)

    @staticmethod
    def make_uniform(extent: int, tile: int, lo: int = 0) -> "TiledRange1":
        return TiledRange1(*(list(range(lo, lo + extent, tile)) + [lo + extent]))

    @property
    def ntiles(self) -> int:
        return len(self.bounds) - 1

    def tile_extent(self, t: int) -> int:
        return self.bounds[t + 1] - self.bounds[t]

    @property
    def extents(self) -> List[int]:
        return [self.tile_extent(t) for t in range(self.ntiles)]

    @property
    def extent(self) -> int:
        return self.bounds[-1] - self.bounds[0]

    def __eq__(self, o):
        return isinstance(o, TiledRange1) and self.bounds == o.bounds

    def __hash__(self):
        return hash(self.bounds)


class TiledRange:
    def __init__(self, dims: Sequence[TiledRange1]):
        self.dims = tuple(dims)

    @property
    def rank(self) -> int:
        return len(self.dims)

    @property
    def tiles_shape(self) -> Tuple[int, ...]:
        return tuple(d.ntiles for d in self.dims)

    @property
    def ntiles(self) -> int:
        return int(np.prod(self.tiles_shape, dtype=np.int64)) if self.dims else 1

    @property
    def elements_shape(self) -> Tuple[int, ...]:
        return tuple(d.extent for d in self.dims)

    def tile_index(self, ordinal: int) -> Tuple[int, ...]:
        return tuple(int(x) for x in np.unravel_index(ordinal, self.tiles_shape)) if self.dims else ()

    def tile_ordinal(self, idx: Sequence[int]) -> int:
        return int(np.ravel_multi_index(tuple(idx), self.tiles_shape)) if self.dims else 0

    def tile_extent(self, idx: Sequence[int]) -> Tuple[int, ...]:
        return tuple(d.tile_extent(t) for d, t in zip(self.dims, idx))

    def tile_slices(self, idx: Sequence[int]):
        return tuple(slice(d.bounds[t] - d.bounds[0], d.bounds[t + 1] - d.bounds[0]) for d, t in zip(self.dims, idx))

    def __eq__(self, o):
        return isinstance(o, TiledRange) and self.dims == o.dims


# ---------------------------------------------------------------------------------------------
class World:
    """One process <-> one GPU. ``size > 1`` needs communicators (``init_comm``)."""

    def __init__(self, device: Optional[Device] = None, rank: int = 0, size: int = 1, device_index: int = 0):
        self.dev = device or Device(device_index)
        self.rank, self.size = rank, size
        self.grid: Optional[Tuple[int, int]] = None  # (Pr, Pc) once communicators exist
        self.lib = self.dev.lib

    def proc_grid(self, rows: int, cols: int, row_size: int, col_size: int, rank: Optional[int] = None) -> ProcGridC:
        g = ProcGridC()
        check(self.lib.tadev_proc_grid_make(self.rank if rank is None else rank, self.size, rows, cols, row_size,
                                            col_size, C.byref(g)))
        return g

    def init_comm(self, Pr: int, Pc: int, unique_id: Optional[bytes] = None) -> None:
        """Create world/row/column NCCL communicators for a Pr x Pc grid. The 128-byte NCCL id is
        generated on rank 0 and distributed with torch.distributed (any backend) unless given."""
        if self.size == 1:
            self.grid = (1, 1)
            return
        if unique_id is None:
            import torch
            import torch.distributed as dist
            buf = (C.c_char * 128)()
            if self.rank == 0:
                check(self.lib.tadev_comm_unique_id(buf))
            t = torch.tensor(list(bytes(buf)), dtype=torch.uint8)
            if dist.get_backend() == "nccl":
                t = t.cuda()
            dist.broadcast(t, 0)
            unique_id = bytes(t.cpu().tolist())
        idbuf = (C.c_char * 128).from_buffer_copy(unique_id)
        check(self.lib.tadev_comm_init(self.dev.ctx, idbuf, self.rank, self.size, Pr, Pc))
        self.grid = (Pr, Pc)

    @property
    def grid_pos(self) -> Tuple[int, int]:
        Pr, Pc = self.grid or (1, 1)
        if self.rank >= Pr * Pc:
            return (-1, -1)
        return (self.rank // Pc, self.rank % Pc)


# ---------------------------------------------------------------------------------------------
def _recursive_outer(size_vectors: Sequence[np.ndarray], inverse: bool) -> np.ndarray:
    dim = len(size_vectors)
    if dim == 1:
        v = np.asarray(size_vectors[0], dtype=f32)
        return (f32(1) / v).astype(f32) if inverse else v.copy()
    middle = (dim >> 1) + (dim & 1)
    left = _recursive_outer(size_vectors[:middle], inverse)
    right = _recursive_outer(size_vectors[middle:], inverse)
    return np.multiply.outer(left, right).astype(f32).ravel()


class SparseShape:
    """SparseShape<float> (sparse_shape.h:77). Norm tensors are tiny and replicated on every
    rank; the arithmetic that defines them (scaling, norm-product GEMM, thresholding, counting)
    runs in the device kernels of shape.cu so the tile lists fed to the GEMM come from the GPU."""

    _threshold = FLT_EPSILON  # static default threshold (sparse_shape.h:1941)

    @classmethod
    def threshold(cls, value: Optional[float] = None) -> float:
        if value is not None:
            cls._threshold = float(value)
        return cls._threshold

    def __init__(self, world: World, tile_norms: np.ndarray, trange: TiledRange, do_not_scale: bool = False,
                 _prebuilt=None):
        self.world = world
        if _prebuilt is not None:
            self.norms, self.size_vectors, self.zero_tile_count, self.my_threshold = _prebuilt
            return
        tile_norms = np.asarray(tile_norms, dtype=f32)
        _ta_assert(tile_norms.size > 0, "SparseShape: empty norm tensor")
        _ta_assert(tuple(tile_norms.shape) == trange.tiles_shape, "SparseShape: norm tensor does not match trange")
        self.size_vectors = [np.asarray(d.extents, dtype=f32) for d in trange.dims]
        self.my_threshold = SparseShape._threshold
        if do_not_scale:
            z = tile_norms < f32(self.my_threshold)
            self.norms = np.where(z, f32(0), tile_norms).astype(f32)
            self.zero_tile_count = int(z.sum())
        else:
            dim = len(self.size_vectors)
            if dim == 1:
                left, right = self.size_vectors[0], None
            else:
                middle = (dim >> 1) + (dim & 1)
                left = _recursive_outer(self.size_vectors[:middle], True)
                right = _recursive_outer(self.size_vectors[middle:], True)
            self.norms, self.zero_tile_count = world.dev.shape_scale(tile_norms, left, right, self.my_threshold)

    # -- queries
    def is_zero(self, ordinal: int) -> bool:
        return bool(self.norms.ravel()[ordinal] < f32(self.my_threshold))

    def is_dense(self) -> bool:
        return False

    def sparsity(self) -> float:
        return self.zero_tile_count / float(self.norms.size)

    def nnz(self) -> int:
        return self.norms.size - self.zero_tile_count

    def data(self) -> np.ndarray:
        return self.norms

    # -- algebra on the contraction path
    def perm(self, perm: Sequence[int]) -> "SparseShape":
        inv = [0] * len(perm)
        for i, p in enumerate(perm):
            inv[p] = i
        norms = np.ascontiguousarray(np.transpose(self.norms, inv))  # metadata shuffle (KBs)
        sv = [None] * len(perm)
        for i, p in enumerate(perm):
            sv[p] = self.size_vectors[i]
        return SparseShape(self.world, None, None, _prebuilt=(norms, sv, self.zero_tile_count, self.my_threshold))

    def gemm(self, other: "SparseShape", factor: float, left_op: int, right_op: int, ncontract: int,
             perm: Optional[Sequence[int]] = None) -> "SparseShape":
        """SparseShape::gemm (sparse_shape.h:1589-1691) with GemmHelper(left_op, right_op, ...)."""
        lr, rr = self.norms.ndim, other.norms.ndim
        lo = (0, lr - ncontract) if left_op == _lib.OP_N else (ncontract, lr)
        li = (lr - ncontract, lr) if left_op == _lib.OP_N else (0, ncontract)
        ro = (ncontract, rr) if right_op == _lib.OP_N else (0, rr - ncontract)
        ri = (0, ncontract) if right_op == _lib.OP_N else (rr - ncontract, rr)
        _ta_assert(list(self.norms.shape[li[0]:li[1]]) == list(other.norms.shape[ri[0]:ri[1]]),
                   "SparseShape::gemm: contracted tile ranges are not congruent")
        res_ext = tuple(self.norms.shape[lo[0]:lo[1]]) + tuple(other.norms.shape[ro[0]:ro[1]])
        res_sv = list(self.size_vectors[lo[0]:lo[1]]) + list(other.size_vectors[ro[0]:ro[1]])
        M = int(np.prod(self.norms.shape[lo[0]:lo[1]], dtype=np.int64))
        N = int(np.prod(other.norms.shape[ro[0]:ro[1]], dtype=np.int64))
        K = int(np.prod(self.norms.shape[li[0]:li[1]], dtype=np.int64))
        thr = SparseShape._threshold
        if ncontract > 0:
            ksz = _recursive_outer(self.size_vectors[li[0]:li[1]], False)
            a = self.norms.reshape((M, K) if left_op == _lib.OP_N else (K, M))
            b = other.norms.reshape((K, N) if right_op == _lib.OP_N else (N, K))
            a = a if left_op == _lib.OP_N else a.T
            b = b if right_op == _lib.OP_N else b.T
            out, nz = self.world.dev.shape_gemm(np.ascontiguousarray(a), np.ascontiguousarray(b), ksz, abs(factor), thr)
        else:
            out, nz = self.world.dev.shape_gemm(self.norms.ravel(), other.norms.ravel(), np.zeros(0, f32), abs(factor), thr)
        res = SparseShape(self.world, None, None, _prebuilt=(out.reshape(res_ext), res_sv, nz, thr))
        return res.perm(perm) if perm is not None else res

    def mask(self, mask_shape: "SparseShape") -> "SparseShape":
        _ta_assert(self.norms.shape == mask_shape.norms.shape, "SparseShape::mask: range mismatch")
        hit = (self.norms >= f32(self.my_threshold)) & (mask_shape.norms < f32(mask_shape.my_threshold))
        out = np.where(hit, f32(0), self.norms).astype(f32)
        return SparseShape(self.world, None, None,
                           _prebuilt=(out, self.size_vectors, self.zero_tile_count + int(hit.sum()), self.my_threshold))


class DenseShape:
    def is_zero(self, ordinal: int) -> bool:
        return False

    def is_dense(self) -> bool:
        return True


# ---------------------------------------------------------------------------------------------
def _split(idx: str) -> List[str]:
    out = [x.strip() for x in idx.split(",")] if idx.strip() else []
    _ta_assert(all(out) and len(set(out)) == len(out), f"bad index list '{idx}'")
    return out


class DistArray:
    """DistArray<Tensor<double>, Dense|SparsePolicy> with device-resident tiles.

    ``owner`` maps a tile ordinal to the owning rank. The default is "everything on rank 0" for a
    single-process world; arrays taking part in a multi-GPU contraction are created with
    :meth:`for_summa`, which applies the cyclic maps of the process grid (proc_grid.h:566-597).
    Local tiles live in ONE arena allocation ordered by ``arena_order`` so that a SUMMA panel is a
    contiguous byte range (broadcast in place, no packing).
    """

    def __init__(self, world: World, trange: TiledRange, shape: Optional[SparseShape] = None,
                 owner=None, arena_order: Optional[Sequence[int]] = None, memory: str = "device",
                 lazy_seed: Optional[int] = None):
        _ta_assert(memory in ("device", "host", "lazy"), "DistArray: memory must be 'device', 'host' or 'lazy'")
        _ta_assert((memory == "lazy") == (lazy_seed is not None), "DistArray: memory='lazy' needs lazy_seed (and only it)")
        self.world, self.trange = world, trange
        self.memory = memory  # "host": tiles live in pinned host memory (TiledArray's default home for
        #                       arrays) and are streamed through the GPU by the SUMMA driver.
        #                       "lazy": tiles are never stored; they are generated on the device when a
        #                       contraction needs them (TiledArray's lazy tiles, array_eval.h:42) from the
        #                       counter RNG keyed by (lazy_seed, tile ordinal) -- tensors larger than HBM.
        self.lazy_seed = lazy_seed
        self.shape = shape if shape is not None else DenseShape()
        self._owner = owner or (lambda ordinal: 0)
        self.tiles: Dict[int, DeviceBuffer] = {}
        self._arena: Optional[DeviceBuffer] = None
        self._arena_order = arena_order

    # -- structure
    def is_local(self, ordinal: int) -> bool:
        return self._owner(ordinal) == self.world.rank

    def is_zero(self, ordinal: int) -> bool:
        return self.shape.is_zero(ordinal)

    def local_nonzero_ordinals(self) -> List[int]:
        return [o for o in range(self.trange.ntiles) if self.is_local(o) and not self.is_zero(o)]

    def tile_elems(self, ordinal: int) -> int:
        return int(np.prod(self.trange.tile_extent(self.trange.tile_index(ordinal)), dtype=np.int64))

    def _allocate(self) -> None:
        if self._arena is not None or self.tiles or self.memory == "lazy":
            return
        ords = self.local_nonzero_ordinals()
        if self._arena_order is not None:
            rank_of = {o: n for n, o in enumerate(self._arena_order)}
            ords.sort(key=lambda o: rank_of.get(o, o))
        sizes = [(self.tile_elems(o) + 1) & ~1 for o in ords]  # keep every tile 16-byte aligned
        total = sum(sizes)
        if self.memory == "host":
            self._arena = HostBuffer(max(total, 2) * 8)
        else:
            self._arena = self.world.dev.alloc(max(total, 2) * 8)
        off = 0
        for o, s in zip(ords, sizes):
            self.tiles[o] = self._arena.view(off * 8, self.tile_elems(o) * 8)
            off += s

    # -- initialisation (dist_array.h:983 fill_local, :1117 init_tiles, :937 set)
    def fill(self, value: float) -> "DistArray":
        self._allocate()
        for o, buf in self.tiles.items():
            n = self.tile_elems(o)
            if self.memory == "host":
                buf.numpy(np.float64, (n,))[:] = value
            else:
                self.world.dev.upload_into(buf, np.full(n, value, dtype=np.float64))
        return self

    def fill_random(self, seed: int) -> "DistArray":
        """uniform(-1,1) generated on the device; element value depends only on (seed, tile
        ordinal, offset in tile) so any distribution of the array holds identical data."""
        _ta_assert(self.memory != "lazy", "fill_random: a lazy array is defined by its lazy_seed")
        self._allocate()
        dev = self.world.dev
        if self.memory == "host":  # generate on the device (one RNG implementation), park on the host
            big = max((self.tile_elems(o) for o in self.tiles), default=0)
            tmp = dev.alloc(max(big, 1) * 8)
            for o, buf in self.tiles.items():
                n = self.tile_elems(o)
                dev.fill_uniform(tmp, n, seed, o << 32)
                check(dev.lib.tadev_memcpy_d2h(dev.ctx, buf.ptr, tmp.ptr, n * 8, dev.stream))
            dev.sync()
            tmp.free()
            return self
        for o, buf in self.tiles.items():
            dev.fill_uniform(buf, self.tile_elems(o), seed, o << 32)
        return self

    def set(self, ordinal: int, tile: np.ndarray) -> None:
        self._allocate()
        _ta_assert(self.is_local(ordinal), "DistArray::set: tile is not local")
        _ta_assert(not self.is_zero(ordinal), "DistArray::set: tile is zero in the shape")
        ext = self.trange.tile_extent(self.trange.tile_index(ordinal))
        _ta_assert(tuple(tile.shape) == tuple(ext), f"DistArray::set: tile extent {tile.shape} != {ext}")
        if self.memory == "host":
            self.tiles[ordinal].numpy(np.float64, ext)[...] = tile
        else:
            self.world.dev.upload_into(self.tiles[ordinal], np.ascontiguousarray(tile, dtype=np.float64))

    def init_from_numpy(self, full: np.ndarray) -> "DistArray":
        _ta_assert(tuple(full.shape) == self.trange.elements_shape, "init_from_numpy: shape mismatch")
        self._allocate()
        for o in self.tiles:
            self.set(o, full[self.trange.tile_slices(self.trange.tile_index(o))])
        return self

    def find(self, ordinal: int) -> np.ndarray:
        """Local tile as a host array (dist_array.h:717)."""
        ext = self.trange.tile_extent(self.trange.tile_index(ordinal))
        if self.memory == "lazy":  # materialise this one tile
            _ta_assert(self.is_local(ordinal) and not self.is_zero(ordinal), "DistArray::find: tile is zero or not local")
            n = self.tile_elems(ordinal)
            tmp = self.world.dev.alloc(n * 8)
            self.world.dev.fill_uniform(tmp, n, self.lazy_seed, ordinal << 32)
            out = self.world.dev.download(tmp, np.float64, ext)
            tmp.free()
            return out
        _ta_assert(ordinal in self.tiles, "DistArray::find: tile is zero or not local")
        if self.memory == "host":
            return self.tiles[ordinal].numpy(np.float64, ext).copy()
        return self.world.dev.download(self.tiles[ordinal], np.float64, ext)

    def to_numpy(self) -> np.ndarray:
        """Dense host copy of the LOCAL tiles (zeros elsewhere)."""
        out = np.zeros(self.trange.elements_shape, dtype=np.float64)
        for o in self.tiles:
            out[self.trange.tile_slices(self.trange.tile_index(o))] = self.find(o)
        return out

    def release(self) -> None:
        if self._arena is not None:
            self._arena.free()
            self._arena = None
        else:
            for b in self.tiles.values():
                b.free()
        self.tiles = {}

    # -- expressions
    def __getitem__(self, idx: str) -> "TsrExpr":
        return TsrExpr(self, _split(idx))

    def __call__(self, idx: str) -> "TsrExpr":
        return self[idx]

    def __setitem__(self, idx: str, expr) -> None:
        TsrExpr(self, _split(idx)).assign(expr)


# ---------------------------------------------------------------------------------------------
class Expr:
    factor = 1.0

    def __mul__(self, other):
        if isinstance(other, (int, float)):
            return ScalExpr(self, float(other))
        return MultExpr(self, other)

    def __rmul__(self, other):
        _ta_assert(isinstance(other, (int, float)), "unsupported operand")
        return ScalExpr(self, float(other))


class TsrExpr(Expr):
    def __init__(self, array: DistArray, indices: List[str]):
        _ta_assert(len(indices) == array.trange.rank, "index list rank does not match the array")
        self.array, self.indices = array, indices

    def assign(self, expr) -> None:
        factor = 1.0
        while isinstance(expr, ScalExpr):
            factor *= expr.scalar
            expr = expr.arg
        _ta_assert(isinstance(expr, MultExpr), "only contraction expressions are implemented (SURVEY §8)")
        _ta_assert(isinstance(expr.left, TsrExpr) and isinstance(expr.right, TsrExpr),
                   "contraction operands must be arrays (nested expressions are out of scope)")
        ContEngine(self, expr.left, expr.right, factor).eval()


class ScalExpr(Expr):
    def __init__(self, arg, scalar: float):
        self.arg, self.scalar = arg, scalar


class MultExpr(Expr):
    def __init__(self, left, right):
        self.left, self.right = left, right


# ---------------------------------------------------------------------------------------------
@dataclass
class ContractionStats:
    nsteps: int = 0
    nsteps_skipped: int = 0
    npairs: int = 0
    nlaunches: int = 0
    flops: float = 0.0
    bcast_bytes: int = 0
    device_ms: float = 0.0
    permute_ms: float = 0.0
    h2d_bytes: int = 0
    d2h_bytes: int = 0
    row_blocks: int = 1
    lazy_tiles: int = 0


class _OperandView:
    def __init__(self):
        self.trange = self.shape = None
        self.ords = np.zeros(0, dtype=np.int64)    # ordinals in the (permuted) tile grid
        self.vals = np.zeros(0, dtype=np.uint64)   # device/host pointers, or provider tokens
        self.provider = None                        # address of a tadev_tile_provider (lazy operand)
        self.user = None                            # its user struct (kept alive here)
        self.keep: list = []
        self.tmp: list = []                         # device buffers to free after the contraction


class ContEngine:
    """ContEngine (expressions/cont_engine.h:354-677) + Summa hand-off (:662-677).

    init_indices/perm_indices -> tadev_plan_contraction; init_struct -> result trange + shape;
    init_distribution -> ProcGrid + cyclic maps; make_dist_eval/eval -> tadev_summa_f64.
    """

    last_stats: Optional[ContractionStats] = None
    depth = 0             # SUMMA pipeline depth in windows (0 = default 2; TA_SUMMA_MAX_DEPTH analogue)
    steps_per_launch = 0  # K steps fused into one grouped-GEMM launch (0 = auto)
    row_blocks = 0        # result row blocks for host-resident results (0 = auto)
    exchange_operands = True  # evaluate C^T = B^T A^T when that needs fewer explicit tile permutations
    stream_permutes = "auto"  # True/False/"auto": permute argument tiles just in time per SUMMA window
    stream_permute_bytes = 8 << 30  # "auto": stream when the permuted copy would exceed this many bytes

    def __init__(self, result: TsrExpr, left: TsrExpr, right: TsrExpr, factor: float):
        self.result, self.left, self.right, self.factor = result, left, right, factor
        self.world = left.array.world
        self.dev = self.world.dev
        plan = ContractionPlanC()
        tgt, li, ri = (",".join(x.indices).encode() for x in (result, left, right))
        if ContEngine.exchange_operands:
            swapped = C.c_int32(0)
            check(self.world.lib.tadev_plan_contraction_opt(tgt, li, ri, C.byref(plan), C.byref(swapped)))
            if swapped.value:
                self.left, self.right = right, left
        else:
            check(self.world.lib.tadev_plan_contraction(tgt, li, ri, C.byref(plan)))
        self.plan = plan

    @staticmethod
    def _perm(arr, rank: int) -> Optional[List[int]]:
        return None if arr[0] < 0 else [int(x) for x in arr[:rank]]

    def _operand_view(self, arr: DistArray, perm: Optional[List[int]]) -> "_OperandView":
        """The operand as the SUMMA driver sees it: permuted trange/shape plus, per local non-zero tile,
        its ordinal in the permuted tile grid and either a pointer (device / pinned host) or a provider
        token. Explicit argument permutations (ArrayEvalImpl/LazyArrayTile + UnaryWrapper<Noop>,
        dist_eval/array_eval.h:42,170) are either materialised up front into one arena (one batched launch
        per distinct tile extent) or, for big operands, performed just in time per SUMMA window by the
        permute provider (no permuted copy in HBM)."""
        v = _OperandView()
        dev = self.dev
        if arr.memory == "lazy":
            _ta_assert(perm is None, "a lazy operand that needs an explicit permutation is not supported")
            ords = np.asarray(arr.local_nonzero_ordinals(), dtype=np.int64)
            v.trange, v.shape, v.ords, v.vals = arr.trange, arr.shape, ords, (ords + 1).astype(np.uint64)
            v.user = _lib.UniformSourceC(dev.ctx.value, arr.lazy_seed)
            v.provider = C.cast(dev.lib.tadev_provider_uniform, C.c_void_p).value
            return v
        ords = np.fromiter(arr.tiles.keys(), dtype=np.int64, count=len(arr.tiles))
        ptrs = np.fromiter((b_.ptr for b_ in arr.tiles.values()), dtype=np.uint64, count=len(arr.tiles))
        if perm is None:
            v.trange, v.shape, v.ords, v.vals = arr.trange, arr.shape, ords, ptrs
            return v
        rank = len(perm)
        dims = [None] * rank
        for i, p in enumerate(perm):
            dims[p] = arr.trange.dims[i]
        v.trange = TiledRange(dims)
        v.shape = arr.shape.perm(perm) if not arr.shape.is_dense() else arr.shape
        # vectorised index arithmetic: tile index, extent and permuted ordinal of every local tile
        tshape = arr.trange.tiles_shape
        idx = np.stack(np.unravel_index(ords, tshape), axis=1) if len(ords) else np.zeros((0, rank), dtype=np.int64)
        ext = np.stack([np.asarray(d.extents, dtype=np.int64)[idx[:, a]] for a, d in enumerate(arr.trange.dims)], axis=1) \
            if len(ords) else np.zeros((0, rank), dtype=np.int64)
        pidx = np.empty_like(idx)
        pidx[:, perm] = idx
        v.ords = np.ravel_multi_index(tuple(pidx.T), v.trange.tiles_shape).astype(np.int64) if len(ords) else ords
        nbytes = int(ext.prod(axis=1).sum()) * 8 if len(ords) else 0
        stream = ContEngine.stream_permutes
        if stream == "auto":
            stream = nbytes > ContEngine.stream_permute_bytes
        if stream:
            v.vals = np.arange(1, len(ords) + 1, dtype=np.uint64)
            v.keep = [np.ascontiguousarray(ext), np.ascontiguousarray(ptrs)]
            src = _lib.PermuteSourceC()
            src.ctx, src.rank = dev.ctx.value, rank
            for i, p in enumerate(perm):
                src.perm[i] = p
            src.extents = v.keep[0].ctypes.data_as(C.POINTER(C.c_int64))
            src.src = v.keep[1].ctypes.data_as(C.POINTER(C.c_void_p))
            v.user = src
            v.provider = C.cast(dev.lib.tadev_provider_permute, C.c_void_p).value
            return v
        elems = ext.prod(axis=1) if len(ords) else np.zeros(0, dtype=np.int64)
        padded = (elems + 1) & ~1
        offs = np.concatenate([[0], np.cumsum(padded)]).astype(np.int64)
        arena = dev.alloc(max(int(offs[-1]), 2) * 8)
        v.vals = (arena.ptr + offs[:-1] * 8).astype(np.uint64)
        v.tmp = [arena]
        # one batched launch per distinct tile extent
        if len(ords):
            uniq, inv = np.unique(ext, axis=0, return_inverse=True)
            inv = inv.ravel()
            for u in range(len(uniq)):
                sel = np.nonzero(inv == u)[0]
                dev.permute_batched_ptrs(tuple(int(x) for x in uniq[u]), perm, 8, ptrs[sel], v.vals[sel])
        return v

    def eval(self) -> ContractionStats:
        P, w, dev = self.plan, self.world, self.dev
        A, B, Cres = self.left.array, self.right.array, self.result.array
        A._allocate()
        B._allocate()
        old_host_arena = None
        if Cres is not A and Cres is not B:
            if Cres.memory == "host" and isinstance(Cres._arena, HostBuffer):
                old_host_arena, Cres._arena = Cres._arena, None  # page-locking is slow: recycle the pinned arena
            Cres.release()  # hand the old result's memory back to the pool BEFORE allocating the new one
        nc = P.inner_rank
        stats = ContractionStats()
        _ta_assert(A.memory != "host" or P.perm_left[0] < 0, "host-resident left operand needs an explicit permutation: not supported")
        _ta_assert(B.memory != "host" or P.perm_right[0] < 0, "host-resident right operand needs an explicit permutation: not supported")
        _ta_assert(Cres.memory == "device" or P.perm_result[0] < 0, "host-resident result needs a result permutation: not supported")
        _ta_assert(Cres.memory != "lazy", "the result of a contraction cannot be a lazy array")
        with dev.timer() as tperm:
            vA = self._operand_view(A, self._perm(P.perm_left, P.left_rank))
            vB = self._operand_view(B, self._perm(P.perm_right, P.right_rank))
        trA, shA, trB, shB = vA.trange, vA.shape, vB.trange, vB.shape
        tmpA, tmpB = vA.tmp, vB.tmp
        stats.permute_ms = tperm.ms if (tmpA or tmpB) else 0.0

        # fused (matrix) views: outer/inner mode ranges of each operand (GemmHelper, gemm_helper.h:62-98)
        lr, rr = trA.rank, trB.rank
        lo = (0, lr - nc) if P.opA == _lib.OP_N else (nc, lr)
        li = (lr - nc, lr) if P.opA == _lib.OP_N else (0, nc)
        ro = (nc, rr) if P.opB == _lib.OP_N else (0, rr - nc)
        ri = (0, nc) if P.opB == _lib.OP_N else (rr - nc, rr)
        for d in range(nc):  # left_right_congruent
            _ta_assert(trA.dims[li[0] + d] == trB.dims[ri[0] + d], "contraction: inner tiled ranges are not congruent")
        res_dims = list(trA.dims[lo[0]:lo[1]]) + list(trB.dims[ro[0]:ro[1]])
        tr_gemm = TiledRange(res_dims)  # result in GEMM order (make_trange, cont_engine.h:593-637)

        def fused_ext(dims):
            ext = np.ones(1, dtype=np.int64)
            for d in dims:
                ext = np.multiply.outer(ext, np.asarray(d.extents, dtype=np.int64)).ravel()
            return ext

        m_ext, n_ext, k_ext = fused_ext(trA.dims[lo[0]:lo[1]]), fused_ext(trB.dims[ro[0]:ro[1]]), fused_ext(trA.dims[li[0]:li[1]])
        Mt, Nt, Kt = len(m_ext), len(n_ext), len(k_ext)

        # result shape (make_shape, cont_engine.h:642-660)
        perm_res = self._perm(P.perm_result, P.result_rank)
        sparse = not (shA.is_dense() and shB.is_dense())
        if sparse:
            _ta_assert(not shA.is_dense() and not shB.is_dense(), "mixed dense/sparse contraction is not supported")
            sh_gemm = shA.gemm(shB, self.factor, P.opA, P.opB, nc)
        else:
            sh_gemm = DenseShape()

        # target structure
        if perm_res is None:
            tr_target = tr_gemm
        else:
            dims = [None] * len(perm_res)
            for i, p in enumerate(perm_res):
                dims[p] = tr_gemm.dims[i]
            tr_target = TiledRange(dims)
        _ta_assert(Cres.trange == tr_target or not Cres.tiles and Cres.trange.rank == tr_target.rank,
                   "result array tiling does not match the expression")

        # distribution (init_distribution, cont_engine.h:537-587)
        Pr, Pc = w.grid or (1, 1)
        r, c = w.grid_pos
        if w.size > 1:
            g = w.proc_grid(Mt, Nt, int(m_ext.sum()), int(n_ext.sum()))
            _ta_assert((g.proc_rows, g.proc_cols) == (Pr, Pc),
                       f"communicators were built for a {Pr}x{Pc} grid but ProcGrid chooses {g.proc_rows}x{g.proc_cols}")

        def norms_2d(sh, rows, cols, transposed):
            if sh.is_dense():
                return None
            n2 = sh.norms.reshape((cols, rows) if transposed else (rows, cols))
            return np.ascontiguousarray(n2.T if transposed else n2, dtype=f32)

        a_n = norms_2d(shA, Mt, Kt, P.opA == _lib.OP_T)
        b_n = norms_2d(shB, Kt, Nt, P.opB == _lib.OP_T)
        c_n = None if sh_gemm.is_dense() else np.ascontiguousarray(sh_gemm.norms.reshape(Mt, Nt), dtype=f32)
        thr = SparseShape._threshold

        # tile tables in fused (row-major) ordinals
        a_tab = np.zeros(max(Mt * Kt, 1), dtype=np.uint64)
        b_tab = np.zeros(max(Kt * Nt, 1), dtype=np.uint64)
        c_tab = np.zeros(max(Mt * Nt, 1), dtype=np.uint64)
        if len(vA.ords):
            oa = vA.ords
            a_tab[oa if P.opA == _lib.OP_N else (oa % Mt) * Kt + oa // Mt] = vA.vals
        if len(vB.ords):
            ob = vB.ords
            b_tab[ob if P.opB == _lib.OP_N else (ob % Kt) * Nt + ob // Kt] = vB.vals
        # result tiles in GEMM order, owned cyclically by (i % Pr, j % Pc)
        # (vectorised: block-sparse results have 10^4 tiles and this is on the timed path)
        if r >= 0:
            I, J = np.arange(r, Mt, Pr), np.arange(c, Nt, Pc)
            keep = np.ones((len(I), len(J)), dtype=bool) if c_n is None else (c_n[np.ix_(I, J)] >= f32(thr))
            ii, jj = np.nonzero(keep)
            ci, cj = I[ii], J[jj]
        else:
            ci = cj = np.zeros(0, dtype=np.int64)
        elems = (m_ext[ci] * n_ext[cj]).astype(np.int64)
        padded = (elems + 1) & ~1
        offs = np.concatenate([[0], np.cumsum(padded)]).astype(np.int64)
        c_on_host = Cres.memory == "host"
        need_bytes = max(int(offs[-1]), 2) * 8
        if c_on_host:
            if old_host_arena is not None and old_host_arena.nbytes >= need_bytes:
                arena = old_host_arena
            else:
                if old_host_arena is not None:
                    old_host_arena.free()
                arena = HostBuffer(need_bytes)
        else:
            arena = dev.alloc(need_bytes)
        keys = (ci * Nt + cj).astype(np.int64)
        c_tab[keys] = (arena.ptr + offs[:-1] * 8).astype(np.uint64)
        local_c = list(zip(ci.tolist(), cj.tolist()))
        sizes = padded.tolist()
        if c_on_host:
            gemm_tiles = {k_: HostBuffer(e_ * 8, arena.ptr + o_ * 8, owned=False)
                          for k_, o_, e_ in zip(keys.tolist(), offs[:-1].tolist(), elems.tolist())}
        else:
            gemm_tiles = {k_: DeviceBuffer(dev, arena.ptr + o_ * 8, e_ * 8, False)
                          for k_, o_, e_ in zip(keys.tolist(), offs[:-1].tolist(), elems.tolist())}

        sp = SummaPlanC()
        sp.Mt, sp.Nt, sp.Kt = Mt, Nt, Kt
        sp.m_ext = m_ext.ctypes.data_as(C.POINTER(C.c_int64))
        sp.n_ext = n_ext.ctypes.data_as(C.POINTER(C.c_int64))
        sp.k_ext = k_ext.ctypes.data_as(C.POINTER(C.c_int64))
        sp.opA, sp.opB, sp.alpha = P.opA, P.opB, self.factor
        fp = C.POINTER(C.c_float)
        sp.a_norms = a_n.ctypes.data_as(fp) if a_n is not None else None
        sp.b_norms = b_n.ctypes.data_as(fp) if b_n is not None else None
        sp.c_norms = c_n.ctypes.data_as(fp) if c_n is not None else None
        sp.threshold = thr
        vpp = C.POINTER(C.c_void_p)
        sp.a_tiles, sp.b_tiles, sp.c_tiles = (a_tab.ctypes.data_as(vpp), b_tab.ctypes.data_as(vpp),
                                              c_tab.ctypes.data_as(vpp))
        sp.accumulate, sp.depth, sp.steps_per_launch = 0, ContEngine.depth, ContEngine.steps_per_launch
        sp.flags = ((_lib.SUMMA_A_ON_HOST if A.memory == "host" else 0) | (_lib.SUMMA_B_ON_HOST if B.memory == "host" else 0) |
                    (_lib.SUMMA_C_ON_HOST if c_on_host else 0) | (_lib.SUMMA_A_LAZY if vA.provider else 0) |
                    (_lib.SUMMA_B_LAZY if vB.provider else 0))
        if vA.provider:
            sp.a_provider, sp.a_user = vA.provider, C.cast(C.pointer(vA.user), C.c_void_p)
        if vB.provider:
            sp.b_provider, sp.b_user = vB.provider, C.cast(C.pointer(vB.user), C.c_void_p)
        sp.row_blocks = ContEngine.row_blocks
        st = SummaStatsC()
        check(w.lib.tadev_summa_f64(dev.ctx, C.byref(sp), C.byref(st)))
        stats.nsteps, stats.nsteps_skipped, stats.npairs = st.nsteps, st.nsteps_skipped, st.npairs
        stats.nlaunches, stats.flops, stats.bcast_bytes, stats.device_ms = st.nlaunches, st.flops, st.bcast_bytes, st.device_ms
        stats.h2d_bytes, stats.d2h_bytes, stats.row_blocks = st.h2d_bytes, st.d2h_bytes, st.row_blocks
        stats.lazy_tiles = st.lazy_tiles
        for b in tmpA + tmpB:
            b.free()

        # hand the result tiles to the target array (finalize, contraction_eval.h:1180-1269; the
        # result permutation is ContractReduce's post-process, contract_reduce.h:370-378)
        Cres.release()
        Cres.trange = tr_target
        if sh_gemm.is_dense():
            Cres.shape = DenseShape()
        else:
            Cres.shape = sh_gemm.perm(perm_res) if perm_res is not None else sh_gemm
        owner_map: Dict[int, int] = {}
        if perm_res is None:
            Cres._arena = arena
            Cres.tiles = gemm_tiles
        else:
            with dev.timer() as tp2:
                out_arena = dev.alloc(arena.nbytes)
                off = 0
                by_extent: Dict[Tuple[int, ...], Tuple[list, list]] = {}
                for (i, j), s in zip(local_c, sizes):
                    o = i * Nt + j
                    idx = tr_gemm.tile_index(o)
                    ext = tr_gemm.tile_extent(idx)
                    pidx = [0] * len(perm_res)
                    for a_, p in enumerate(perm_res):
                        pidx[p] = idx[a_]
                    dst = out_arena.view(off * 8, gemm_tiles[o].nbytes)
                    srcs, dsts = by_extent.setdefault(tuple(ext), ([], []))
                    srcs.append(gemm_tiles[o])
                    dsts.append(dst)
                    Cres.tiles[tr_target.tile_ordinal(pidx)] = dst
                    off += s
                for ext, (srcs, dsts) in by_extent.items():
                    dev.permute_batched(ext, perm_res, 8, srcs, dsts)
            stats.permute_ms += tp2.ms
            arena.free()
            Cres._arena = out_arena
        # ownership of the target ordinals follows the GEMM-order cyclic map
        gemm_shape = tr_gemm.tiles_shape

        def owner(ordinal: int, _perm=perm_res, _tr=tr_target, _Pr=Pr, _Pc=Pc, _Nt=Nt, _gs=gemm_shape) -> int:
            idx = _tr.tile_index(ordinal)
            if _perm is not None:
                gidx = [idx[p] for p in _perm]
            else:
                gidx = list(idx)
            go = int(np.ravel_multi_index(tuple(gidx), _gs)) if _gs else 0
            return ((go // _Nt) % _Pr) * _Pc + (go % _Nt) % _Pc

        Cres._owner = owner
        ContEngine.last_stats = stats
        return stats


# ---------------------------------------------------------------------------------------------
def summa_arrays(world: World, trA: TiledRange, trB: TiledRange, shapeA=None, shapeB=None,
                 memory: str = "device") -> Tuple[DistArray, DistArray]:
    """Create the operands of ``C[m,n] = A[m,k] * B[k,n]`` (matrices) distributed with the
    process grid's cyclic maps (make_row_phase_pmap / make_col_phase_pmap, proc_grid.h:566-597)
    and with arenas ordered so that SUMMA panels are contiguous."""
    Pr, Pc = world.grid or (1, 1)
    Mt, Kt = trA.tiles_shape
    Kt2, Nt = trB.tiles_shape
    _ta_assert(Kt == Kt2, "inner tilings differ")
    ownA = lambda o: ((o // Kt) % Pr) * Pc + (o % Kt) % Pc  # noqa: E731
    ownB = lambda o: ((o // Nt) % Pr) * Pc + (o % Nt) % Pc  # noqa: E731
    orderA = [i * Kt + k for k in range(Kt) for i in range(Mt)]  # column panels contiguous
    orderB = [k * Nt + j for k in range(Kt) for j in range(Nt)]  # row panels contiguous
    return (DistArray(world, trA, shapeA, ownA, orderA, memory), DistArray(world, trB, shapeB, ownB, orderB, memory))
