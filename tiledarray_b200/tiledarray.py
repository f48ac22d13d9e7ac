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
        """TiledRange1::make_uniform (tiled_range1.h:289-313): ceil(extent / tile) tiles, as uniform as possible
        (the first tiles are one element larger), e.g. make_uniform(55, 10) == {0,10,19,28,37,46,55}."""
        _ta_assert(extent > 0 and tile > 0, "TiledRange1::make_uniform: positive extent and tile size required")
        ntiles = (extent + tile - 1) // tile
        quot, rem = divmod(extent + ntiles - 1, ntiles)
        avg, nplus = quot - 1, rem + 1
        bounds, e = [], lo
        for i in range(ntiles):
            bounds.append(e)
            e += avg + 1 if i < nplus else avg
        bounds.append(lo + extent)
        return TiledRange1(*bounds)

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

    def _with(self, norms: np.ndarray, nzero: int) -> "SparseShape":
        return SparseShape(self.world, None, None, _prebuilt=(norms.astype(f32), self.size_vectors, int(nzero), self.my_threshold))

    def scale(self, factor: float) -> "SparseShape":
        """SparseShape::scale (sparse_shape.h:1243): norm * |factor|, hard zero below the threshold."""
        out = (self.norms * f32(abs(factor))).astype(f32)
        z = out < f32(self.my_threshold)
        return self._with(np.where(z, f32(0), out), z.sum())

    def add(self, other: "SparseShape", factor: float = 1.0) -> "SparseShape":
        """SparseShape::add / subt (sparse_shape.h:1309,1370,1499): (a + b) * |factor|, thresholded."""
        _ta_assert(self.norms.shape == other.norms.shape, "SparseShape::add: range mismatch")
        out = (self.norms + other.norms).astype(f32)
        if factor != 1.0:
            out = (out * f32(abs(factor))).astype(f32)
        z = out < f32(SparseShape._threshold)
        return self._with(np.where(z, f32(0), out), z.sum())

    def mult(self, other: "SparseShape", factor: float = 1.0) -> "SparseShape":
        """SparseShape::mult (sparse_shape.h:1522,1551): a * b (* |factor|), then scaled by the tile volume
        (scale_tile_norms<ScaleBy::Volume>, :149-217; on the device) and thresholded."""
        _ta_assert(self.norms.shape == other.norms.shape, "SparseShape::mult: range mismatch")
        prod = (self.norms * other.norms).astype(f32)
        if factor != 1.0:
            prod = (prod * f32(abs(factor))).astype(f32)
        dim = len(self.size_vectors)
        if dim == 1:
            left, right = self.size_vectors[0], np.ones(1, dtype=f32)
        else:
            middle = (dim >> 1) + (dim & 1)
            left = _recursive_outer(self.size_vectors[:middle], False)
            right = _recursive_outer(self.size_vectors[middle:], False)
        out, nz = self.world.dev.shape_scale(prod, left, right, SparseShape._threshold)
        return self._with(out, nz)

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

    def tile_norms(self) -> np.ndarray:
        """Frobenius norms of the local tiles over the tile grid (zeros elsewhere), computed on the device."""
        _ta_assert(self.memory == "device", "tile_norms: device-resident arrays only")
        out = np.zeros(max(self.trange.ntiles, 1), dtype=np.float64)
        if self.tiles:
            ords = np.fromiter(self.tiles.keys(), dtype=np.int64, count=len(self.tiles))
            ptrs = np.fromiter((b.ptr for b in self.tiles.values()), dtype=np.uint64, count=len(self.tiles))
            elems = np.fromiter((b.nbytes // 8 for b in self.tiles.values()), dtype=np.int64, count=len(self.tiles))
            out[ords] = np.sqrt(self.world.dev.tile_sqnorms(ptrs, elems))
        return out.reshape(self.trange.tiles_shape)

    def truncate(self) -> "DistArray":
        """DistArray::truncate (dist_array.h:1553, conversions/truncate.h): recompute the shape from the
        true tile norms and drop the tiles that fall below the threshold. Dense arrays are unchanged."""
        if self.shape.is_dense():
            return self
        norms = self.tile_norms().astype(f32)
        if self.world.size > 1:  # shapes are replicated: every rank needs every tile's norm (gop.max, sparse_shape.h:416)
            norms = self.world.dev.allreduce_max_f32(norms)
        new_shape = SparseShape(self.world, norms, self.trange)
        for o in [o for o in self.tiles if new_shape.is_zero(o)]:
            del self.tiles[o]  # views into the arena: the memory is reclaimed with the arena
        self.shape = new_shape
        return self

    def release(self) -> None:
        for x in getattr(self, "_extra_arenas", []):
            x.free()
        self._extra_arenas = []
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
        if expr is _ASSIGNED:  # c["m,n"] += ... already evaluated by TsrExpr.__iadd__
            return
        TsrExpr(self, _split(idx)).assign(expr)


# ---------------------------------------------------------------------------------------------
_ASSIGNED = object()


class Expr:
    factor = 1.0
    mask = None  # result-shape override (Expr::set_shape, expressions/expr.h:116)

    def set_shape(self, shape: "SparseShape"):
        """``(a("m,k") * b("k,n")).set_shape(mask)``: the result shape is masked by ``mask``
        (ContEngine::init_struct, cont_engine.h:526-528 -> SparseShape::mask, sparse_shape.h:653-676)."""
        _ta_assert(isinstance(shape, SparseShape), "set_shape: a SparseShape is required")
        self.mask = shape
        return self

    def __mul__(self, other):
        if isinstance(other, (int, float)):
            return ScalExpr(self, float(other))
        return MultExpr(self, other)

    def __rmul__(self, other):
        _ta_assert(isinstance(other, (int, float)), "unsupported operand")
        return ScalExpr(self, float(other))

    def __add__(self, other):
        return AddExpr(self, other, 1.0)

    def __sub__(self, other):
        return AddExpr(self, other, -1.0)

    def __neg__(self):
        return ScalExpr(self, -1.0)


class TsrExpr(Expr):
    def __init__(self, array: DistArray, indices: List[str]):
        _ta_assert(len(indices) == array.trange.rank, "index list rank does not match the array")
        self.array, self.indices = array, indices

    def __iadd__(self, expr):
        """c("m,n") += a("m,k") * b("k,n"): the product is accumulated into the existing tiles (beta = 1)."""
        self.assign(expr, accumulate=True)
        return _ASSIGNED

    def assign(self, expr, accumulate: bool = False) -> None:
        outer_mask = expr.mask
        factor, expr = _peel(expr)
        mask = expr.mask if expr.mask is not None else outer_mask
        if isinstance(expr, MultExpr):
            fl, left = _peel(expr.left)
            fr, right = _peel(expr.right)
            _ta_assert(isinstance(left, TsrExpr) and isinstance(right, TsrExpr),
                       "product operands must be (scaled) arrays (nested expressions are out of scope)")
            factor *= fl * fr
            if set(left.indices) == set(right.indices) == set(self.indices):
                # every index is shared and kept: Hadamard product (mult_engine.h), no contraction
                _ta_assert(not accumulate, "+= of a Hadamard product is not implemented")
                _ta_assert(mask is None, "set_shape is implemented for contractions only")
                ElementwiseEngine(self, _lib.EW_MULT, factor, left, 1.0, right).eval()
            else:
                ContEngine(self, left, right, factor, accumulate, mask).eval()
            return
        _ta_assert(not accumulate, "+= is implemented for contractions only")
        _ta_assert(mask is None, "set_shape is implemented for contractions only")
        if isinstance(expr, AddExpr):
            fl, left = _peel(expr.left)
            fr, right = _peel(expr.right)
            _ta_assert(isinstance(left, TsrExpr) and isinstance(right, TsrExpr),
                       "sum operands must be (scaled) arrays (nested expressions are out of scope)")
            ElementwiseEngine(self, _lib.EW_AXPBY, factor * fl, left, factor * fr * expr.sign, right).eval()
            return
        _ta_assert(isinstance(expr, TsrExpr), "unsupported expression")
        ElementwiseEngine(self, _lib.EW_AXPBY, factor, expr, 0.0, None).eval()  # c(idx) = f * a(idx'): scale / copy / permute


class ScalExpr(Expr):
    def __init__(self, arg, scalar: float):
        self.arg, self.scalar = arg, scalar


class MultExpr(Expr):
    def __init__(self, left, right):
        self.left, self.right = left, right


class AddExpr(Expr):
    def __init__(self, left, right, sign: float):
        self.left, self.right, self.sign = left, right, sign


def _peel(expr):
    """(accumulated scalar factor, inner expression) of nested ScalExpr nodes."""
    factor = 1.0
    while isinstance(expr, ScalExpr):
        factor *= expr.scalar
        expr = expr.arg
    return factor, expr


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
    gemm_ms: float = 0.0
    list_ms: float = 0.0
    swapped: bool = False


class ElementwiseEngine:
    """c(idx) = alpha * a(idx_a) (+ beta * b(idx_b) | .* b(idx_b)): the reference's AddEngine / SubtEngine /
    ScalEngine / MultEngine-Hadamard (expressions/add_engine.h, subt_engine.h, scal_engine.h, mult_engine.h)
    for device arrays. The engine is native (csrc/cont_engine.cpp behind tadev_elementwise_create/_eval):
    operands whose index order differs from the target are permuted first (one batched launch per tile
    extent); then ONE tadev_tiles_binary_f64 launch produces every result tile. Result shape:
    SparseShape::scale / add / mult. This class only describes the arrays and adopts the result."""

    last_ms: float = 0.0

    def __init__(self, result: "TsrExpr", op: int, alpha: float, a: "TsrExpr", beta: float, b: Optional["TsrExpr"]):
        self.result, self.op, self.alpha, self.a, self.beta, self.b = result, op, alpha, a, beta, b
        self.world = a.array.world
        self.dev = self.world.dev

    def eval(self) -> None:
        dev, w, lib = self.dev, self.world, self.world.lib
        Cres = self.result.array
        _ta_assert(Cres.memory == "device", "element-wise expressions: device-resident result only")
        keep: list = []
        self.a.array._allocate()
        dA = ContEngine._describe(self.a.array, keep)
        dB = None
        if self.b is not None:
            self.b.array._allocate()
            dB = ContEngine._describe(self.b.array, keep)
        handle = C.c_void_p()
        rc = lib.tadev_elementwise_create(dev.ctx, self.op, ",".join(self.result.indices).encode(), self.alpha,
                                          ",".join(self.a.indices).encode(), C.byref(dA), self.beta,
                                          ",".join(self.b.indices).encode() if self.b is not None else None,
                                          C.byref(dB) if dB is not None else None, SparseShape._threshold, C.byref(handle))
        if rc == _lib.EINVAL:
            raise TiledArrayException(lib.tadev_last_error().decode())
        check(rc)
        try:
            info = _lib.ContractionInfoC()
            check(lib.tadev_elementwise_info_get(handle, C.byref(info)))
            dims, off = [], 0
            for d in range(info.rank):
                n = info.ntiles[d]
                dims.append(TiledRange1(*[info.bounds[off + t] for t in range(n + 1)]))
                off += n + 1
            tr = TiledRange(dims)
            nloc = info.nlocal
            ords = np.ctypeslib.as_array(info.ordinals, shape=(nloc,)).copy() if nloc else np.zeros(0, dtype=np.int64)
            elems = np.ctypeslib.as_array(info.elems, shape=(nloc,)).copy() if nloc else np.zeros(0, dtype=np.int64)
            offs = np.ctypeslib.as_array(info.offsets, shape=(nloc,)).copy() if nloc else np.zeros(0, dtype=np.int64)
            if info.norms:
                norms = np.ctypeslib.as_array(info.norms, shape=(max(tr.ntiles, 1),)).copy().reshape(tr.tiles_shape)
                sv = [np.asarray(d.extents, dtype=f32) for d in tr.dims]
                shape = SparseShape(w, None, None, _prebuilt=(norms, sv, int(info.nzero), SparseShape._threshold))
            else:
                shape = DenseShape()
            arena = dev.alloc(int(info.arena_elems) * 8)
            ms = C.c_float()
            check(lib.tadev_elementwise_eval(handle, arena.ptr, C.byref(ms)))
            ElementwiseEngine.last_ms = ms.value
        finally:
            lib.tadev_elementwise_destroy(handle)
        own = self.a.array._owner if w.size > 1 else (lambda o: 0)
        Cres.release()  # the old tiles (possibly operands of this expression: stream-ordered free after the kernel)
        Cres.trange, Cres.shape, Cres._arena = tr, shape, arena
        Cres.tiles = {o_: DeviceBuffer(dev, arena.ptr + f_ * 8, e_ * 8, False)
                      for o_, f_, e_ in zip(ords.tolist(), offs.tolist(), elems.tolist())}
        Cres._owner = own


class _Contraction:
    """Owner of a native tadev_contraction handle (destroyed with the last reference)."""

    def __init__(self, lib, handle):
        self.lib, self.handle = lib, handle

    def owner(self, ordinal: int) -> int:
        o = C.c_int()
        check(self.lib.tadev_contraction_owner(self.handle, ordinal, C.byref(o)))
        return o.value

    def __del__(self):
        if self.handle:
            self.lib.tadev_contraction_destroy(self.handle)
            self.handle = None


class ContEngine:
    """ContEngine (expressions/cont_engine.h:354-677) + Summa hand-off (:662-677).

    The engine itself is native (csrc/cont_engine.cpp behind tadev_contraction_create/_eval): index
    analysis and operand exchange, op flags, permuted operand structure, result trange, device shape
    screening, ProcGrid check, result layout, argument/result tile permutations and the SUMMA call.
    This class only describes the arrays to it and adopts the result.
    """

    last_stats: Optional[ContractionStats] = None
    depth = 0             # SUMMA pipeline depth in windows (0 = default 2; TA_SUMMA_MAX_DEPTH analogue)
    steps_per_launch = 0  # K steps fused into one grouped-GEMM launch (0 = auto)
    row_blocks = 0        # result row blocks (0 = auto)
    exchange_operands = True  # evaluate C^T = B^T A^T when that needs fewer explicit tile permutations
    stream_permutes = "auto"  # True/False/"auto": permute argument tiles just in time per SUMMA window
    stream_permute_bytes = 8 << 30  # "auto": stream when the permuted copy would exceed this many bytes

    def __init__(self, result: TsrExpr, left: TsrExpr, right: TsrExpr, factor: float, accumulate: bool = False,
                 mask: Optional[SparseShape] = None):
        self.result, self.left, self.right, self.factor, self.accumulate = result, left, right, factor, accumulate
        self.mask = mask
        self.world = left.array.world
        self.dev = self.world.dev

    @staticmethod
    def _describe(arr: DistArray, keep: list) -> "_lib.ArrayDescC":
        """tadev_array_desc of a DistArray (the arrays it points to are appended to `keep`)."""
        d = _lib.ArrayDescC()
        tr = arr.trange
        bounds = np.asarray([b for dim in tr.dims for b in dim.bounds], dtype=np.int64)
        ntiles = np.asarray([dim.ntiles for dim in tr.dims], dtype=np.int32)
        keep += [bounds, ntiles]
        d.rank = tr.rank
        d.memory = {"device": _lib.MEM_DEVICE, "host": _lib.MEM_HOST, "lazy": _lib.MEM_LAZY}[arr.memory]
        d.bounds = bounds.ctypes.data_as(C.POINTER(C.c_int64))
        d.ntiles = ntiles.ctypes.data_as(C.POINTER(C.c_int32))
        if not arr.shape.is_dense():
            norms = np.ascontiguousarray(arr.shape.norms, dtype=f32)
            keep.append(norms)
            d.norms = norms.ctypes.data_as(C.POINTER(C.c_float))
        table = np.zeros(max(tr.ntiles, 1), dtype=np.uint64)
        if arr.memory == "lazy":
            d.lazy_seed = arr.lazy_seed
            loc = np.asarray(arr.local_nonzero_ordinals(), dtype=np.int64)
            table[loc] = 1  # marks the local tiles
        elif arr.tiles:
            ords = np.fromiter(arr.tiles.keys(), dtype=np.int64, count=len(arr.tiles))
            table[ords] = np.fromiter((b_.ptr for b_ in arr.tiles.values()), dtype=np.uint64, count=len(arr.tiles))
        keep.append(table)
        d.tiles = table.ctypes.data_as(C.POINTER(C.c_void_p))
        if arr.world.size > 1 and arr.memory == "device":
            # the array's process map: lets the engine redistribute tiles that are not where SUMMA needs them
            owners = np.fromiter((arr._owner(o) for o in range(tr.ntiles)), dtype=np.int32, count=tr.ntiles)
            keep.append(owners)
            d.owners = owners.ctypes.data_as(C.POINTER(C.c_int32))
        return d

    def eval(self) -> ContractionStats:
        w, dev, lib = self.world, self.dev, self.world.lib
        A, B, Cres = self.left.array, self.right.array, self.result.array
        _ta_assert(Cres.memory != "lazy", "the result of a contraction cannot be a lazy array")
        A._allocate()
        B._allocate()
        old_host_arena = None
        old_result = None
        if self.accumulate:
            _ta_assert(Cres is not A and Cres is not B, "c += a*b with c among the arguments is not supported")
            old_result = Cres
        elif Cres is not A and Cres is not B:
            if Cres.memory == "host" and isinstance(Cres._arena, HostBuffer):
                old_host_arena, Cres._arena = Cres._arena, None  # page-locking is slow: recycle the pinned arena
            Cres.release()  # hand the old result's memory back to the pool BEFORE allocating the new one
        keep: list = []
        dA, dB = self._describe(A, keep), self._describe(B, keep)
        opt = _lib.ContractOptionsC()
        check(lib.tadev_contract_options_default(C.byref(opt)))
        opt.exchange_operands = int(ContEngine.exchange_operands)
        opt.stream_permutes = -1 if ContEngine.stream_permutes == "auto" else int(bool(ContEngine.stream_permutes))
        opt.stream_permute_bytes = ContEngine.stream_permute_bytes
        opt.depth, opt.steps_per_launch, opt.row_blocks = ContEngine.depth, ContEngine.steps_per_launch, ContEngine.row_blocks
        opt.threshold = SparseShape._threshold
        if self.mask is not None:
            mask_norms = np.ascontiguousarray(self.mask.norms, dtype=f32)
            keep.append(mask_norms)
            opt.mask_norms = mask_norms.ctypes.data_as(C.POINTER(C.c_float))
            opt.mask_threshold = self.mask.my_threshold
        handle = C.c_void_p()
        rc = lib.tadev_contraction_create(dev.ctx, ",".join(self.result.indices).encode(), ",".join(self.left.indices).encode(),
                                          ",".join(self.right.indices).encode(), C.byref(dA), C.byref(dB), self.factor,
                                          C.byref(opt), C.byref(handle))
        if rc == _lib.EINVAL:  # the reference TA_ASSERTs on malformed expressions
            raise TiledArrayException(lib.tadev_last_error().decode())
        check(rc)
        eng = _Contraction(lib, handle)
        info = _lib.ContractionInfoC()
        check(lib.tadev_contraction_info_get(handle, C.byref(info)))

        # result structure
        dims, off = [], 0
        for d in range(info.rank):
            n = info.ntiles[d]
            dims.append(TiledRange1(*[info.bounds[off + t] for t in range(n + 1)]))
            off += n + 1
        tr_target = TiledRange(dims)
        _ta_assert(Cres.trange == tr_target or not Cres.tiles and Cres.trange.rank == tr_target.rank,
                   "result array tiling does not match the expression")
        nloc = info.nlocal
        ords = np.ctypeslib.as_array(info.ordinals, shape=(nloc,)).copy() if nloc else np.zeros(0, dtype=np.int64)
        elems = np.ctypeslib.as_array(info.elems, shape=(nloc,)).copy() if nloc else np.zeros(0, dtype=np.int64)
        offs = np.ctypeslib.as_array(info.offsets, shape=(nloc,)).copy() if nloc else np.zeros(0, dtype=np.int64)
        if info.norms:
            norms = np.ctypeslib.as_array(info.norms, shape=(max(tr_target.ntiles, 1),)).copy().reshape(tr_target.tiles_shape)
            sv = [np.asarray(d.extents, dtype=f32) for d in tr_target.dims]
            new_shape = SparseShape(w, None, None, _prebuilt=(norms, sv, int(info.nzero), SparseShape._threshold))
        else:
            new_shape = DenseShape()

        # result arena
        c_on_host = Cres.memory == "host"
        need_bytes = int(info.arena_elems) * 8
        st = _lib.ContractStatsC()
        mem_c = _lib.MEM_HOST if c_on_host else _lib.MEM_DEVICE
        if self.accumulate:
            # c += a*b: the product is added into the EXISTING tiles through their own pointers (any arena layout).
            # Product tiles that are zero in c so far are created (zero-filled) first: the result shape is the sum
            # of the two shapes (AddEngine semantics, SparseShape::add).
            _ta_assert(old_result.trange == tr_target, "c += a*b: the existing result must have the tiling of the product")
            _ta_assert(old_result.memory == Cres.memory and old_result.memory != "lazy", "c += a*b: bad result memory")
            old_result._allocate()
            missing = [(o_, e_) for o_, e_ in zip(ords.tolist(), elems.tolist()) if o_ not in old_result.tiles]
            if missing:
                _ta_assert(not old_result.shape.is_dense(), "c += a*b: a dense result is missing tiles")
                tot = sum((e_ + 1) & ~1 for _, e_ in missing)
                extra = HostBuffer(tot * 8) if c_on_host else dev.alloc(tot * 8)
                if c_on_host:
                    extra.numpy(np.float64, (tot,))[:] = 0.0
                else:
                    dev.memset(extra, 0)
                old_result._extra_arenas = getattr(old_result, "_extra_arenas", []) + [extra]
                off = 0
                for o_, e_ in missing:
                    old_result.tiles[o_] = extra.view(off * 8, e_ * 8)
                    off += (e_ + 1) & ~1
            table = (C.c_void_p * max(nloc, 1))(*[old_result.tiles[o_].ptr for o_ in ords.tolist()])
            check(lib.tadev_contraction_eval_tiles(handle, table, mem_c, 1, C.byref(st)))
            if not old_result.shape.is_dense():
                old_result.shape = old_result.shape.add(new_shape)
            arena = None
        else:
            if c_on_host:
                if old_host_arena is not None and old_host_arena.nbytes >= need_bytes:
                    arena = old_host_arena
                else:
                    if old_host_arena is not None:
                        old_host_arena.free()
                    arena = HostBuffer(need_bytes)
            else:
                arena = dev.alloc(need_bytes)
            check(lib.tadev_contraction_eval(handle, arena.ptr, mem_c, 0, C.byref(st)))
        stats = ContractionStats()
        sm = st.summa
        stats.nsteps, stats.nsteps_skipped, stats.npairs = sm.nsteps, sm.nsteps_skipped, sm.npairs
        stats.nlaunches, stats.flops, stats.bcast_bytes, stats.device_ms = sm.nlaunches, sm.flops, sm.bcast_bytes, sm.device_ms
        stats.h2d_bytes, stats.d2h_bytes, stats.row_blocks, stats.lazy_tiles = sm.h2d_bytes, sm.d2h_bytes, sm.row_blocks, sm.lazy_tiles
        stats.gemm_ms, stats.list_ms = sm.gemm_ms, sm.list_ms
        stats.permute_ms = st.permute_ms
        stats.swapped = bool(info.swapped)

        # adopt the result (finalize, contraction_eval.h:1180-1269)
        if not self.accumulate:
            Cres.release()
            Cres.trange, Cres.shape, Cres._arena = tr_target, new_shape, arena
            if c_on_host:
                Cres.tiles = {o_: HostBuffer(e_ * 8, arena.ptr + f_ * 8, owned=False)
                              for o_, f_, e_ in zip(ords.tolist(), offs.tolist(), elems.tolist())}
            else:
                Cres.tiles = {o_: DeviceBuffer(dev, arena.ptr + f_ * 8, e_ * 8, False)
                              for o_, f_, e_ in zip(ords.tolist(), offs.tolist(), elems.tolist())}
            Cres._owner = eng.owner
        ContEngine.last_stats = stats
        return stats


# ---------------------------------------------------------------------------------------------
def summa_arrays(world: World, trA: TiledRange, trB: TiledRange, shapeA=None, shapeB=None,
                 memory: str = "device", lazy_seeds=(None, None)) -> Tuple[DistArray, DistArray]:
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
    memA = "lazy" if lazy_seeds[0] is not None else memory
    memB = "lazy" if lazy_seeds[1] is not None else memory
    return (DistArray(world, trA, shapeA, ownA, orderA, memA, lazy_seeds[0]),
            DistArray(world, trB, shapeB, ownB, orderB, memB, lazy_seeds[1]))


def contraction_layout(world_size: int, target: str, lidx: str, trL: TiledRange, ridx: str, trR: TiledRange,
                       exchange_operands: Optional[bool] = None):
    """[host] tadev_contraction_layout: the ProcGrid and, per tile of each operand, its position in the fused
    GEMM-side tile grid (-> owner rank and panel-contiguous arena order). Needs no device."""
    lib = _lib.load()
    keep: list = []

    def desc(tr):
        d = _lib.ArrayDescC()
        bounds = np.asarray([b for dim in tr.dims for b in dim.bounds], dtype=np.int64)
        ntiles = np.asarray([dim.ntiles for dim in tr.dims], dtype=np.int32)
        keep.extend([bounds, ntiles])
        d.rank, d.memory = tr.rank, _lib.MEM_DEVICE
        d.bounds = bounds.ctypes.data_as(C.POINTER(C.c_int64))
        d.ntiles = ntiles.ctypes.data_as(C.POINTER(C.c_int32))
        return d

    dL, dR = desc(trL), desc(trR)
    opt = _lib.ContractOptionsC()
    check(lib.tadev_contract_options_default(C.byref(opt)))
    opt.exchange_operands = int(ContEngine.exchange_operands if exchange_operands is None else exchange_operands)
    info = _lib.LayoutInfoC()
    lr, lc = np.zeros(max(trL.ntiles, 1), np.int32), np.zeros(max(trL.ntiles, 1), np.int32)
    rr, rc_ = np.zeros(max(trR.ntiles, 1), np.int32), np.zeros(max(trR.ntiles, 1), np.int32)
    rc = lib.tadev_contraction_layout(target.encode(), lidx.encode(), ridx.encode(), C.byref(dL), C.byref(dR), C.byref(opt),
                                      world_size, C.byref(info), lr.ctypes.data, lc.ctypes.data, rr.ctypes.data, rc_.ctypes.data)
    if rc == _lib.EINVAL:
        raise TiledArrayException(lib.tadev_last_error().decode())
    check(rc)
    return info, (lr, lc), (rr, rc_)


def contraction_arrays(world: World, target: str, lidx: str, trL: TiledRange, ridx: str, trR: TiledRange,
                       shapeL=None, shapeR=None, memory=("device", "device"), lazy_seeds=(None, None)):
    """Create the two operands of ``c(target) = a(lidx) * b(ridx)`` already distributed the way SUMMA wants them
    (cyclic maps of the ProcGrid over the fused tile grids, proc_grid.h:566-597) with panel-contiguous arenas, for any
    index lists (permuted / transposed operands, exchanged operands). Returns (a, b, (Pr, Pc)); the caller then
    calls ``world.init_comm(Pr, Pc)``. The analogue of constructing TA::DistArrays on the pmaps ContEngine would pick."""
    info, (lr, lc), (rr, rc_) = contraction_layout(world.size, target, lidx, trL, ridx, trR)
    Pr, Pc = info.Pr, info.Pc
    out = []
    for tr, shape, fr, fc, role, mem, seed in ((trL, shapeL, lr, lc, info.left_role, memory[0], lazy_seeds[0]),
                                               (trR, shapeR, rr, rc_, info.right_role, memory[1], lazy_seeds[1])):
        owners = ((fr.astype(np.int64) % Pr) * Pc + fc.astype(np.int64) % Pc).tolist()
        # role 0 (A(i,k)): column panels contiguous -> order by (k = fcol, i = frow); role 1 (B(k,j)): by (k = frow, j = fcol)
        key = fc.astype(np.int64) * (int(fr.max()) + 1) + fr if role == 0 else fr.astype(np.int64) * (int(fc.max()) + 1) + fc
        order = np.argsort(key, kind="stable").tolist()
        mem = "lazy" if seed is not None else mem
        out.append(DistArray(world, tr, shape, (lambda o, ow=owners: ow[o]), order, mem, seed))
    return out[0], out[1], (Pr, Pc)
