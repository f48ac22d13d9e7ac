"""Thin object wrapper over the libtadev C ABI: one ``Device`` = one ``tadev_ctx`` = one GPU.

Host data are numpy arrays; device tiles are opaque integer device pointers wrapped in
``DeviceBuffer``. Nothing here computes on the CPU: every method forwards to a C-ABI entry.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import GemmGroup, GemmTask, OP_N, OP_T, check


class DeviceBuffer:
    """A device allocation (stream-ordered pool) holding ``nbytes`` bytes."""

    __slots__ = ("dev", "ptr", "nbytes", "_owned")

    def __init__(self, dev: "Device", ptr: int, nbytes: int, owned: bool = True):
        self.dev, self.ptr, self.nbytes, self._owned = dev, ptr, nbytes, owned

    def view(self, byte_offset: int, nbytes: int) -> "DeviceBuffer":
        assert 0 <= byte_offset and byte_offset + nbytes <= self.nbytes
        return DeviceBuffer(self.dev, self.ptr + byte_offset, nbytes, owned=False)

    def free(self) -> None:
        if self._owned and self.ptr and self.dev.ctx:
            self.dev.free(self)
        self.ptr = 0


class HostBuffer:
    """Pinned (page-locked) host memory from tadev_host_alloc: the home of host-resident tiles."""

    __slots__ = ("ptr", "nbytes", "_owned")

    def __init__(self, nbytes: int = 0, ptr: int = 0, owned: bool = True):
        if owned:
            p = C.c_void_p()
            check(_lib.load().tadev_host_alloc(max(nbytes, 1), C.byref(p)))
            ptr = p.value
        self.ptr, self.nbytes, self._owned = ptr, nbytes, owned

    def view(self, byte_offset: int, nbytes: int) -> "HostBuffer":
        assert 0 <= byte_offset and byte_offset + nbytes <= self.nbytes
        return HostBuffer(nbytes, self.ptr + byte_offset, owned=False)

    def numpy(self, dtype, shape) -> np.ndarray:
        n = int(np.prod(shape, dtype=np.int64))
        assert n * np.dtype(dtype).itemsize <= self.nbytes
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)

    def free(self) -> None:
        if self._owned and self.ptr:
            check(_lib.load().tadev_host_free(self.ptr))
        self.ptr = 0


class Device:
    def __init__(self, device: int = 0, pool_bytes: int = 0):
        self.lib = _lib.load()
        ctx = C.c_void_p()
        check(self.lib.tadev_init(device, pool_bytes, C.byref(ctx)))
        self.ctx = ctx
        self.device = device
        n = C.c_int()
        check(self.lib.tadev_num_streams(self.ctx, C.byref(n)))
        self.streams = []
        for i in range(n.value):
            s = C.c_void_p()
            check(self.lib.tadev_get_stream(self.ctx, i, C.byref(s)))
            self.streams.append(s)
        self.stream = self.streams[0]

    # ---- lifetime -------------------------------------------------------------------------
    def close(self) -> None:
        if self.ctx:
            check(self.lib.tadev_finalize(self.ctx))
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def sync(self, stream=None) -> None:
        check(self.lib.tadev_stream_sync(self.ctx, stream or self.stream))

    # ---- timing ---------------------------------------------------------------------------
    class _Timer:
        def __init__(self, dev, stream):
            self.dev, self.stream = dev, stream
            self.e0, self.e1 = C.c_void_p(), C.c_void_p()
            check(dev.lib.tadev_event_create(dev.ctx, C.byref(self.e0)))
            check(dev.lib.tadev_event_create(dev.ctx, C.byref(self.e1)))
            self.ms = None

        def __enter__(self):
            check(self.dev.lib.tadev_event_record(self.dev.ctx, self.e0, self.stream))
            return self

        def __exit__(self, *exc):
            check(self.dev.lib.tadev_event_record(self.dev.ctx, self.e1, self.stream))
            ms = C.c_float()
            check(self.dev.lib.tadev_event_elapsed_ms(self.dev.ctx, self.e0, self.e1, C.byref(ms)))
            self.ms = ms.value
            self.dev.lib.tadev_event_destroy(self.dev.ctx, self.e0)
            self.dev.lib.tadev_event_destroy(self.dev.ctx, self.e1)

    def timer(self, stream=None) -> "Device._Timer":
        """CUDA-event timer on ``stream`` (default: the ctx's first compute stream)."""
        return Device._Timer(self, stream or self.stream)

    # ---- memory ---------------------------------------------------------------------------
    def alloc(self, nbytes: int, stream=None) -> DeviceBuffer:
        p = C.c_void_p()
        check(self.lib.tadev_alloc(self.ctx, nbytes, C.byref(p), stream or self.stream))
        return DeviceBuffer(self, p.value or 0, nbytes)

    def free(self, buf: DeviceBuffer, stream=None) -> None:
        check(self.lib.tadev_free(self.ctx, buf.ptr, stream or self.stream))

    def upload(self, arr: np.ndarray, stream=None) -> DeviceBuffer:
        arr = np.ascontiguousarray(arr)
        buf = self.alloc(arr.nbytes, stream)
        if arr.nbytes:
            check(self.lib.tadev_memcpy_h2d(self.ctx, buf.ptr, arr.ctypes.data, arr.nbytes, stream or self.stream))
            self.sync(stream)  # pageable source: keep it alive until the copy is done
        return buf

    def upload_into(self, buf: DeviceBuffer, arr: np.ndarray, stream=None, sync: bool = True) -> None:
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= buf.nbytes
        check(self.lib.tadev_memcpy_h2d(self.ctx, buf.ptr, arr.ctypes.data, arr.nbytes, stream or self.stream))
        if sync:
            self.sync(stream)

    def download(self, buf: DeviceBuffer, dtype, shape, stream=None) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= buf.nbytes, (out.nbytes, buf.nbytes)
        if out.nbytes:
            check(self.lib.tadev_memcpy_d2h(self.ctx, out.ctypes.data, buf.ptr, out.nbytes, stream or self.stream))
        self.sync(stream)
        return out

    def memset(self, buf: DeviceBuffer, byte: int = 0, stream=None) -> None:
        check(self.lib.tadev_memset(self.ctx, buf.ptr, byte, buf.nbytes, stream or self.stream))

    def fill_uniform(self, buf: DeviceBuffer, n: int, seed: int, offset: int = 0, stream=None) -> None:
        check(self.lib.tadev_fill_uniform_f64(self.ctx, stream or self.stream, buf.ptr, n, seed, offset))

    # ---- tile GEMM ------------------------------------------------------------------------
    def gemm(self, opA: int, opB: int, m: int, n: int, k: int, alpha: float, A: DeviceBuffer, B: DeviceBuffer,
             beta: float, Cbuf: DeviceBuffer, stream=None) -> None:
        check(self.lib.tadev_gemm_f64(self.ctx, stream or self.stream, opA, opB, m, n, k, alpha, A.ptr, B.ptr, beta,
                                      Cbuf.ptr))

    @staticmethod
    def make_groups(groups: Sequence[tuple]):
        """Pack [(C_ptr, m, n, accumulate, [(A_ptr, B_ptr, k), ...]), ...] into C-ABI descriptor arrays."""
        ng = len(groups)
        nt = sum(len(g[4]) for g in groups)
        G = (GemmGroup * max(ng, 1))()
        T = (GemmTask * max(nt, 1))()
        t = 0
        for gi, (cptr, m, n, acc, tl) in enumerate(groups):
            G[gi] = GemmGroup(cptr, m, n, t, t + len(tl), int(bool(acc)), 0)
            for (a, b, k) in tl:
                T[t] = GemmTask(a, b, k, 0)
                t += 1
        return G, ng, T, nt

    def gemm_grouped_packed(self, opA: int, opB: int, alpha: float, packed, stream=None) -> None:
        G, ng, T, nt = packed
        check(self.lib.tadev_gemm_grouped_f64(self.ctx, stream or self.stream, opA, opB, alpha, G, ng, T, nt))

    def gemm_grouped(self, opA: int, opB: int, alpha: float, groups: Sequence[tuple], stream=None) -> None:
        """groups: sequence of (C_ptr, m, n, accumulate, [(A_ptr, B_ptr, k), ...])."""
        self.gemm_grouped_packed(opA, opB, alpha, self.make_groups(groups), stream)

    # ---- permutation / elementwise ---------------------------------------------------------
    def permute(self, extent: Sequence[int], perm: Sequence[int], elem_bytes: int, src: DeviceBuffer,
                dst: DeviceBuffer, stream=None) -> None:
        rank = len(extent)
        ext = (C.c_int64 * max(rank, 1))(*extent)
        pm = (C.c_int32 * max(rank, 1))(*perm)
        check(self.lib.tadev_permute(self.ctx, stream or self.stream, rank, ext, pm, elem_bytes, src.ptr, dst.ptr))

    def permute_batched(self, extent: Sequence[int], perm: Sequence[int], elem_bytes: int,
                        srcs: Sequence[DeviceBuffer], dsts: Sequence[DeviceBuffer], stream=None) -> None:
        """One launch for many tiles of identical extents (tadev_permute_batched)."""
        rank, n = len(extent), len(srcs)
        ext = (C.c_int64 * max(rank, 1))(*extent)
        pm = (C.c_int32 * max(rank, 1))(*perm)
        ins = (C.c_void_p * max(n, 1))(*[b.ptr for b in srcs])
        outs = (C.c_void_p * max(n, 1))(*[b.ptr for b in dsts])
        check(self.lib.tadev_permute_batched(self.ctx, stream or self.stream, rank, ext, pm, elem_bytes, n, ins, outs))

    def permute_batched_ptrs(self, extent: Sequence[int], perm: Sequence[int], elem_bytes: int, src_ptrs: np.ndarray,
                             dst_ptrs: np.ndarray, stream=None) -> None:
        """tadev_permute_batched on arrays of raw device pointers (uint64)."""
        rank, n = len(extent), len(src_ptrs)
        ext = (C.c_int64 * max(rank, 1))(*extent)
        pm = (C.c_int32 * max(rank, 1))(*perm)
        ins = np.ascontiguousarray(src_ptrs, dtype=np.uint64)
        outs = np.ascontiguousarray(dst_ptrs, dtype=np.uint64)
        vpp = C.POINTER(C.c_void_p)
        check(self.lib.tadev_permute_batched(self.ctx, stream or self.stream, rank, ext, pm, elem_bytes, n,
                                             ins.ctypes.data_as(vpp), outs.ctypes.data_as(vpp)))

    def tiles_binary(self, op: int, out_ptrs: np.ndarray, x_ptrs: np.ndarray, y_ptrs: np.ndarray, elems: np.ndarray,
                     alpha: float, beta: float, stream=None) -> None:
        """tadev_tiles_binary_f64 on arrays of raw device pointers (0 = zero tile)."""
        n = len(out_ptrs)
        vpp, ip = C.POINTER(C.c_void_p), C.POINTER(C.c_int64)
        o = np.ascontiguousarray(out_ptrs, dtype=np.uint64)
        x = np.ascontiguousarray(x_ptrs, dtype=np.uint64)
        y = np.ascontiguousarray(y_ptrs, dtype=np.uint64)
        e = np.ascontiguousarray(elems, dtype=np.int64)
        check(self.lib.tadev_tiles_binary_f64(self.ctx, stream or self.stream, op, n, o.ctypes.data_as(vpp), x.ctypes.data_as(vpp),
                                              y.ctypes.data_as(vpp), e.ctypes.data_as(ip), alpha, beta))

    def tile_sqnorms(self, ptrs: np.ndarray, elems: np.ndarray, stream=None) -> np.ndarray:
        """Squared Frobenius norms of tiles given by raw device pointers (tadev_tile_sqnorms_f64)."""
        n = len(ptrs)
        if n == 0:
            return np.zeros(0)
        ptrs = np.ascontiguousarray(ptrs, dtype=np.uint64)
        elems = np.ascontiguousarray(elems, dtype=np.int64)
        out = np.empty(n)
        for lo in range(0, n, 65535):  # the entry takes at most 65535 tiles per launch
            hi = min(n, lo + 65535)
            d_p = self.upload(ptrs[lo:hi], stream)
            d_s = self.upload(elems[lo:hi], stream)
            d_o = self.alloc(8 * (hi - lo), stream)
            check(self.lib.tadev_tile_sqnorms_f64(self.ctx, stream or self.stream, hi - lo, d_p.ptr, d_s.ptr,
                                                  int(elems[lo:hi].max()), d_o.ptr))
            out[lo:hi] = self.download(d_o, np.float64, (hi - lo,), stream)
            for b in (d_p, d_s, d_o):
                b.free()
        return out

    def add_to(self, n: int, result: DeviceBuffer, arg: DeviceBuffer, stream=None) -> None:
        check(self.lib.tadev_add_to_f64(self.ctx, stream or self.stream, n, result.ptr, arg.ptr))

    def scale(self, n: int, x: DeviceBuffer, factor: float, stream=None) -> None:
        check(self.lib.tadev_scale_f64(self.ctx, stream or self.stream, n, x.ptr, factor))

    # ---- shapes ---------------------------------------------------------------------------
    def shape_scale(self, norms: np.ndarray, left: np.ndarray, right: Optional[np.ndarray], threshold: float):
        """Device SparseShape::scale_tile_norms<InverseVolume>; returns (scaled norms, zero count)."""
        norms = np.ascontiguousarray(norms, dtype=np.float32)
        d_n = self.upload(norms.ravel())
        d_l = self.upload(np.ascontiguousarray(left, dtype=np.float32))
        d_r = self.upload(np.ascontiguousarray(right, dtype=np.float32)) if right is not None else None
        d_z = self.upload(np.zeros(1, dtype=np.uint64))
        check(self.lib.tadev_shape_scale_f32(self.ctx, self.stream, d_n.ptr, d_l.ptr, len(left),
                                             d_r.ptr if d_r else None, len(right) if right is not None else 0,
                                             threshold, d_z.ptr))
        out = self.download(d_n, np.float32, norms.shape)
        nz = int(self.download(d_z, np.uint64, (1,))[0])
        for b in (d_n, d_l, d_r, d_z):
            if b:
                b.free()
        return out, nz

    def shape_gemm(self, a: np.ndarray, b: np.ndarray, ksz: np.ndarray, abs_factor: float, threshold: float):
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        Kt = len(ksz)
        if Kt:
            Mt, Nt = a.shape[0], b.shape[1]
        else:
            Mt, Nt = a.size, b.size
        d_a, d_b = self.upload(a.ravel()), self.upload(b.ravel())
        d_k = self.upload(np.ascontiguousarray(ksz, dtype=np.float32)) if Kt else None
        d_o = self.alloc(4 * Mt * Nt)
        d_z = self.upload(np.zeros(1, dtype=np.uint64))
        check(self.lib.tadev_shape_gemm_f32(self.ctx, self.stream, Mt, Nt, Kt, d_a.ptr, d_b.ptr,
                                            d_k.ptr if d_k else None, abs_factor, threshold, d_o.ptr, d_z.ptr))
        out = self.download(d_o, np.float32, (Mt, Nt))
        nz = int(self.download(d_z, np.uint64, (1,))[0])
        for x in (d_a, d_b, d_k, d_o, d_z):
            if x:
                x.free()
        return out, nz

    def allreduce_max_f32(self, x: np.ndarray) -> np.ndarray:
        """Element-wise max over all ranks through the library's NCCL world communicator (shape replication)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        d = self.upload(x.ravel())
        check(self.lib.tadev_shape_allreduce_max_f32(self.ctx, self.stream, d.ptr, x.size))
        out = self.download(d, np.float32, x.shape)
        d.free()
        return out

    def build_pairlist(self, k: int, Pr: int, Pc: int, r: int, c: int, a: Optional[np.ndarray],
                       b: Optional[np.ndarray], cn: Optional[np.ndarray], Mt: int, Nt: int, Kt: int, threshold: float):
        d_a = self.upload(np.ascontiguousarray(a, dtype=np.float32).ravel()) if a is not None else None
        d_b = self.upload(np.ascontiguousarray(b, dtype=np.float32).ravel()) if b is not None else None
        d_c = self.upload(np.ascontiguousarray(cn, dtype=np.float32).ravel()) if cn is not None else None
        cap = max(1, Mt * Nt)
        d_i, d_j, d_n = self.alloc(4 * cap), self.alloc(4 * cap), self.alloc(4)
        check(self.lib.tadev_build_pairlist(self.ctx, self.stream, k, Pr, Pc, r, c, Mt, Nt, Kt,
                                            d_a.ptr if d_a else None, d_b.ptr if d_b else None,
                                            d_c.ptr if d_c else None, threshold, d_i.ptr, d_j.ptr, d_n.ptr))
        n = int(self.download(d_n, np.int32, (1,))[0])
        pi = self.download(d_i, np.int32, (cap,))[:n]
        pj = self.download(d_j, np.int32, (cap,))[:n]
        for x in (d_a, d_b, d_c, d_i, d_j, d_n):
            if x:
                x.free()
        return pi, pj

    def build_tile_lists(self, ksteps: Sequence[int], Pr: int, Pc: int, r: int, c: int, a: Optional[np.ndarray],
                         b: Optional[np.ndarray], cn: Optional[np.ndarray], Mt: int, Nt: int, Kt: int, threshold: float):
        """tadev_build_tile_lists: (group_begin [nrl*ncl + 1], task_k [ntasks]) of a window of SUMMA steps."""
        d_a = self.upload(np.ascontiguousarray(a, dtype=np.float32).ravel()) if a is not None else None
        d_b = self.upload(np.ascontiguousarray(b, dtype=np.float32).ravel()) if b is not None else None
        d_c = self.upload(np.ascontiguousarray(cn, dtype=np.float32).ravel()) if cn is not None else None
        ks = np.ascontiguousarray(ksteps, dtype=np.int32)
        d_k = self.upload(ks) if len(ks) else None
        nrl = (Mt - r + Pr - 1) // Pr if r < Mt else 0
        ncl = (Nt - c + Pc - 1) // Pc if c < Nt else 0
        ng = nrl * ncl
        cap = max(1, ng * max(len(ks), 1))
        d_g, d_t, d_n = self.alloc(4 * (ng + 1)), self.alloc(4 * cap), self.alloc(4)
        check(self.lib.tadev_build_tile_lists(self.ctx, self.stream, Pr, Pc, r, c, Mt, Nt, Kt, d_k.ptr if d_k else None, len(ks),
                                              d_a.ptr if d_a else None, d_b.ptr if d_b else None, d_c.ptr if d_c else None,
                                              threshold, d_g.ptr, d_t.ptr, cap, d_n.ptr))
        n = int(self.download(d_n, np.int32, (1,))[0])
        gb = self.download(d_g, np.int32, (ng + 1,))
        tk = self.download(d_t, np.int32, (cap,))[:n]
        for x in (d_a, d_b, d_c, d_k, d_g, d_t, d_n):
            if x:
                x.free()
        return gb, tk

    # ---- probes ---------------------------------------------------------------------------
    def probe_fp64_peak(self, kind: int = 0, iters: int = 20000):
        t, ms = C.c_double(), C.c_float()
        check(self.lib.tadev_probe_fp64_peak(self.ctx, kind, iters, C.byref(t), C.byref(ms)))
        return t.value, ms.value

    def probe_copy_gbs(self, nbytes: int = 1 << 30, iters: int = 5) -> float:
        g = C.c_double()
        check(self.lib.tadev_probe_copy_gbs(self.ctx, nbytes, iters, C.byref(g)))
        return g.value

    def probe_pcie_gbs(self, nbytes: int = 1 << 30) -> dict:
        v = [C.c_double() for _ in range(4)]
        check(self.lib.tadev_probe_pcie_gbs(self.ctx, nbytes, *[C.byref(x) for x in v]))
        return dict(zip(("h2d", "d2h", "h2d_bidir", "d2h_bidir"), (x.value for x in v)))

    def launch_count(self) -> int:
        n = C.c_int64()
        check(self.lib.tadev_launch_count(self.ctx, C.byref(n)))
        return n.value


def device_count() -> int:
    n = C.c_int()
    check(_lib.load().tadev_device_count(C.byref(n)))
    return n.value
