"""The five BASELINE.json configurations as synthetic workloads of the public API (bench.py, scripts/).

Each builder returns a Workload whose operands are created ALREADY distributed the way SUMMA wants them for
the world's size (tiledarray.contraction_arrays -> tadev_contraction_layout: ProcGrid + cyclic maps over the
fused tile grids) and filled by the device counter RNG (element = f(seed, tile ordinal, offset), so any
distribution holds identical data and any rank can regenerate any operand element on the host for the parity
check, tests/sampled_parity.py).

  C1  dense N=4096 tile=256                                    (examples/gemm/ta_dense.cpp 4096 256)
  C2  dense N=32768 tile=1024                                  (the headline config)
  C3  block-sparse N=65536 tile=512, 10 % random tile density  (examples/gemm/ta_sparse.cpp; true tile norms)
  C3m the same with the example's own expression: (a*b).set_shape(a.shape) and fill(1.0)
  C4  CCSD PPL R(a,b,i,j)=T(c,d,i,j)*V(a,b,c,d) o=100 v=800 tile=64; V (3.28 TB) is a lazy array
  C5  C(i,a,j,b)=A(i,k,a,c)*B(j,c,k,b), i=j=k=128 (tile 16), a=b=c=512 (tile 64); operand tiles permuted
  reduced variants for quick runs: C4r (v=256), C5r (i=j=k=64)
"""
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from tiledarray_b200.tiledarray import (ContEngine, DistArray, SparseShape, TiledRange, TiledRange1, World,
                                        contraction_arrays)


@dataclass
class Workload:
    name: str
    label: str
    target: str
    lidx: str
    ridx: str
    a: DistArray
    b: DistArray
    c: DistArray
    grid: tuple
    seeds: tuple
    flops: float                      # algorithmic flop of one step
    note: str = ""
    mask: Optional[SparseShape] = None
    apparent_flops: Optional[float] = None
    steps_hint: int = 3               # default timed steps / warm-ups for a run that finishes in minutes
    warmup_hint: int = 3
    e2e_ok: bool = True
    extra: dict = field(default_factory=dict)

    def step(self):
        prod = self.a[self.lidx] * self.b[self.ridx]
        if self.mask is not None:
            prod = prod.set_shape(self.mask)
        self.c[self.target] = prod
        return ContEngine.last_stats

    def release(self):
        for x in (self.a, self.b, self.c):
            x.release()


def example_tiling(range_size: int, tile: int) -> TiledRange1:
    """make_uniform_tiling of examples/gemm/ta_cc_abcd.cpp:56-65: tile-size steps, the last tile takes the remainder
    (v=800, tile=64 -> 12 x 64 + 32; NOT TiledRange1::make_uniform)."""
    b = list(range(0, range_size + 1, tile))
    if b[-1] != range_size:
        b.append(range_size)
    return TiledRange1(*b)


def _setup(world: World, target, lidx, trL, ridx, trR, shapeL=None, shapeR=None, memory=("device", "device"),
           lazy_seeds=(None, None)):
    a, b, (Pr, Pc) = contraction_arrays(world, target, lidx, trL, ridx, trR, shapeL, shapeR, memory, lazy_seeds)
    if world.grid is None:
        world.init_comm(Pr, Pc)
    assert world.grid == (Pr, Pc), f"communicators were built for {world.grid}, this expression needs {(Pr, Pc)}"
    return a, b, (Pr, Pc)


def dense(world: World, n: int, tile: int, seeds, name: str, label: str, memory: str = "device") -> Workload:
    t1 = TiledRange1.make_uniform(n, tile)
    tr = TiledRange([t1, t1])
    a, b, grid = _setup(world, "m,n", "m,k", tr, "k,n", tr, memory=(memory, memory))
    if memory == "device":
        a.fill_random(seeds[0])
        b.fill_random(seeds[1])
    c = DistArray(world, tr, memory=memory)
    return Workload(name, label, "m,n", "m,k", "k,n", a, b, c, grid, seeds, 2.0 * float(n) ** 3)


def _true_norm_shape(world: World, arr: DistArray) -> SparseShape:
    """SparseShape from the true Frobenius norms of the local tiles (device reduction) replicated with the
    library's all-reduce(max) (sparse_shape.h:416)."""
    norms = arr.tile_norms().astype(np.float32)
    if world.size > 1:
        norms = world.dev.allreduce_max_f32(norms)
    return SparseShape(world, norms, arr.trange)


def block_sparse(world: World, n: int = 65536, tile: int = 512, density: float = 0.10, seeds=(5, 6), masked: bool = False,
                 name: str = "C3", memory: str = "device") -> Workload:
    t1 = TiledRange1.make_uniform(n, tile)
    tr = TiledRange([t1, t1])
    nt = t1.ntiles

    def pattern(seed):
        rng = np.random.default_rng(seed)
        nz = rng.permutation(nt * nt)[: int(density * nt * nt)]  # first 10 % of a seeded permutation
        p = np.zeros(nt * nt, dtype=np.float32)
        # the example's synthetic norm sqrt(bs^2) (ta_sparse.cpp:150): the Frobenius norm of a tile of ones
        p[nz] = np.float32(np.sqrt(np.float32(tile * tile)))
        return p.reshape(nt, nt)

    shA, shB = SparseShape(world, pattern(seeds[0]), tr), SparseShape(world, pattern(seeds[1]), tr)
    a, b, grid = _setup(world, "m,n", "m,k", tr, "k,n", tr, shA, shB, memory=(memory, memory))
    mask = None
    if masked:  # exactly ta_sparse.cpp:162-190: fill(1.0), synthetic norms, result masked by the left shape
        if memory == "device":
            a.fill(1.0)
            b.fill(1.0)
        mask = shA
    elif memory == "device":
        a.fill_random(seeds[0])
        b.fill_random(seeds[1])
        a.shape, b.shape = _true_norm_shape(world, a), _true_norm_shape(world, b)
    c = DistArray(world, tr, memory=memory)
    za = (a.shape.norms >= np.float32(SparseShape.threshold())).astype(np.int64)
    zb = (b.shape.norms >= np.float32(SparseShape.threshold())).astype(np.int64)
    counts = za @ zb
    if masked:
        counts = counts * (shA.norms >= np.float32(SparseShape.threshold()))
    pairs = int(counts.sum())
    w = Workload(name, f"block-sparse DGEMM N={n} tile={tile} {int(density * 100)}% tile density"
                 + (" , (a*b).set_shape(a.shape) as ta_sparse.cpp:190" if masked else "") + " (BASELINE configs[2])",
                 "m,n", "m,k", "k,n", a, b, c, grid, seeds, 2.0 * pairs * float(tile) ** 3,
                 note=f"{pairs} tile pairs (exact count from the tile lists, ta_sparse.cpp:193 flop convention)",
                 mask=mask, apparent_flops=2.0 * float(n) ** 3)
    w.extra["pairs"] = pairs
    return w


def ccsd_ppl(world: World, o: int = 100, v: int = 800, tile: int = 64, seeds=(7, 8), lazy_v: bool = True, name: str = "C4") -> Workload:
    o1, v1 = example_tiling(o, tile), example_tiling(v, tile)
    trT, trV, trR = TiledRange([v1, v1, o1, o1]), TiledRange([v1, v1, v1, v1]), TiledRange([v1, v1, o1, o1])
    a, b, grid = _setup(world, "a,b,i,j", "c,d,i,j", trT, "a,b,c,d", trV, lazy_seeds=(None, seeds[1] if lazy_v else None))
    a.fill_random(seeds[0])
    if not lazy_v:
        b.fill_random(seeds[1])
    c = DistArray(world, trR)
    full = (o, v) == (100, 800)
    return Workload(name, f"CCSD PPL R(a,b,i,j)=T(c,d,i,j)*V(a,b,c,d) o={o} v={v} tile={tile}"
                    + (" (BASELINE configs[3])" if full else " (reduced v)"), "a,b,i,j", "c,d,i,j", "a,b,c,d", a, b, c, grid,
                    seeds, 2.0 * float(o) ** 2 * float(v) ** 4,
                    note=("V (3.28 TB) is a lazy array generated per tile on the device inside the timed region; " if lazy_v else "")
                    + "operands exchanged by the engine (R[ab,ij] = V[ab,cd] T[cd,ij], NN, no permutation)",
                    steps_hint=1 if (full or lazy_v) else 2, warmup_hint=0 if full else 1, e2e_ok=False)


def permuted_4index(world: World, small: int = 128, big: int = 512, seeds=(9, 10), name: str = "C5") -> Workload:
    s1, b1 = TiledRange1.make_uniform(small, 16), TiledRange1.make_uniform(big, 64)
    trA, trB, trC = TiledRange([s1, s1, b1, b1]), TiledRange([s1, b1, s1, b1]), TiledRange([s1, b1, s1, b1])
    a, b, grid = _setup(world, "i,a,j,b", "i,k,a,c", trA, "j,c,k,b", trB)
    a.fill_random(seeds[0])
    b.fill_random(seeds[1])
    c = DistArray(world, trC)
    full = (small, big) == (128, 512)
    w = Workload(name, f"permuted 4-index C(i,a,j,b)=A(i,k,a,c)*B(j,c,k,b) i=j=k={small} a=b=c={big} tiles 16/64"
                 + (" (BASELINE configs[4])" if full else " (reduced i,j,k)"), "i,a,j,b", "i,k,a,c", "j,c,k,b", a, b, c, grid, seeds,
                 2.0 * (float(small) * float(big)) ** 3,
                 note="both operands need explicit tile permutations (A -> (i,a,c,k), B -> (c,k,j,b)); done per SUMMA window "
                      "by the permute provider when the operand exceeds ContEngine.stream_permute_bytes, else up front",
                 steps_hint=2 if full else 3, warmup_hint=1 if full else 3, e2e_ok=False)
    w.extra["permute_bytes"] = 2.0 * 2.0 * 8.0 * (float(small) * float(big)) ** 2  # read + write of both operands
    return w


def build(world: World, config: str, n: Optional[int] = None, tile: Optional[int] = None, memory: str = "device") -> Workload:
    if config == "C1":
        return dense(world, n or 4096, tile or 256, (1, 2), "C1", f"dense DGEMM N={n or 4096} tile={tile or 256} FP64, c(m,n)=a(m,k)*b(k,n) "
                     "(BASELINE configs[0])", memory)
    if config == "C2":
        return dense(world, n or 32768, tile or 1024, (3, 4), "C2", f"dense DGEMM N={n or 32768} tile={tile or 1024} FP64, "
                     "c(m,n)=a(m,k)*b(k,n) (BASELINE configs[1])", memory)
    if config == "C3":
        return block_sparse(world, n or 65536, tile or 512, memory=memory)
    if config == "C3m":
        return block_sparse(world, n or 65536, tile or 512, masked=True, name="C3m", memory=memory)
    if config == "C4":
        return ccsd_ppl(world)
    if config == "C4h":  # half-size virtual space, V still lazy (205 GB): the config-4 code path in ~15 s
        return ccsd_ppl(world, v=400, lazy_v=True, name="C4h")
    if config == "C4r":
        return ccsd_ppl(world, v=256, lazy_v=False, name="C4r")
    if config == "C5":
        return permuted_4index(world)
    if config == "C5r":
        return permuted_4index(world, small=64, name="C5r")
    raise ValueError(f"unknown config {config}")
