"""The other BASELINE.json configs (parity-test cases, not bench lines): each is run once through
the public API on one GPU, timed with CUDA events, and spot-checked against the oracle.
  C1   dense N=4096 tile=256                         (the reference's CPU-runnable case)
  C3   block-sparse N=65536 tile=512, 10 % tile density, true Frobenius tile norms
  C4r  CCSD PPL R(a,b,i,j)=T(c,d,i,j)*V(a,b,c,d), o=100 (64+36), v REDUCED to 256 (4x64):
       V at v=800 is 3.28 TB and does not fit in HBM; same code path, same tiling
  C5r  C(i,a,j,b)=A(i,k,a,c)*B(j,c,k,b), i=j=k REDUCED to 64 (tile 16), a=b=c=512 (tile 64)
  C4   the full config: o=100 v=800 tile=64. V (3.28 TB) is a LAZY array: its tiles are generated on
       the device from the counter RNG when a SUMMA window needs them; the engine exchanges the
       operands (R[ab,ij] = V[ab,cd] T[cd,ij], plain NN, no permutation) and walks R in row blocks
  C5   the full config: i=j=k=128 (tile 16), a=b=c=512 (tile 64); argument permutations are
       performed just in time per SUMMA window (no permuted copies: 3 x 34.4 GB stay resident)
python scripts/bench_configs.py [C1,C3,C4r,C5r,C4,C5] [out.jsonl]"""
import itertools
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ta_oracle as O  # noqa: E402  (checker only)
from tests import util_rng  # noqa: E402
from tiledarray_b200 import Device  # noqa: E402
from tiledarray_b200.tiledarray import ContEngine, DistArray, SparseShape, TiledRange, TiledRange1, World  # noqa: E402

which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["C1", "C3", "C4r", "C5r"]
outp = sys.argv[2] if len(sys.argv) > 2 else None
dev = Device(0)
world = World(device=dev)
world.init_comm(1, 1)
peak = max(dev.probe_fp64_peak(0, 40000)[0] for _ in range(2))


def example_tiling(range_size, tile):
    """make_uniform_tiling of examples/gemm/ta_cc_abcd.cpp:56-65: tile-size steps, the last tile takes the remainder
    (v=800, tile=64 -> 12 x 64 + 32; NOT TiledRange1::make_uniform, which would give 7 x 62 + 6 x 61)."""
    b = list(range(0, range_size + 1, tile))
    if b[-1] != range_size:
        b.append(range_size)
    return TiledRange1(*b)


def host_tile(arr, ordinal, seed):
    ext = arr.trange.tile_extent(arr.trange.tile_index(ordinal))
    return util_rng.tile_fill(ordinal, int(np.prod(ext)), seed).reshape(ext)


def spot_check(c, a, b, spec, seeds, ordinal):
    """Recompute one result tile on the host from regenerated operand tiles."""
    lidx, rest = spec.split(",")
    ridx, tidx = rest.split("->")
    tix = dict(zip(tidx, c.trange.tile_index(ordinal)))
    inner = [x for x in lidx if x in ridx]
    dims = {x: a.trange.dims[lidx.index(x)].ntiles for x in inner}
    ref = None
    for combo in itertools.product(*[range(dims[x]) for x in inner]):
        env = dict(tix)
        env.update(dict(zip(inner, combo)))
        oa = a.trange.tile_ordinal([env[x] for x in lidx])
        ob = b.trange.tile_ordinal([env[x] for x in ridx])
        if a.is_zero(oa) or b.is_zero(ob):
            continue
        term = np.einsum(spec, host_tile(a, oa, seeds[0]), host_tile(b, ob, seeds[1]), optimize=True)
        ref = term if ref is None else ref + term
    return O.rel_frobenius(c.find(ordinal), ref)


def run(name, a, b, c, target, lidx, ridx, flops, seeds, reps=2, note="", spot="middle"):
    best_total, st = 1e30, None
    for _ in range(reps):
        dev.sync()
        t0 = time.perf_counter()
        with dev.timer() as tm:
            c[target] = a[lidx] * b[ridx]
        best_total = min(best_total, tm.ms)
        st = ContEngine.last_stats
        wall = time.perf_counter() - t0
    spec = f"{lidx.replace(',', '')},{ridx.replace(',', '')}->{target.replace(',', '')}"
    o = sorted(c.tiles)[len(c.tiles) // 2] if spot == "middle" else sorted(c.tiles)[-1]
    t0 = time.perf_counter()
    err = spot_check(c, a, b, spec, seeds, o)
    spot_s = time.perf_counter() - t0
    rec = {"config": name, "note": note, "algorithmic_flop": flops, "executed_flop": st.flops, "npairs": st.npairs,
           "gemm_launches": st.nlaunches, "row_blocks": st.row_blocks, "lazy_tiles": st.lazy_tiles, "spot_check_s": spot_s,
           "gemm_ms": st.device_ms, "permute_ms": st.permute_ms, "total_ms": best_total,
           "wall_s_last": wall, "tflops_effective": flops / (best_total * 1e-3) / 1e12,
           "tflops_gemm": st.flops / (st.device_ms * 1e-3) / 1e12, "fp64_peak_tflops": peak,
           "frac_of_peak_gemm": st.flops / (st.device_ms * 1e-3) / 1e12 / peak, "spot_rel_frobenius": err}
    print(json.dumps(rec), flush=True)
    if outp:
        with open(outp, "a") as f:
            f.write(json.dumps(rec) + "\n")
    assert err < 1e-12, err


if "C1" in which:
    t = TiledRange1.make_uniform(4096, 256)
    tr = TiledRange([t, t])
    a, b, c = DistArray(world, tr).fill_random(1), DistArray(world, tr).fill_random(2), DistArray(world, tr)
    run("C1 dense N=4096 tile=256", a, b, c, "m,n", "m,k", "k,n", 2.0 * 4096 ** 3, (1, 2), reps=3)
    for x in (a, b, c):
        x.release()

if "C3" in which:
    N, T, dens = 65536, 512, 0.10
    t = TiledRange1.make_uniform(N, T)
    tr = TiledRange([t, t])
    nt = t.ntiles

    def sparse(seed):
        rng = np.random.default_rng(seed)
        nz = rng.permutation(nt * nt)[: int(dens * nt * nt)]  # first 10 % of a seeded permutation
        pattern = np.zeros(nt * nt, dtype=np.float32)
        pattern[nz] = 1.0
        arr = DistArray(world, tr, SparseShape(world, pattern.reshape(nt, nt) * T, tr, do_not_scale=False))
        arr.fill_random(seed)
        # true Frobenius norms of the tiles, computed on the device
        import ctypes as C
        ords = sorted(arr.tiles)
        ptrs = np.array([arr.tiles[o].ptr for o in ords], dtype=np.uint64)
        sizes = np.full(len(ords), T * T, dtype=np.int64)
        d_p, d_s, d_o = dev.upload(ptrs), dev.upload(sizes), dev.alloc(8 * len(ords))
        from tiledarray_b200._lib import check
        check(dev.lib.tadev_tile_sqnorms_f64(dev.ctx, dev.stream, len(ords), d_p.ptr, d_s.ptr, d_o.ptr))
        sq = dev.download(d_o, np.float64, (len(ords),))
        norms = np.zeros(nt * nt, dtype=np.float32)
        norms[ords] = np.sqrt(sq).astype(np.float32)
        arr.shape = SparseShape(world, norms.reshape(nt, nt), tr)
        for x in (d_p, d_s, d_o):
            x.free()
        return arr

    a, b = sparse(5), sparse(6)
    c = DistArray(world, tr)
    za = (a.shape.norms >= np.float32(SparseShape.threshold())).astype(np.int64)
    zb = (b.shape.norms >= np.float32(SparseShape.threshold())).astype(np.int64)
    pairs = int((za @ zb).sum())
    for variant in os.environ.get("C3_VARIANTS", "default").split(","):
        if variant == "raster0":
            os.environ["TADEV_RASTER_S"] = "0"
        else:
            os.environ.pop("TADEV_RASTER_S", None)
        run(f"C3 block-sparse N=65536 tile=512 density=10% [{variant}]", a, b, c, "m,n", "m,k", "k,n", 2.0 * pairs * T ** 3, (5, 6),
            note=f"{pairs} tile pairs (exact count from the tile lists); apparent 2N^3 = {2.0 * N ** 3:.3e}")
    print(json.dumps({"config": "C3", "result_nnz_tiles": len(c.tiles), "result_sparsity": c.shape.sparsity()}), flush=True)
    for x in (a, b, c):
        x.release()

if "C4r" in which:
    o1 = example_tiling(100, 64)
    v1 = example_tiling(256, 64)
    T2 = DistArray(world, TiledRange([v1, v1, o1, o1])).fill_random(7)
    V = DistArray(world, TiledRange([v1, v1, v1, v1])).fill_random(8)
    R = DistArray(world, TiledRange([v1, v1, o1, o1]))
    run("C4r CCSD PPL o=100 v=256 tile=64", T2, V, R, "a,b,i,j", "c,d,i,j", "a,b,c,d", 2.0 * 100 ** 2 * 256 ** 4, (7, 8),
        note="v reduced from 800 (V would be 3.28 TB); opA=T, opB=T, result permute (i,j,a,b)->(a,b,i,j)")
    for x in (T2, V, R):
        x.release()

if "C4" in which:
    o1 = example_tiling(100, 64)
    v1 = example_tiling(800, 64)
    T2 = DistArray(world, TiledRange([v1, v1, o1, o1])).fill_random(7)
    V = DistArray(world, TiledRange([v1, v1, v1, v1]), memory="lazy", lazy_seed=8)
    R = DistArray(world, TiledRange([v1, v1, o1, o1]))
    run("C4 CCSD PPL o=100 v=800 tile=64", T2, V, R, "a,b,i,j", "c,d,i,j", "a,b,c,d", 2.0 * 100 ** 2 * 800.0 ** 4, (7, 8), reps=1,
        note="full size; V (3.28 TB) lazy: generated per tile on the device inside the timed region; operands exchanged "
             "(NN, no permutation); R in row blocks", spot="last")
    for x in (T2, V, R):
        x.release()

if "C5" in which:
    s1 = TiledRange1.make_uniform(128, 16)
    b1 = TiledRange1.make_uniform(512, 64)
    A = DistArray(world, TiledRange([s1, s1, b1, b1])).fill_random(9)
    B = DistArray(world, TiledRange([s1, b1, s1, b1])).fill_random(10)
    Cc = DistArray(world, TiledRange([s1, b1, s1, b1]))
    run("C5 permuted 4-index i=j=k=128 a=b=c=512 tiles 16/64", A, B, Cc, "i,a,j,b", "i,k,a,c", "j,c,k,b",
        2.0 * (128.0 * 512) ** 3, (9, 10), reps=1,
        note="full size; both operands explicitly permuted (general), just in time per SUMMA window (permute provider); "
             "permute time is inside gemm_ms")
    for x in (A, B, Cc):
        x.release()

if "C5r" in which:
    s1 = TiledRange1.make_uniform(64, 16)
    b1 = TiledRange1.make_uniform(512, 64)
    A = DistArray(world, TiledRange([s1, s1, b1, b1])).fill_random(9)
    B = DistArray(world, TiledRange([s1, b1, s1, b1])).fill_random(10)
    Cc = DistArray(world, TiledRange([s1, b1, s1, b1]))
    run("C5r permuted 4-index i=j=k=64 a=b=c=512 tiles 16/64", A, B, Cc, "i,a,j,b", "i,k,a,c", "j,c,k,b",
        2.0 * (64 * 512) ** 3, (9, 10), note="i,j,k reduced from 128; both operands explicitly permuted (general)")
    for x in (A, B, Cc):
        x.release()
dev.close()
