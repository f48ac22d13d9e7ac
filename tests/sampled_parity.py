"""Sampled-element parity check of a contraction result against a host recomputation (numpy).

Test infrastructure (used by tests/, bench.py's parity leg and scripts/): for a result tile of
``c(target) = a(lidx) * b(ridx)`` whose operands were filled by DistArray.fill_random / lazy seeds
(element value = f(seed, tile ordinal << 32 + offset in tile), tests/util_rng.py), pick a few random
positions along every target index, regenerate ONLY the operand elements those result elements depend
on (all contracted positions, the sampled outer positions) and contract them with numpy.einsum. The
cost per tile is O(samples x contracted extent), independent of the tile size, so every rank of a
multi-GPU run can verify many of its own tiles at BASELINE's full sizes; random positions inside the
tile expose block- or index-level errors that fill(1)/linearity properties cannot see.
"""
import itertools

import numpy as np

from tests import util_rng


def fill_uniform_at(offsets: np.ndarray, seed: int) -> np.ndarray:
    """tadev_fill_uniform_f64's value at arbitrary global element offsets (any shape)."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        base = (np.uint64(seed) * np.uint64(0xD1342543DE82EF95)) & M
        r = util_rng._splitmix64((base + offsets.astype(np.uint64)) & M)
    return (r >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0


def _operand_slice(arr, idx_names, tile_idx, positions, seed):
    """Elements of tile `tile_idx` of `arr` at the cartesian product of positions[name] (per dimension)."""
    ext = arr.trange.tile_extent(tile_idx)
    strides = [1] * len(ext)
    for d in range(len(ext) - 2, -1, -1):
        strides[d] = strides[d + 1] * ext[d + 1]
    off = np.zeros((), dtype=np.uint64)
    for d, name in enumerate(idx_names):
        p = np.asarray(positions[name], dtype=np.uint64) * np.uint64(strides[d])
        off = off[..., None] + p  # broadcast outer sum
    if seed is None:  # DistArray.fill(1.0)
        return np.ones(off.shape)
    ordinal = arr.trange.tile_ordinal(tile_idx)
    return fill_uniform_at(off + np.uint64(ordinal << 32), seed)


def check_tile(c, a, b, target, lidx, ridx, seeds, ordinal, nsample, rng, factor=1.0):
    """(sum |diff|^2, sum |ref|^2, number of operand tile pairs) over sampled elements of result tile `ordinal`."""
    T, L, R = target.split(","), lidx.split(","), ridx.split(",")
    tidx = c.trange.tile_index(ordinal)
    text = c.trange.tile_extent(tidx)
    positions = {}
    tile_of = {}
    for d, name in enumerate(T):
        n = min(nsample, text[d])
        positions[name] = np.sort(rng.choice(text[d], size=n, replace=False))
        tile_of[name] = tidx[d]
    inner = [x for x in L if x in R]
    inner_dims = [a.trange.dims[L.index(x)] for x in inner]
    names = {x: chr(97 + i) for i, x in enumerate(dict.fromkeys(T + L + R))}
    spec = f"{''.join(names[x] for x in L)},{''.join(names[x] for x in R)}->{''.join(names[x] for x in T)}"
    ref = np.zeros([len(positions[x]) for x in T])
    npairs = 0
    for combo in itertools.product(*[range(d.ntiles) for d in inner_dims]):
        env = dict(tile_of)
        env.update(dict(zip(inner, combo)))
        ia, ib = tuple(env[x] for x in L), tuple(env[x] for x in R)
        if a.is_zero(a.trange.tile_ordinal(ia)) or b.is_zero(b.trange.tile_ordinal(ib)):
            continue
        pos = dict(positions)
        for x, dim, t in zip(inner, inner_dims, combo):
            pos[x] = np.arange(dim.tile_extent(t))
        ref += np.einsum(spec, _operand_slice(a, L, ia, pos, seeds[0]), _operand_slice(b, R, ib, pos, seeds[1]), optimize=True)
        npairs += 1
    got = c.find(ordinal)[np.ix_(*[positions[x] for x in T])]
    ref *= factor
    return float(np.sum((got - ref) ** 2)), float(np.sum(ref ** 2)), npairs


def check_local_tiles(c, a, b, target, lidx, ridx, seeds, ntiles=16, nsample=4, seed=1234, factor=1.0):
    """Verify up to `ntiles` of this rank's result tiles (evenly spread over the sorted local ordinals, first and
    last included, so every row block of the driver is covered). Returns a dict with the worst per-tile relative
    Frobenius error over the sampled elements, the number of tiles and elements checked."""
    ords = sorted(c.tiles)
    if not ords:
        return {"worst_rel": 0.0, "tiles": 0, "elements": 0}
    pick = sorted({ords[int(round(x))] for x in np.linspace(0, len(ords) - 1, min(ntiles, len(ords)))})
    rng = np.random.default_rng(seed + 7919 * c.world.rank)
    worst, nel = 0.0, 0
    for o in pick:
        num, den, _ = check_tile(c, a, b, target, lidx, ridx, seeds, o, nsample, rng, factor)
        worst = max(worst, np.sqrt(num / den) if den > 0 else np.sqrt(num))
        nel += int(np.prod([min(nsample, e) for e in c.trange.tile_extent(c.trange.tile_index(o))]))
    return {"worst_rel": float(worst), "tiles": len(pick), "elements": nel}
