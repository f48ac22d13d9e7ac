"""CPU test of the sampled-element parity checker (tests/sampled_parity.py) itself: on host-built
"arrays" whose tiles follow the device RNG convention, the sampled check must accept the exact einsum
result and reject a result with one wrong 2x2 block."""
import numpy as np

from tests import sampled_parity as SP
from tests import util_rng
from tiledarray_b200.tiledarray import TiledRange, TiledRange1


class _World:
    rank = 0


class FakeArray:
    def __init__(self, trange, seed=None, zero=()):
        self.trange, self.seed, self.zero, self.world = trange, seed, set(zero), _World()
        self.tiles = {}

    def is_zero(self, o):
        return o in self.zero

    def full(self):
        out = np.zeros(self.trange.elements_shape)
        for o in range(self.trange.ntiles):
            if o in self.zero:
                continue
            idx = self.trange.tile_index(o)
            ext = self.trange.tile_extent(idx)
            out[self.trange.tile_slices(idx)] = util_rng.tile_fill(o, int(np.prod(ext)), self.seed).reshape(ext)
        return out

    def find(self, o):
        return self.tiles[o]


def _result(tr, full):
    c = FakeArray(tr)
    for o in range(tr.ntiles):
        c.tiles[o] = full[tr.tile_slices(tr.tile_index(o))].copy()
    return c


def test_fill_uniform_at_matches_contiguous_fill():
    offs = np.arange(1000, dtype=np.uint64) + np.uint64(5 << 32)
    assert np.array_equal(SP.fill_uniform_at(offs, 9), util_rng.fill_uniform(1000, 9, 5 << 32))


def test_sampled_check_matrix_product_sparse():
    t = TiledRange1(0, 5, 12, 20)
    tr = TiledRange([t, t])
    a, b = FakeArray(tr, 3, zero={1, 5}), FakeArray(tr, 4, zero={0})
    C = a.full() @ b.full()
    c = _result(tr, C)
    r = SP.check_local_tiles(c, a, b, "m,n", "m,k", "k,n", (3, 4), ntiles=9, nsample=3)
    assert r["tiles"] == 9 and r["worst_rel"] < 1e-13
    # a wrong block anywhere in a tile is caught when every position is sampled
    c.tiles[4][2:4, 1:3] += 1e-6
    r = SP.check_local_tiles(c, a, b, "m,n", "m,k", "k,n", (3, 4), ntiles=9, nsample=8)
    assert r["worst_rel"] > 1e-9


def test_sampled_check_permuted_4index():
    s, v = TiledRange1(0, 2, 5), TiledRange1(0, 3, 7)
    trA, trB, trC = TiledRange([s, s, v, v]), TiledRange([s, v, s, v]), TiledRange([s, v, s, v])
    a, b = FakeArray(trA, 9), FakeArray(trB, 10)
    C = np.einsum("ikac,jckb->iajb", a.full(), b.full())
    c = _result(trC, C)
    r = SP.check_local_tiles(c, a, b, "i,a,j,b", "i,k,a,c", "j,c,k,b", (9, 10), ntiles=5, nsample=2)
    assert r["worst_rel"] < 1e-13 and r["elements"] > 0
    r2 = SP.check_local_tiles(c, a, b, "i,a,j,b", "i,k,a,c", "j,c,k,b", (9, 10), ntiles=5, nsample=2, factor=2.0)
    assert r2["worst_rel"] > 0.1
