"""numpy restatement of tadev_fill_uniform_f64's counter RNG (permute.cu: splitmix64 keyed by
(seed, global element offset)) so that tests can regenerate device-filled tiles on the host."""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return x ^ (x >> np.uint64(31))


def fill_uniform(n: int, seed: int, offset: int = 0) -> np.ndarray:
    with np.errstate(over="ignore"):
        base = (np.uint64(seed) * np.uint64(0xD1342543DE82EF95)) & _M
        idx = (np.arange(n, dtype=np.uint64) + np.uint64(offset)) & _M
        r = _splitmix64((base + idx) & _M)
    return (r >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0


def tile_fill(tile_ordinal: int, nelems: int, seed: int) -> np.ndarray:
    """DistArray.fill_random: offset = tile_ordinal << 32."""
    return fill_uniform(nelems, seed, tile_ordinal << 32)
