"""Known-answer cases restated from the reference's own tests for the contraction path.

The reference keeps no golden files: every expectation is recomputed in-test from a closed
form. The closed forms are restated here once and used twice: to pin the CPU
oracle (tests/test_oracle_known_answers.py, no GPU) and as parity cases for the CUDA path
(tests/test_gpu_*.py). Citations are file:line under /root/reference.
"""
import numpy as np

# ---- tile permutation: tests/librett.cpp --------------------------------------------------
# (extent, perm, einsum-like description "b(<out index order>) = a(i,j,...)")
#   :475-505  rank 2  {1,0}           b(j,i)         = a(i,j)      A = 10
#   :507-537  rank 2  {1,0} non-sym   b(j,i)         = a(i,j)      A x B = 10 x 5
#   :539-592  rank 3  {1,2,0}         b(k,i,j)       = a(i,j,k)    10 x 5 x 2
#             rank 3  {1,0,2}         b(j,i,k)       = a(i,j,k)
#   :594-661  rank 4  {0,3,2,1}       b(i,l,k,j)     = a(i,j,k,l)  2 x 3 x 6 x 4
#             rank 4  {1,0,3,2}       b(j,i,l,k)     = a(i,j,k,l)
#   :663-745  rank 6  {0,3,2,1,5,4}   b(i,l,k,j,n,m) = a(i,j,k,l,m,n)  2 x 3 x 6 x 4 x 5 x 7
#             rank 6  {1,0,4,3,2,5}   b(j,i,m,l,k,n) = a(i,j,k,l,m,n)
LIBRETT_CASES = [
    ((10, 10), (1, 0), "ji"),
    ((10, 5), (1, 0), "ji"),
    ((10, 5, 2), (1, 2, 0), "kij"),
    ((10, 5, 2), (1, 0, 2), "jik"),
    ((2, 3, 6, 4), (0, 3, 2, 1), "ilkj"),
    ((2, 3, 6, 4), (1, 0, 3, 2), "jilk"),
    ((2, 3, 6, 4, 5, 7), (0, 3, 2, 1, 5, 4), "ilkjnm"),
    ((2, 3, 6, 4, 5, 7), (1, 0, 4, 3, 2, 5), "jimlkn"),
]


def librett_input(extent):
    """tile_a[iter] = iter in row-major order (tests/librett.cpp:607-620)."""
    return np.arange(int(np.prod(extent)), dtype=np.float64).reshape(extent)


def librett_check(b: np.ndarray, extent, out_order: str) -> bool:
    """BOOST_CHECK_EQUAL(tile_b(<out_order>), iter) for iter running over a's row-major order."""
    letters = "ijklmn"[:len(extent)]
    it = 0
    for idx in np.ndindex(*extent):
        env = dict(zip(letters, idx))
        if b[tuple(env[c] for c in out_order)] != it:
            return False
        it += 1
    return True


# ---- TiledRangeFixture: tests/range_fixture.h:67-132, tests/global_fixture.h:62-76 ----------
PRIMES = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29]
FIXTURE_BOUNDS = tuple(int(x) for x in np.concatenate([[0], np.cumsum(PRIMES[:5])]))  # 0,2,5,10,17,28
FIXTURE_RANK = 3  # TEST_DIM
SPARSE_FIXTURE_THRESHOLD = 0.001  # tests/sparse_shape_fixture.h:111


def fixture_norms(rng: np.random.Generator, tiles_shape, extents_per_dim, fill_percent: float, threshold: float):
    """make_norm_tensor (tests/sparse_shape_fixture.h:55-75): norm = sqrt(v^2 * volume), v in
    0..100, then a fraction overwritten by threshold*0.1. (world.rand() is replaced by a seeded
    numpy generator: MADWorld's RNG is not reproducible outside MADNESS.)"""
    norms = np.empty(tiles_shape, dtype=np.float32)
    for idx in np.ndindex(*tiles_shape):
        vol = float(np.prod([extents_per_dim[d][t] for d, t in enumerate(idx)]))
        v = float(rng.integers(0, 101))
        norms[idx] = np.float32(np.sqrt(v * v * vol))
    flat = norms.reshape(-1)
    n = int(float(flat.size) * (1.0 - fill_percent))
    for _ in range(n):
        flat[int(rng.integers(0, flat.size))] = np.float32(threshold * 0.1)
    return norms


# ---- ContractReduce: tests/tile_op_contract_reduce.cpp:109-220 ----------------------------------
# m = 18, k = 27, n = 36, factor 3, integer tiles (exact), all four op combinations.
CONTRACT_REDUCE_MNK = (18, 36, 27)
CONTRACT_REDUCE_FACTOR = 3.0


def int_tile(rng: np.random.Generator, shape):
    """Integer-valued tiles (the reference's `rand() % 101`-style fills, tests/expressions_fixture.h:
    47-266): all products and sums stay exactly representable in FP64, so results must be EXACT."""
    return rng.integers(0, 101, size=shape).astype(np.float64)
