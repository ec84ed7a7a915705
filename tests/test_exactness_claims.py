"""CPU (-m "not gpu"): the float32 identities the kernels' refactorings rest on, checked
in IEEE round-to-nearest float32 arithmetic (numpy) on random and extreme operands.

1. max distributes over a common addend bit for bit: max(fl(a+s), fl(b+s)) == fl(max(a,b)+s).
2. Decisions as sign bits: for finite a <= b, sign(fl(a-b)) is set exactly when a != b, and fl(a-a) is +0;
   for any finite a, b, sign(fl(b-a)) is set exactly when a > b (the INSERTION decision zm > zi).
3. -FLT_MAX (semiring zero) stays -FLT_MAX under the penalties the recurrence adds to it.
"""
import numpy as np

F32_MAX = np.finfo(np.float32).max


def _operands(rng, n):
    """Scores as the lattice sees them: log-odds sums up to 1e5, gap penalties, -FLT_MAX, tiny values."""
    pools = [
        rng.uniform(-1e5, 1e5, n), rng.uniform(-20, 20, n), rng.uniform(-1e-30, 1e-30, n),
        np.full(n, -F32_MAX), np.full(n, -F32_MAX) + rng.uniform(-10, 0, n),
        rng.standard_normal(n) * 10.0 ** rng.uniform(-40, 38, n),
    ]
    x = np.concatenate(pools).astype(np.float32)
    rng.shuffle(x)
    return x[np.isfinite(x)]


def _sign(x):
    return (x.view(np.uint32) >> 31).astype(bool)


def test_max_distributes_over_a_common_addend():
    rng = np.random.RandomState(1)
    a, b = _operands(rng, 200000), _operands(rng, 200000)
    n = min(len(a), len(b))
    a, b = a[:n], b[:n]
    s = rng.uniform(-30, 30, n).astype(np.float32)
    with np.errstate(over="ignore"):
        lhs = np.maximum(a + s, b + s)
        rhs = np.maximum(a, b) + s
    assert np.array_equal(lhs.view(np.uint32), rhs.view(np.uint32))
    # three operands, as in X = max3((M+ng)+ng, D+gs, (I+gs)+ng) -> M(r+1,c+1) = X + subst
    c = _operands(rng, 200000)[:n]
    with np.errstate(over="ignore"):
        lhs3 = np.maximum(np.maximum(a + s, b + s), c + s)
        rhs3 = np.maximum(np.maximum(a, b), c) + s
    assert np.array_equal(lhs3.view(np.uint32), rhs3.view(np.uint32))


def test_decisions_are_sign_bits_of_differences():
    rng = np.random.RandomState(2)
    x, y = _operands(rng, 300000), _operands(rng, 300000)
    n = min(len(x), len(y))
    x, y = x[:n], y[:n]
    x[: n // 4] = y[: n // 4]  # plenty of exact ties
    # neighbours in float32 (differences of one ulp, including subnormal differences)
    with np.errstate(over="ignore"):
        x[n // 4: n // 2] = np.nextafter(y[n // 4: n // 2], np.float32(-np.inf))
    x = np.maximum(x, np.float32(-F32_MAX))  # scores never leave [-FLT_MAX, FLT_MAX]
    m = np.maximum(x, y)  # the maximum the kernel subtracts (X or Y): every operand is <= m
    with np.errstate(over="ignore"):
        for a in (x, y):
            d = a - m
            assert np.array_equal(_sign(d), a != m)          # bit set <=> "differs from the maximum"
            assert not _sign(d[a == m]).any()                # a - a is +0, never -0
        assert np.array_equal(_sign(y - x), x > y)           # plane 4: zm > zi <=> sign(zi - zm)
        assert not np.isnan(x - m).any()


def test_semiring_zero_is_absorbing_for_the_penalties():
    low = np.float32(-F32_MAX)
    for g, e in ((0.001, 5.0 / 6.0), (0.5, 0.99), (1e-6, 1e-3)):
        pen = np.array([np.log1p(-g), np.log1p(-e), np.log(g), np.log(e)], dtype=np.float32)
        for p in pen:
            assert np.float32(low + p) == low
            assert np.float32(np.float32(low + p) + p) == low
    # and differences against it stay finite or overflow to -inf with the sign set (never NaN)
    with np.errstate(over="ignore"):
        d = np.float32(low) - np.float32(1e5)
    assert np.signbit(d) and not np.isnan(d)
