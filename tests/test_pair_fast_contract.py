"""The arithmetic claim behind the LOCAL kernel's one-pair fast path (kernels.cuh, `e.pair_fast`), checked with
exact rational arithmetic on the CPU:

  general:  C = (double) acc * 2^-40 per group;  d = fl(C_a - C_b);  d2 = fma(dz, dz, fma(dy, dy, fl(dx dx)));
            u = fl(d2 - fl(r r));  active <=> fl(k u) >= 0
  fast:     n = acc_a - acc_b (int64);  f = (double) n;  D = fma(fz, fz, fma(fy, fy, fl(fx fx)));
            active <=> (k > 0 ? D >= fl(fl(r r) 2^80) : D <= fl(fl(r r) 2^80))

are the same predicate whenever |acc_a|, |acc_b| < 2^53 -- including the ties, where d2 == r r exactly."""
from fractions import Fraction

import numpy as np


def fl(x: Fraction) -> float:
    return float(x)                      # correctly rounded (round-half-even) conversion of an exact rational


def fma(a: float, b: float, c: float) -> float:
    return fl(Fraction(a) * Fraction(b) + Fraction(c))


def general(acc_a, acc_b, r, k):
    d = [fl(Fraction(float(a)) * Fraction(1, 2 ** 40) - Fraction(float(b)) * Fraction(1, 2 ** 40)) for a, b in zip(acc_a, acc_b)]
    d2 = fma(d[2], d[2], fma(d[1], d[1], fl(Fraction(d[0]) * Fraction(d[0]))))
    u = fl(Fraction(d2) - Fraction(fl(Fraction(r) * Fraction(r))))
    return fl(Fraction(k) * Fraction(u)) >= 0.0


def fast(acc_a, acc_b, r, k):
    f = [float(a - b) for a, b in zip(acc_a, acc_b)]        # int -> double: round to nearest even, like I2F.F64.S64
    D = fma(f[2], f[2], fma(f[1], f[1], fl(Fraction(f[0]) * Fraction(f[0]))))
    t80 = fl(Fraction(fl(Fraction(r) * Fraction(r))) * 2 ** 80)
    return D >= t80 if k > 0 else D <= t80


def test_random_sums_and_radii():
    rng = np.random.default_rng(20261020)
    for _ in range(4000):
        scale = int(rng.choice([30, 41, 46, 52]))                       # |acc| up to 2^52
        acc_a = [int(rng.integers(-2 ** scale, 2 ** scale)) for _ in range(3)]
        near = rng.random() < 0.7                                       # mostly pairs a fraction of a nm apart
        acc_b = [a + int(rng.integers(-2 ** 39, 2 ** 39)) for a in acc_a] if near else [int(rng.integers(-2 ** scale, 2 ** scale)) for _ in range(3)]
        acc_b = [max(-2 ** 53 + 1, min(2 ** 53 - 1, b)) for b in acc_b]
        dist = float(np.sqrt(sum(((a - b) * 2.0 ** -40) ** 2 for a, b in zip(acc_a, acc_b))))
        r = float(dist * (1.0 + rng.choice([0.0, 1e-16, -1e-16, 3e-16, 1e-9, -1e-9, 0.3, -0.3]))) or 0.25
        for k in (1.0, -1.0, 2.5, -1e-3):
            assert general(acc_a, acc_b, r, k) == fast(acc_a, acc_b, r, k), (acc_a, acc_b, r, k)


def test_exact_ties():
    """d2 == r r exactly (difference along one axis, a power-of-two multiple of a small integer): u = 0, k u = +-0 >= 0, active for
    both signs of k -- and the comparison form says so too."""
    for m, e in ((3, 38), (5, 36), (1, 40), (7, 20), (12345, 25)):
        n = m << e
        acc_a, acc_b = [n + 17, -5, 9], [17, -5, 9]
        r = m * 2.0 ** (e - 40)
        for k in (1.0, -1.0):
            assert general(acc_a, acc_b, r, k) and fast(acc_a, acc_b, r, k)
        # one ulp inside / outside flips exactly one side
        for r2, inside in ((np.nextafter(r, 0.0), False), (np.nextafter(r, np.inf), True)):
            assert general(acc_a, acc_b, float(r2), -1.0) == fast(acc_a, acc_b, float(r2), -1.0) == inside
            assert general(acc_a, acc_b, float(r2), 1.0) == fast(acc_a, acc_b, float(r2), 1.0) == (not inside)
