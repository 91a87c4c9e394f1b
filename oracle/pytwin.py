"""Pure-Python second restatement of the oracle's arithmetic (small cases only).

TEST INFRASTRUCTURE ONLY (same rules as pyoracle.py).  It exists so that the C oracle is
cross-checked by an independently written model: exact rational arithmetic with explicit
IEEE-754 rounding instead of the host FPU.  Follows the same reference lines as vs_oracle.c:
J/util/Distances.java:48-153, J/pq/PqEncoder.java:18-37, J/pq/PqTrainer.java:28-91,
java.util.Random (JDK core).
"""
from __future__ import annotations

import math
import struct
from fractions import Fraction

MASK48 = (1 << 48) - 1
MULT = 0x5DEECE66D


class JavaRandom:
    def __init__(self, seed: int):
        self.state = (seed ^ MULT) & MASK48

    def next(self, bits: int) -> int:
        self.state = (self.state * MULT + 0xB) & MASK48
        v = (self.state >> (48 - bits)) & 0xFFFFFFFF
        return v - (1 << 32) if v >= (1 << 31) else v

    def next_int(self, bound: int | None = None) -> int:
        if bound is None:
            return self.next(32)
        if bound <= 0:
            raise ValueError("bound must be positive")
        r = self.next(31)
        m = bound - 1
        if bound & m == 0:
            return (bound * r) >> 31
        u = r
        while True:
            r = u % bound
            t = (u - r + m) & 0xFFFFFFFF
            if t < (1 << 31):  # non-negative as a Java int
                return r
            u = self.next(31)

    def next_float(self) -> float:
        return self.next(24) / float(1 << 24)


def _round_f32(x: Fraction) -> float:
    """Round an exact rational to the nearest binary32 (ties to even); returns a Python float."""
    if x == 0:
        return 0.0
    sign = -1 if x < 0 else 1
    x = abs(x)
    e = x.numerator.bit_length() - x.denominator.bit_length()
    if Fraction(2) ** e > x:
        e -= 1
    e = max(e, -126)  # subnormal range shares the exponent of the smallest normal
    scaled = x / (Fraction(2) ** (e - 23))  # in [2^23, 2^24) for normals
    n = scaled.numerator // scaled.denominator
    rem = scaled - n
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and n & 1):
        n += 1
    val = Fraction(n) * (Fraction(2) ** (e - 23))
    if val >= Fraction(2) ** 128:
        return sign * math.inf
    return sign * float(val)


def f32(x: float) -> float:
    return struct.unpack("<f", struct.pack("<f", x))[0]


def fmaf(a: float, b: float, c: float) -> float:
    return _round_f32(Fraction(a) * Fraction(b) + Fraction(c))


def fsub(a: float, b: float) -> float:
    return _round_f32(Fraction(a) - Fraction(b))


def fadd(a: float, b: float) -> float:
    return _round_f32(Fraction(a) + Fraction(b))


def _reduce(acc):
    s = 0.0
    for v in acc:
        s = fadd(s, v)
    return s


def l2_squared(a, b, lanes: int = 16) -> float:
    n = len(a)
    ub = n - n % lanes
    acc = [0.0] * lanes
    for i in range(0, ub, lanes):
        for l in range(lanes):
            d = fsub(a[i + l], b[i + l])
            acc[l] = fmaf(d, d, acc[l])
    s = float(_reduce(acc))
    for i in range(ub, n):
        d = float(a[i]) - float(b[i])  # doubles; exact for fp32 inputs of similar scale
        s += d * d
    return s


def dot(a, b, lanes: int = 16) -> float:
    n = len(a)
    ub = n - n % lanes
    acc = [0.0] * lanes
    for i in range(0, ub, lanes):
        for l in range(lanes):
            acc[l] = fmaf(a[i + l], b[i + l], acc[l])
    s = float(_reduce(acc))
    for i in range(ub, n):
        s += float(a[i]) * float(b[i])
    return s


def norm(a, lanes: int = 16) -> float:
    return math.sqrt(dot(a, a, lanes))


def l2(a, b, lanes: int = 16) -> float:
    return math.sqrt(l2_squared(a, b, lanes))


def cosine(a, b, lanes: int = 16) -> float:
    n = norm(a, lanes) * norm(b, lanes)
    if n == 0.0:
        return 0.0
    return dot(a, b, lanes) / n


def pq_encode(centroids, v, lanes: int = 16):
    """centroids: nested [M][K][subDim] of Python floats holding fp32 values."""
    codes = []
    for s, book in enumerate(centroids):
        sub = len(book[0])
        x = v[s * sub:(s + 1) * sub]
        best, best_d = 0, math.inf
        for ci, c in enumerate(book):
            d = l2_squared(x, c, lanes)
            if d < best_d:
                best_d, best = d, ci
        codes.append(best & 0xFF)
    return codes


def pq_train(vectors, D: int, M: int, K: int, iterations: int, seed: int, lanes: int = 16):
    if M <= 0 or K <= 0 or D <= 0:
        raise ValueError("Invalid PQ params (m,k,dimension)")
    if D % M != 0:
        raise ValueError("dimension must be divisible by m")
    sub = D // M
    rnd = JavaRandom(seed)
    n = len(vectors)
    out = []
    for s in range(M):
        data = [list(v[s * sub:(s + 1) * sub]) for v in vectors]
        cent = [list(data[rnd.next_int(n)]) for _ in range(K)]
        for _ in range(iterations):
            assign = []
            for x in data:
                best, best_d = 0, math.inf
                for ci in range(K):
                    d = l2_squared(x, cent[ci], lanes)
                    if d < best_d:
                        best_d, best = d, ci
                assign.append(best)
            new_c = [[0.0] * sub for _ in range(K)]
            counts = [0] * K
            for x, a in zip(data, assign):
                for d in range(sub):
                    new_c[a][d] = fadd(new_c[a][d], x[d])
                counts[a] += 1
            for ci in range(K):
                if counts[ci] == 0:
                    new_c[ci] = list(data[rnd.next_int(n)])
                else:
                    for d in range(sub):
                        new_c[ci][d] = _round_f32(Fraction(new_c[ci][d]) / Fraction(f32(float(counts[ci]))))
            cent = new_c
        out.append(cent)
    return out
