"""R-compatible Mersenne-Twister `set.seed()` + `sample(n)`.

Host-side helper of the Python mirror of the R glue.  The engine never draws random numbers: the R
glue runs `set.seed(seeds[i]); sample(nSamples)` itself (reference `R/calculatePvalues.R:315-318`)
and hands the `[B x n]` index matrix to the native library.  This module exists so the Python
tests / benchmarks can draw *the same* permutations R would, which is what lets the golden
`genomeRegions` null distribution be replayed without an R interpreter.

Follows R's `src/main/RNG.c` (`set.seed` scrambling, `MT_genrand`, `fixup`) and
`src/main/unique.c`/`random.c` (`do_sample` without replacement, `R_unif_index`).
"""
from __future__ import annotations

import math

import numpy as np

_N, _M = 624, 397
_MATRIX_A, _UPPER, _LOWER = 0x9908B0DF, 0x80000000, 0x7FFFFFFF
_I2_32M1 = 2.328306437080797e-10


class RRandom:
    """R's default RNG (Mersenne-Twister, inversion) restricted to unif_rand()/sample()."""

    def __init__(self, seed: int, sample_kind: str = "Rejection"):
        if sample_kind not in ("Rejection", "Rounding"):
            raise ValueError("sample_kind must be 'Rejection' (R >= 3.6) or 'Rounding' (R < 3.6)")
        self.sample_kind = sample_kind
        # RNG_Init: initial scrambling, then fill i_seed[0..624] with the LCG
        s = seed & 0xFFFFFFFF
        for _ in range(50):
            s = (69069 * s + 1) & 0xFFFFFFFF
        iseed = []
        for _ in range(_N + 1):
            s = (69069 * s + 1) & 0xFFFFFFFF
            iseed.append(s)
        # FixupSeeds: dummy[0] = mti = N  -> state must be regenerated on first draw
        self.mt = iseed[1:]
        self.mti = _N

    def _genrand(self) -> int:
        mt = self.mt
        if self.mti >= _N:
            for kk in range(_N - _M):
                y = (mt[kk] & _UPPER) | (mt[kk + 1] & _LOWER)
                mt[kk] = mt[kk + _M] ^ (y >> 1) ^ (_MATRIX_A if y & 1 else 0)
            for kk in range(_N - _M, _N - 1):
                y = (mt[kk] & _UPPER) | (mt[kk + 1] & _LOWER)
                mt[kk] = mt[kk + (_M - _N)] ^ (y >> 1) ^ (_MATRIX_A if y & 1 else 0)
            y = (mt[_N - 1] & _UPPER) | (mt[0] & _LOWER)
            mt[_N - 1] = mt[_M - 1] ^ (y >> 1) ^ (_MATRIX_A if y & 1 else 0)
            self.mti = 0
        y = mt[self.mti]
        self.mti += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def unif_rand(self) -> float:
        v = self._genrand() * 2.3283064365386963e-10
        # fixup(): keep strictly inside (0, 1)
        if v <= 0.0:
            return 0.5 * _I2_32M1
        if 1.0 - v <= 0.0:
            return 1.0 - 0.5 * _I2_32M1
        return v

    def _rbits(self, bits: int) -> float:
        v = 0
        n = 0
        while n <= bits:
            v1 = int(math.floor(self.unif_rand() * 65536))
            v = 65536 * v + v1
            n += 16
        if bits < 64:
            v &= (1 << bits) - 1
        return float(v)

    def unif_index(self, dn: float) -> float:
        if self.sample_kind == "Rounding":
            return math.floor(dn * self.unif_rand())
        if dn <= 0:
            return 0.0
        bits = int(math.ceil(math.log2(dn)))
        while True:
            dv = self._rbits(bits)
            if dv < dn:
                return dv

    def sample(self, n: int) -> np.ndarray:
        """`sample(n)` / `sample(seq_len(n))`: a permutation of 1..n, returned 0-based."""
        x = list(range(n))
        out = np.empty(n, dtype=np.int32)
        m = n
        for i in range(n):
            j = int(self.unif_index(float(m)))
            out[i] = x[j]
            m -= 1
            x[j] = x[m]
        return out


def permutation_indices(seeds, n: int, sample_kind: str = "Rejection") -> np.ndarray:
    """[B x n] int32, 0-based: row i == `set.seed(seeds[i]); sample(n) - 1`.

    (Reference `R/calculatePvalues.R:315-322`.)  A seed of None / NA is not supported here because
    it would continue whatever RNG stream the R session had.
    """
    out = np.empty((len(seeds), n), dtype=np.int32)
    for i, s in enumerate(seeds):
        out[i] = RRandom(int(s), sample_kind).sample(n)
    return out
