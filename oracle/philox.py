"""Philox4x32-10 counter-based RNG and the uniform-draw contract (TEST INFRASTRUCTURE).

This file is part of the parity oracle: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may import it.  The product path (``hyperfluorescence_b200``)
never does; the same contract is implemented on the device in
``hyperfluorescence_b200/csrc/hf_philox.cuh`` and in plain C in ``oracle/frm_oracle.c``.

The reference (``/root/reference``) has no counter-based RNG: it uses unseeded MT19937
instances, one per ``Lattice`` (``reseau.py:113``) and one per molecule (``molecule.py:72``).
Parity is defined by *injecting* uniforms through ``random.Random.random()``;
this module defines which uniform every consumer receives.

Draw contract
-------------
``uniform(seed, stream, ident, d)`` is the d-th uniform of stream ``stream`` of object ``ident``:

    key     = (seed & 0xffffffff, seed >> 32)
    b       = d >> 1                                          one Philox block yields two uniforms
    counter = (b & 0xffffffff, b >> 32, ident, stream)
    w       = philox4x32_10(counter, key)
    (lo,hi) = (w[0], w[1]) if d is even else (w[2], w[3])
    u       = (((hi << 32) | lo) >> 11) * 2**-53              in [0, 1), 53-bit like CPython's random()

Streams:

    STREAM_TRAJ    = 0  ident = trajectory   Lattice._seed after lattice creation: charge
                                             injection samples (reseau.py:219,225), one uniform
                                             per unblocked candidate hop (reseau.py:284,316),
                                             re-injection choices (reseau.py:455,467)
    STREAM_SPIN    = 1  ident = trajectory   k-th exciton formed in the trajectory
                                             (molecule.py:92, the molecule-private draw)
    STREAM_SPECIES = 2  ident = realisation  Lattice._seed during the species shuffle
                                             (reseau.py:164-168)
    STREAM_ENERGY  = 3  ident = realisation  d = 4*site + j, the j-th uniform consumed by the
                                             site's four gauss() calls (molecule.py:180-183)
"""
from __future__ import annotations

import random

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = 0xFFFFFFFF

STREAM_TRAJ = 0
STREAM_SPIN = 1
STREAM_SPECIES = 2
STREAM_ENERGY = 3

TWO_M53 = 2.0 ** -53


def philox4x32_10(counter, key):
    """Ten rounds of Philox-4x32 (Salmon et al., SC'11); returns four 32-bit words."""
    c0, c1, c2, c3 = counter
    k0, k1 = key
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = (
            ((p1 >> 32) ^ c1 ^ k0) & MASK32,
            p1 & MASK32,
            ((p0 >> 32) ^ c3 ^ k1) & MASK32,
            p0 & MASK32,
        )
        k0 = (k0 + W0) & MASK32
        k1 = (k1 + W1) & MASK32
    return c0, c1, c2, c3


def mantissa(seed: int, stream: int, ident: int, d: int) -> int:
    """The 53-bit integer m with uniform = m * 2**-53."""
    b = d >> 1
    w = philox4x32_10((b & MASK32, (b >> 32) & MASK32, ident & MASK32, stream & MASK32),
                      (seed & MASK32, (seed >> 32) & MASK32))
    lo, hi = (w[0], w[1]) if (d & 1) == 0 else (w[2], w[3])
    return ((hi << 32) | lo) >> 11


def uniform(seed: int, stream: int, ident: int, d: int) -> float:
    return mantissa(seed, stream, ident, d) * TWO_M53


class UniformStream:
    """A sequential reader of one (seed, stream, ident) Philox stream, or of a replay buffer."""

    def __init__(self, seed: int = 0, stream: int = 0, ident: int = 0, replay=None):
        self.seed, self.stream, self.ident = seed, stream, ident
        self.replay = replay
        self.count = 0

    def next(self) -> float:
        d = self.count
        self.count += 1
        if self.replay is not None:
            return float(self.replay[d])
        return uniform(self.seed, self.stream, self.ident, d)


class StreamRandom(random.Random):
    """``random.Random`` whose ``random()`` replays a :class:`UniformStream`.

    Only ``random()`` is overridden, so CPython routes ``_randbelow`` through
    ``_randbelow_without_getrandbits`` and ``sample``/``choice``/``gauss`` all consume the
    same uniform stream (SURVEY.md §8c "RNG draw contract").
    """

    def __init__(self, source: UniformStream):
        self.source = source
        super().__init__(0)

    def random(self) -> float:  # noqa: D401
        return self.source.next()

    def seed(self, *args, **kwargs):  # keep gauss_next reset, ignore the seed value
        super().seed(0)
