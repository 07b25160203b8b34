"""Event and geometry types of the reference's ``event.py`` (codes 43-57, ``Point`` 60-91,
``Vector`` 94-111, ``Event`` 114-173), kept as the Python-level data model of the façade.

On the device an event is ``(initial site, final site, tau, kind, particule)`` in
struct-of-arrays form (``include/hfkmc.h``: ``hf_event``); these classes are what
``Lattice._cache`` and ``get_particules_positions`` hand back to user code.
"""
from __future__ import annotations

from dataclasses import dataclass
from math import sqrt

# event.py:43-51 — the integer codes are part of the C-ABI (HF_EV_*)
EVENTS: dict[str, int] = {"move": 1, "bound": 2, "ISC": 3, "Forster": 4, "decay": 5, "unbound": 6, "capture": 7}
# event.py:53-57 (HF_P_*)
PARTICULES: dict[str, int] = {"electron": 1, "hole": 2, "exciton": 3}

_NUMBER = (float, int)


def _bad_operand(other):
    return TypeError(f"other must be of type Vector, float or int, got {type(other)}")


@dataclass
class Point:
    """A grid position.  Adding or subtracting points (or a number) yields a :class:`Vector`."""
    x: int
    y: int
    z: int

    def _combine(self, other, sign):
        if isinstance(other, Point):
            return Vector(self.x + sign * other.x, self.y + sign * other.y, self.z + sign * other.z)
        if isinstance(other, _NUMBER):
            return Vector(self.x + sign * other, self.y + sign * other, self.z + sign * other)
        raise _bad_operand(other)

    def __add__(self, other):
        return self._combine(other, 1)

    def __sub__(self, other):
        return self._combine(other, -1)


@dataclass
class Vector(Point):
    """A displacement.  ``v * w`` is the dot product, ``v * number`` a scaled vector."""
    x: float
    y: float
    z: float

    def __mul__(self, other):
        if isinstance(other, Point):
            return self.x * other.x + self.y * other.y + self.z * other.z
        if isinstance(other, _NUMBER):
            return Vector(self.x * other, self.y * other, self.z * other)
        raise _bad_operand(other)

    def __rmul__(self, other):
        return self * other

    def norm(self) -> float:
        return sqrt(self.x**2 + self.y**2 + self.z**2)


def _as_event(other):
    if not isinstance(other, Event):
        raise TypeError(f"other must be of type event, got {type(other)}")
    return other


@dataclass(eq=False)
class Event:
    """One pending or fired KMC event.

    Ordering compares ``tau`` only.  Equality is the reference's collision test
    (``event.py:141-147``): same kind and particle, and — for ``move`` / ``Forster`` — same
    initial *or* same final site; for the other kinds all four sites equal.  This is the rule
    the device applies when it drops pending events (quirk Q4, ``hf_remove_events``).
    """
    initial: Point
    final: Point
    tau: float
    kind: int
    particule: int = 0

    def __eq__(self, other) -> bool:
        other = _as_event(other)
        if self.kind != other.kind or self.particule != other.particule:
            return False
        if self.kind in (EVENTS["move"], EVENTS["Forster"]):
            return self.initial == other.initial or self.final == other.final
        return self.initial == other.initial == self.final == other.final

    def __ne__(self, other) -> bool:
        return not self == other

    def __lt__(self, other) -> bool:
        return self.tau < _as_event(other).tau

    def __gt__(self, other) -> bool:
        return self.tau > _as_event(other).tau

    def __le__(self, other) -> bool:
        return self.tau <= _as_event(other).tau

    def __ge__(self, other) -> bool:
        return self.tau >= _as_event(other).tau
