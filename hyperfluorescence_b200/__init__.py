"""hyperfluorescence_b200 — B200-native kinetic-Monte-Carlo hot path of Pehnny/hyperfluorescence.

The Python surface mirrors the reference's modules (``reseau``, ``molecule``, ``event``,
``constants``, ``plot``, ``OLED``); the work happens in ``libhfkmc.so`` (hand-written sm_100a
CUDA, C-ABI in ``include/hfkmc.h``).  Importing the package does not need a GPU; constructing a
``Lattice`` does, and there is no CPU fallback.

Drop-in use of an unmodified reference script::

    PYTHONPATH=/path/to/hyperfluorescence_b200/dropin python OLED.py
"""
from .event import EVENTS, PARTICULES, Event, Point, Vector  # noqa: F401
from .molecule import EXCITON, Fluorescent, Host, Molecule, Proportion, TADF  # noqa: F401

__all__ = ["Lattice", "EVENTS", "PARTICULES", "Event", "Point", "Vector", "EXCITON", "Fluorescent", "Host",
           "Molecule", "Proportion", "TADF"]


def __getattr__(name):
    if name == "Lattice":          # lazy: keeps `import hyperfluorescence_b200` free of ctypes loading
        from .reseau import Lattice
        return Lattice
    raise AttributeError(name)
