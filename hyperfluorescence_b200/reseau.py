"""``Lattice`` — drop-in for the reference's ``reseau.Lattice`` (``reseau.py:55-621``) whose
kinetic-Monte-Carlo loop runs in hand-written sm_100a CUDA kernels behind ``libhfkmc.so``.

Same constructor, same methods (``operations``, ``get_IQE``, ``get_particules_positions``),
same readable attributes (``_emission``, ``_injection``, ``_recombination``, ``_step``, ``_IQE``,
``_cache``, ``_electrons_locations`` ... — ``OLED.py:19-21`` reads private ones, so they are part
of the surface) and the same exceptions.  Everything else is different: the lattice is a
device-resident struct-of-arrays, the step loop is one kernel launch per ``operations`` call,
and one object can carry many independent trajectories (keyword-only extras below).

There is no CPU path: constructing a ``Lattice`` without an sm_100 GPU raises.
"""
from __future__ import annotations

import os
from collections import deque
from math import prod

import numpy as np

from . import _native as nat
from . import constants as cst  # noqa: F401  (same module-level name as the reference)
from .event import EVENTS, PARTICULES, Event, Point, Vector
from .molecule import (EXCITON, SPECIES_CLASSES, Fluorescent, Host, Molecule, Proportion, TADF,  # noqa: F401
                       molecule_from_device)


def born_von_karman(position: int, distance: int, size: int, axe: str) -> list[int]:
    """Neighbour coordinates along one axis, periodic in x and y, open in z, with the exact edge
    behaviour of ``reseau.py:190-205`` at any distance (quirk Q3: only distance 1 is symmetric)."""
    lower, upper = position - distance, position + distance
    if lower > -1 and upper < size:
        return list(range(lower, upper + 1))
    if lower < 0:
        out = list(range(distance + 1))
        return out + list(range(size - distance, size)) if axe != "z" else out
    out = list(range(position - distance, size))
    return out + list(range(distance)) if axe != "z" else out


class _Row:
    def __init__(self, lattice, z, y):
        self._l, self._z, self._y = lattice, z, y

    def __len__(self):
        return self._l._dimension.x

    def __getitem__(self, x):
        if isinstance(x, slice):
            return [self[i] for i in range(*x.indices(len(self)))]
        if x < 0:
            x += len(self)
        if not 0 <= x < len(self):
            raise IndexError("list index out of range")
        return self._l._molecule_at(x, self._y, self._z)

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class _Plane:
    def __init__(self, lattice, z):
        self._l, self._z = lattice, z

    def __len__(self):
        return self._l._dimension.y

    def __getitem__(self, y):
        if isinstance(y, slice):
            return [self[i] for i in range(*y.indices(len(self)))]
        if y < 0:
            y += len(self)
        if not 0 <= y < len(self):
            raise IndexError("list index out of range")
        return _Row(self._l, self._z, y)

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class _Grid:
    """``Lattice._grid[z][y][x]`` without materialising X*Y*Z Python objects."""

    def __init__(self, lattice):
        self._l = lattice

    def __len__(self):
        return self._l._dimension.z

    def __getitem__(self, z):
        if isinstance(z, slice):
            return [self[i] for i in range(*z.indices(len(self)))]
        if z < 0:
            z += len(self)
        if not 0 <= z < len(self):
            raise IndexError("list index out of range")
        return _Plane(self._l, z)

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class Lattice:
    """Hyperfluorescent-OLED lattice (host / TADF / fluorescent emitter) stepped by the First
    Reaction Method on the GPU.

    Lattice(dimension, proportions, electric_field=0.1, charges=10, charge_tranfer_distance=1,
            architecture=NotImplemented)                                   # reseau.py:109-111

    Keyword-only extras (none changes the behaviour of a positional reference-style call):
      seed            Philox key; default: $HF_SEED if set, else fresh OS entropy (the reference is unseeded too)
      trajectories    independent trajectories carried by this object (default 1).  Tallies and
                      IQE are then sums over all of them.
      realisations    independent lattice realisations shared by the trajectories (default 1)
      first_trajectory / first_realisation   global ids of the first local ones (multi-GPU sharding)
      device          CUDA device ordinal
      lattice         dict(species, homo, lumo[, s1, t1]) of flat z,y,x arrays to upload instead
                      of generating the lattice on the device (e.g. taken from a reference run)
      replay          uniforms replacing the trajectory stream (parity testing)
      excitons        None (default): the reference as it is — an exciton decays at tau = 0 on the site it formed on
                      (reseau.py:529-533).  True or a dict of exciton parameters (``_native.x_params_from_dict``:
                      cutoff, forster_radius, dexter_prefactor, k_isc, k_risc, k_rad, k_nonrad ...): the exciton
                      formed by a `bound` event hops, crosses and decays in the SAME First Reaction Method as the
                      carriers — the slots the reference reserves (`_move_exciton_events`, `_isc_events`,
                      `_decay_events`, reseau.py:233-235, 526-550).  Own model.
                      "instant": the default behaviour routed through those slots (parity gate).
    """

    def __init__(self, dimension: tuple[int, int, int], proportions: tuple[float, float, float],
                 electric_field: float = 10.**(-1), charges: int = 10, charge_tranfer_distance: int = 1,
                 architecture: str = NotImplemented, *, seed: int | None = None, trajectories: int = 1,
                 realisations: int = 1, first_trajectory: int = 0, first_realisation: int = 0, device: int = 0,
                 lattice: dict | None = None, replay=None, excitons=None) -> None:
        self._init_raises(dimension, proportions)
        self._lattice_parameters_creation(dimension, proportions, electric_field, charges)
        molecules = self._dimension.x * self._dimension.y
        if self._charges > molecules:                                              # reseau.py:212-213
            raise ValueError(f"Required {self._charges} charges but only {molecules} molecules available.")
        # any distance: the device builds _born_von_karman's table verbatim, seams and duplicates included (Q3)
        self._distance = charge_tranfer_distance
        if seed is None and os.environ.get("HF_SEED"):         # unmodified reference scripts cannot pass seed=
            seed = int(os.environ["HF_SEED"], 0)
        self._seed_value = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed)
        self._trajectories = int(trajectories)
        self._sim = nat.Sim(dimension, (self._proportions.host, self._proportions.tadf, self._proportions.fluo),
                            electric_field, charges, charge_tranfer_distance, trajectories, realisations,
                            self._seed_value, first_trajectory, first_realisation, device)
        if lattice is None:
            self._sim.generate_lattice()                                           # reseau.py:115
        else:
            for r in range(realisations):
                src = lattice[r] if isinstance(lattice, (list, tuple)) else lattice
                self._sim.upload_lattice(r, src["species"], src["homo"], src["lumo"], src.get("s1"), src.get("t1"))
        if replay is not None:
            self._sim.set_replay(replay)
        self._exciton_params = None
        if isinstance(excitons, str) and excitons == "instant":
            self._sim.p_excitons(1)
        elif excitons:
            self._exciton_params = nat.x_params_from_dict({} if excitons is True else dict(excitons))
            self._sim.x_configure(self._exciton_params)
            self._sim.p_excitons(2)
        self._sim.inject()                                                         # reseau.py:116-117
        self._IQE: float = 0.
        self._time: float = 0.                                                     # never advanced (Q9)
        self._lattice_cache = None
        self._grid = _Grid(self)

    # ---- constructor helpers (reseau.py:126-152) ------------------------------------------------
    def _init_raises(self, dimension, proportions) -> None:
        if not isinstance(dimension, tuple):
            raise TypeError(f"Expected type(dimension) to be tuple, got {type(dimension)}.")
        elif len(dimension) != 3:
            raise IndexError(f"Expected dimension length to be 3, got {len(dimension)}.")
        elif dimension[-1] < 3:
            raise ValueError(f"Expected the last component of dimension to be > 2, got {dimension[-1]}.")
        if len(proportions) != 3:
            raise IndexError(f"Expected proportions length to be 3, got {len(dimension)}.")
        if min(proportions) * prod(dimension) < 1:
            minimum = 1 / min(proportions)
            raise ValueError("Not enough molecules for current proportions and dimension.\n "
                             f"Expected at least {minimum}, got {prod(dimension)}.")

    def _lattice_parameters_creation(self, dimension, proportions, electric_field, charges) -> None:
        if sum(proportions) != 1.:                                                 # quirk Q1
            norm = sum(dimension)
            proportions = (dimension[0] / norm, dimension[1] / norm, dimension[2] / norm)
        self._dimension: Point = Point(*dimension)
        self._proportions: Proportion = Proportion(*proportions)
        self._electric_field: Vector = Vector(0, 0, electric_field)                # [eV/nm]
        self._lattice_constant: float = 1.                                         # [nm]
        self._charge_transfer_rate: float = 10.**13                                # [Hz]
        self._temperature: float = 300.                                            # [K]
        self._charges: int = charges

    # ---- the step loop ---------------------------------------------------------------------------
    def operations(self, stop: int) -> None:
        """``stop`` KMC steps (``reseau.py:580-595``) for every trajectory, in one kernel launch."""
        self._sim.run(stop)
        self._sim.sync()
        if self._trajectories == 1:
            status = self._sim.info(0).status
            if status == nat.TRAJ_ZERO_DIVISION:         # swallowed after dumping the cache (587-589)
                print(self._cache)
                return
            if status == nat.TRAJ_NO_EVENTS:             # 590-591: IQE is NOT refreshed
                return
            self._raise_for_status(status)
        else:
            if self._sim.tallies()["stopped"]:
                for t in range(self._trajectories):
                    self._raise_for_status(self._sim.info(t).status)
        t = self._sim.tallies()
        self._IQE = 100. * 2. * float(t["emission"]) / float(t["injection"])       # reseau.py:595

    @staticmethod
    def _raise_for_status(status):
        if status == nat.TRAJ_VALUE_ERROR:
            raise ValueError("min() arg is an empty sequence, or list.remove(x): x not in list "
                             "(no free neighbour, or carrier flags out of step with the position lists)")
        if status == nat.TRAJ_REPLAY_EXHAUSTED:
            raise RuntimeError("replay buffer exhausted")
        if status == nat.TRAJ_CAPACITY:
            raise RuntimeError("internal flag-set capacity exceeded")

    # ---- getters -----------------------------------------------------------------------------------
    def get_IQE(self) -> float:
        return self._IQE

    def get_particules_positions(self):
        """Three lists of ``(Point, molecule class)``: electrons, holes, excitons (``reseau.py:617-621``)."""
        return tuple([(p, self._get_molecule_type(p)) for p in locs]
                     for locs in (self._electrons_locations, self._holes_locations, self._excitons_locations))

    def _point(self, site: int) -> Point:
        X, Y = self._dimension.x, self._dimension.y
        return Point(int(site) % X, (int(site) // X) % Y, int(site) // (X * Y))

    def _site(self, p: Point) -> int:
        return (p.z * self._dimension.y + p.y) * self._dimension.x + p.x

    def _positions(self, which, trajectory=0):
        return [self._point(s) for s in self._sim.positions(trajectory, which)]

    @property
    def _electrons_locations(self):
        return self._positions(0)

    @property
    def _holes_locations(self):
        return self._positions(1)

    @property
    def _excitons_locations(self):
        return self._positions(2)

    def _tally(self, name):
        return self._sim.tallies()[name]

    _emission = property(lambda self: self._tally("emission"))
    _recombination = property(lambda self: self._tally("recombination"))
    _injection = property(lambda self: self._tally("injection"))
    _step = property(lambda self: self._tally("steps"))

    def _make_event(self, rec) -> Event:
        return Event(self._point(rec["initial"]), self._point(rec["final"]), float(rec["tau"]),
                     int(rec["kind"]), int(rec["particule"]))

    @property
    def _cache(self):
        """The last 10 fired events of trajectory 0 (``reseau.py:124, 489-490``)."""
        recs = self._sim.last_events(0)
        return deque([None] * (10 - len(recs)) + [self._make_event(r) for r in recs], 10)

    def _move_events(self, which):
        ini, fin, tau = self._sim.move_events(0, which)
        part = PARTICULES["electron"] if which == 0 else PARTICULES["hole"]
        return [Event(self._point(i), self._point(f), float(t), EVENTS["move"], part) for i, f, t in zip(ini, fin, tau)]

    _move_electron_events = property(lambda self: self._move_events(0))
    _move_hole_events = property(lambda self: self._move_events(1))

    def _special_events(self, kind):
        info = self._sim.info(0)
        if info.special_kind != kind:
            return []
        p = self._point(info.special_site)
        return [Event(p, p, 0., kind, info.special_particule)]

    def _exciton_events(self, kinds):
        """Pending events of the exciton slots whose kind is in ``kinds`` (integrated excitons), in slot order."""
        x = self._sim.excitons(0)
        return [Event(self._point(s), self._point(f), float(t), int(k), PARTICULES["exciton"])
                for s, f, t, k in zip(x["site"], x["final"], x["tau"], x["kind"]) if int(k) in kinds]

    def _decay_event_list(self):
        if getattr(self._sim, "exciton_mode", 0):
            return self._exciton_events((EVENTS["decay"],))
        return self._special_events(EVENTS["decay"])

    _binding_events = property(lambda self: self._special_events(EVENTS["bound"]))
    _decay_events = property(_decay_event_list)
    _capture_events = property(lambda self: self._special_events(EVENTS["capture"]))
    # reseau.py:233-235: empty in the reference; filled when the excitons live in the device loop (excitons=...)
    _move_exciton_events = property(lambda self: self._exciton_events((EVENTS["move"], EVENTS["Forster"]))
                                    if getattr(self._sim, "exciton_mode", 0) else [])
    _isc_events = property(lambda self: self._exciton_events((EVENTS["ISC"],))
                           if getattr(self._sim, "exciton_mode", 0) else [])

    def get_EQE(self, outcoupling: float | None = None) -> float:
        """External quantum efficiency in percent: IQE (reseau.py:595) times the outcoupling efficiency (the
        exciton parameters' ``outcoupling``, 0.2 by default).  Not in the reference, which stops at the IQE."""
        if outcoupling is None:
            outcoupling = self._exciton_params.outcoupling if self._exciton_params is not None else 0.2
        return float(outcoupling) * self._IQE

    def _get_all_events(self):
        return (self._move_electron_events + self._move_hole_events + self._move_exciton_events + self._isc_events
                + self._decay_events + self._binding_events + self._capture_events)   # reseau.py:608-612

    # ---- lattice read-back -----------------------------------------------------------------------
    def _lattice_arrays(self):
        if self._lattice_cache is None:
            self._lattice_cache = self._sim.download_lattice(0)
        return self._lattice_cache

    def _neighbourhood(self, position: Point, distance: int = 1) -> list[Point]:
        d = self._dimension
        xs = born_von_karman(position.x, distance, d.x, "x")
        ys = born_von_karman(position.y, distance, d.y, "y")
        zs = born_von_karman(position.z, distance, d.z, "z")
        me = (position.x, position.y, position.z)
        return [Point(x, y, z) for x in xs for y in ys for z in zs if (x, y, z) != me]   # reseau.py:184-188

    def _molecule_at(self, x, y, z):
        a = self._lattice_arrays()
        p = Point(x, y, z)
        s = self._site(p)
        info = self._sim.info(0)
        exciton = info.exciton_state if (info.special_kind == EVENTS["decay"] and info.special_site == s) else 0
        if getattr(self._sim, "exciton_mode", 0):
            xs = self._sim.excitons(0)
            here = [int(sp) for st, sp in zip(xs["site"], xs["spin"]) if int(st) == s]
            exciton = here[0] if here else 0
        return molecule_from_device(int(a["species"][s]), p, self._neighbourhood(p, self._distance),
                                    float(a["homo"][s]), float(a["lumo"][s]), float(a["s1"][s]), float(a["t1"][s]),
                                    bool(s in self._sim.flags(0, 0)), bool(s in self._sim.flags(0, 1)), int(exciton))

    def _get_molecule(self, position: Point):
        return self._molecule_at(position.x, position.y, position.z)

    def _get_molecule_type(self, position: Point) -> type:
        return SPECIES_CLASSES[int(self._lattice_arrays()["species"][self._site(position)])]

    # ---- batch extras ------------------------------------------------------------------------------
    def tallies(self) -> dict:
        """All box totals (sums over this object's trajectories), including per-kind event counts
        and per-species recombination / emission channels."""
        return self._sim.tallies()

    def close(self) -> None:
        self._sim.close()
