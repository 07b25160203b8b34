"""Exciton trajectories on the device lattice: Förster and Dexter hops, ISC / RISC on the TADF
sensitizer, radiative and non-radiative decay, rejection-free (BKL) selection.

This fills the slots the reference leaves empty (``event.py:43-51`` ISC / Forster / decay codes,
``reseau.py:233-235`` the empty exciton event lists, ``reseau.py:526-550`` the ``...`` branches,
``reseau.py:533`` the routing comment "move ou unbound si Host ; ISC ou Forster si TADF ; Decay si
Fluorescent").  The reference has no rate laws or parameters for any of this, so the model and its
defaults are this package's own (specification: ``oracle/exciton_oracle.c``).
"""
from __future__ import annotations

import os
from math import prod

from . import _native as nat
from .event import EVENTS, PARTICULES  # noqa: F401  (event codes used by the trace)


class ExcitonEnsemble:
    """``trajectories`` independent single-exciton trajectories on ``realisations`` lattices of the
    given composition.  ``run(n)`` advances every trajectory by n KMC events; excitons that decay
    are replaced at once, so tallies accumulate over many exciton lifetimes.

    ``parameters`` is a dict overriding any field of ``hf_x_params`` (cutoff, forster_radius,
    dexter_prefactor, dexter_length, k_isc, k_risc, k_rad, k_nonrad, singlet_fraction,
    outcoupling); see ``default_parameters()``.
    """

    def __init__(self, dimension, proportions=(0.84, 0.15, 0.01), trajectories=1, realisations=1, parameters=None,
                 seed=None, device=0, first_trajectory=0, first_realisation=0, lattice=None):
        if not isinstance(dimension, tuple) or len(dimension) != 3:
            raise TypeError("dimension must be a 3-tuple")
        if dimension[-1] < 3:
            raise ValueError(f"Expected the last component of dimension to be > 2, got {dimension[-1]}.")
        if min(proportions) * prod(dimension) < 1:
            raise ValueError("Not enough molecules for current proportions and dimension.")
        if sum(proportions) != 1.:                                          # quirk Q1, reseau.py:143-145
            norm = sum(dimension)
            proportions = tuple(d / norm for d in dimension)
        self.seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed)
        self.parameters = nat.x_params_from_dict(parameters or {})
        self._sim = nat.Sim(dimension, proportions, 0.1, 1, 1, trajectories, realisations, self.seed,
                            first_trajectory, first_realisation, device)
        if lattice is None:
            self._sim.generate_lattice()
        else:
            for r in range(realisations):
                src = lattice[r] if isinstance(lattice, (list, tuple)) else lattice
                self._sim.upload_lattice(r, src["species"], src["homo"], src["lumo"], src["s1"], src["t1"])
        self._sim.x_configure(self.parameters)
        self._sim.x_spawn()

    @staticmethod
    def default_parameters() -> dict:
        return nat.x_params_to_dict(nat.x_default_params())

    def run(self, n_events: int) -> None:
        self._sim.x_run(n_events)
        self._sim.sync()

    def tallies(self) -> dict:
        return self._sim.x_tallies()

    def summary(self) -> dict:
        """Photons per exciton (IQE), EQE = outcoupling x IQE, and the emission / loss channel
        fractions over the excitons that have decayed so far."""
        t = self.tallies()
        rad = t["rad_host"] + t["rad_tadf"] + t["rad_fluo"]
        lost_s = t["nonrad_s_host"] + t["nonrad_s_tadf"] + t["nonrad_s_fluo"]
        lost_t = t["nonrad_t_host"] + t["nonrad_t_tadf"] + t["nonrad_t_fluo"]
        decayed = rad + lost_s + lost_t
        frac = (lambda v: v / decayed) if decayed else (lambda v: 0.0)
        return dict(excitons_decayed=decayed, iqe=frac(rad), eqe=self.parameters.outcoupling * frac(rad),
                    channels={k: frac(t[k]) for k in t if k.startswith(("rad_", "nonrad_"))},
                    events=t["events"], forster=t["forster"], dexter=t["dexter"], isc=t["isc"], risc=t["risc"])

    def close(self):
        self._sim.close()


class DenseExcitonEnsemble:
    """``devices`` independent devices, each holding ``excitons_per_device`` INTERACTING excitons on one
    of ``realisations`` lattices: a 2-bit-per-site occupancy mask, transfers towards occupied sites
    are annihilation channels (``annihilation[donor spin][acceptor spin]`` scales their rate, 0 =
    pure blocking; spin index 0 singlet, 1 triplet), a triplet hit by a triplet is promoted to a
    singlet with probability ``p_tta``.  Lost excitons are replaced at once on a random empty
    interior site, so the exciton density stays at excitons_per_device / sites.

    BASELINE configs[3] ("high-density excitons with occupancy blocking and exciton-exciton
    annihilation").  The reference has nothing of the kind — its only occupancy rule is same-type
    charge blocking (``reseau.py:385, 394``) — so the model is this package's own (DESIGN.md §10;
    specification: ``oracle/exciton_dense_oracle.c``)."""

    def __init__(self, dimension, proportions=(0.84, 0.15, 0.01), devices=1, excitons_per_device=32, realisations=1,
                 parameters=None, annihilation=((1.0, 1.0), (1.0, 1.0)), p_tta=0.25, seed=None, device=0,
                 first_device=0, first_realisation=0):
        if not isinstance(dimension, tuple) or len(dimension) != 3:
            raise TypeError("dimension must be a 3-tuple")
        if dimension[-1] < 3:
            raise ValueError(f"Expected the last component of dimension to be > 2, got {dimension[-1]}.")
        if not 1 <= excitons_per_device <= 1024 or 2 * excitons_per_device > dimension[0] * dimension[1] * (dimension[2] - 2):
            raise ValueError("excitons_per_device must be in [1, 1024] and at most half the interior sites")
        if sum(proportions) != 1.:                                          # quirk Q1, reseau.py:143-145
            norm = sum(dimension)
            proportions = tuple(d / norm for d in dimension)
        self.seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed)
        self.parameters = nat.x_params_from_dict(parameters or {})
        self.excitons_per_device = int(excitons_per_device)
        self._sim = nat.Sim(dimension, proportions, 0.1, 1, 1, devices, realisations, self.seed, first_device,
                            first_realisation, device)
        self._sim.generate_lattice()
        self._sim.x_configure(self.parameters)
        self._sim.xd_configure(self.excitons_per_device, annihilation, p_tta)
        self._sim.xd_spawn()

    def run(self, n_events: int) -> None:
        """n_events KMC events per device."""
        self._sim.xd_run(n_events)
        self._sim.sync()

    def tallies(self) -> dict:
        return self._sim.xd_tallies()

    def device(self, index=0) -> dict:
        return self._sim.xd_device(index)

    def summary(self) -> dict:
        t = self.tallies()
        rad = t["rad_host"] + t["rad_tadf"] + t["rad_fluo"]
        ann = t["ann_ss"] + t["ann_st"] + t["ann_ts"] + t["ann_tt"]
        lost = rad + ann + sum(t[k] for k in t if k.startswith("nonrad_"))
        frac = (lambda v: v / lost) if lost else (lambda v: 0.0)
        return dict(excitons_lost=lost, annihilated=ann, annihilation_fraction=frac(ann), iqe=frac(rad),
                    eqe=self.parameters.outcoupling * frac(rad), tta_up=t["tta_up"], events=t["events"],
                    channels={k: frac(t[k]) for k in t if k.startswith(("rad_", "nonrad_", "ann_"))})

    def close(self):
        self._sim.close()
