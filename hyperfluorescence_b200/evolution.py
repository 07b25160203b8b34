"""Evolutionary optimisation of the host / TADF / emitter proportions.

The reference only states the goal ("optimize the device through the molecules proportions with
an evolutionnist algorithm", ``README.md:2``; ``Proportion`` ``molecule.py:403-418``); nothing is
implemented there, so the algorithm below is this package's own.  What it must
keep from the reference is the *fitness*: the internal quantum efficiency a ``Lattice`` reports,
``IQE = 100 * 2 * emission / injection`` (``reseau.py:595``), for a candidate composition.

Mapping to the GPU: a population of P candidates is ONE handle with P lattice realisations, each
generated with its own composition (``hf_generate_lattice_mixed``), carrying T trajectories each;
one ``hf_run`` launch steps all P*T trajectories and ``hf_read_realisation_tallies`` returns the
per-candidate emission / injection sums.  Across GPUs the population is sharded by candidate
(global realisation ids, so a candidate's fitness does not depend on the rank that evaluates it)
and the fitness vector is all-gathered.

A candidate is stored as integer dopant counts (n_tadf, n_fluo) out of the N = X*Y*Z sites, which
is what the reference's lattice builder reduces proportions to anyway (``reseau.py:159-161``);
``proportions()`` returns fractions that reproduce exactly those counts under ``int(N * p)``.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class GAConfig:
    population: int = 256
    trajectories_per_candidate: int = 10_000
    steps: int = 1_000                      # KMC events per trajectory per fitness evaluation
    elite: int = 8                          # candidates copied unchanged into the next generation
    tournament: int = 3
    crossover_rate: float = 0.7
    mutation_rate: float = 0.3
    mutation_scale: float = 0.05            # std-dev of the Gaussian mutation, in site fraction
    max_dopant_fraction: float = 0.6        # tadf + fluo <= this fraction of the INTERIOR sites


class Compositions:
    """Integer compositions of a lattice of ``dims`` and their validity rules."""

    def __init__(self, dims, max_dopant_fraction=0.6):
        self.dims = tuple(int(d) for d in dims)
        self.N = self.dims[0] * self.dims[1] * self.dims[2]
        self.interior = self.dims[0] * self.dims[1] * (self.dims[2] - 2)
        self.max_dopants = max(2, int(self.interior * max_dopant_fraction))

    def clip(self, counts):
        """Project (n_tadf, n_fluo) pairs onto the valid set: each >= 1 (``reseau.py:137``), host
        count of the interior >= 0 and the dopant cap."""
        c = np.maximum(np.rint(np.asarray(counts, dtype=np.float64)).astype(np.int64), 1)
        over = c.sum(axis=1) > self.max_dopants
        if over.any():
            scale = self.max_dopants / c[over].sum(axis=1, keepdims=True)
            c[over] = np.maximum((c[over] * scale).astype(np.int64), 1)
        return c

    def proportions(self, counts):
        """(host, tadf, fluo) fractions with int(N * p) == count for both dopants."""
        c = np.asarray(counts, dtype=np.int64)
        tadf = (c[:, 0] + 0.5) / self.N
        fluo = (c[:, 1] + 0.5) / self.N
        props = np.stack([1.0 - tadf - fluo, tadf, fluo], axis=1)
        assert np.array_equal((self.N * props[:, 1]).astype(np.int64), c[:, 0])
        assert np.array_equal((self.N * props[:, 2]).astype(np.int64), c[:, 1])
        return props

    def random(self, rng, n):
        frac = rng.dirichlet((6.0, 2.0, 1.0), size=n)            # mostly host, like the reference default
        return self.clip(np.stack([frac[:, 1], frac[:, 2]], axis=1) * self.interior)


def next_generation(counts, fitness, comps: Compositions, cfg: GAConfig, rng):
    """Elitism + tournament selection + blend crossover + Gaussian mutation on the count pairs."""
    counts = np.asarray(counts, dtype=np.int64)
    fitness = np.asarray(fitness, dtype=np.float64)
    n = len(counts)
    order = np.argsort(-fitness, kind="stable")
    out = [counts[i] for i in order[:cfg.elite]]
    def pick():
        idx = rng.integers(0, n, size=cfg.tournament)
        return counts[idx[np.argmax(fitness[idx])]].astype(np.float64)
    while len(out) < n:
        a, b = pick(), pick()
        if rng.random() < cfg.crossover_rate:
            w = rng.random()
            child = w * a + (1.0 - w) * b
        else:
            child = a
        if rng.random() < cfg.mutation_rate:
            child = child + rng.normal(0.0, cfg.mutation_scale * comps.interior, size=2)
        out.append(child)
    return comps.clip(np.array(out, dtype=np.float64))


class DeviceFitness:
    """IQE of a population of compositions on the GPU(s): the fitness kernel."""

    def __init__(self, dims, charges=1, electric_field=0.1, cfg: GAConfig | None = None, seed=0, device=0,
                 rank=0, world=1):
        self.dims, self.charges, self.field = tuple(dims), int(charges), float(electric_field)
        self.cfg = cfg or GAConfig()
        self.seed, self.device, self.rank, self.world = int(seed), int(device), int(rank), int(world)
        self.comps = Compositions(dims, self.cfg.max_dopant_fraction)
        if self.cfg.population % self.world:
            raise ValueError("world size must divide the population")
        self.local = self.cfg.population // self.world
        self.generation = 0
        self._sim = None
        self.last_tallies = None

    def _handle(self):
        if self._sim is None:
            from . import _native as nat
            t = self.cfg.trajectories_per_candidate
            self._sim = nat.Sim(self.dims, (0.84, 0.15, 0.01), self.field, self.charges, 1,
                                n_trajectories=self.local * t, n_realisations=self.local, seed=self.seed,
                                first_trajectory=self.rank * self.local * t,
                                first_realisation=self.rank * self.local, device=self.device)
        return self._sim

    def __call__(self, counts):
        """Fitness (IQE, percent) of every candidate of the population; all ranks get all values."""
        counts = np.asarray(counts, dtype=np.int64)
        if len(counts) != self.cfg.population:
            raise ValueError("population size mismatch")
        sim = self._handle()
        mine = counts[self.rank * self.local:(self.rank + 1) * self.local]
        sim.generate_lattice_mixed(self.comps.proportions(mine))
        sim.inject()
        sim.run(self.cfg.steps)
        t = sim.realisation_tallies()
        self.last_tallies = t
        local = 100.0 * 2.0 * t["emission"].astype(np.float64) / t["injection"].astype(np.float64)   # reseau.py:595
        self.generation += 1
        return all_gather_fitness(local, self.world)

    def close(self):
        if self._sim is not None:
            self._sim.close()
            self._sim = None


class ExcitonFitness(DeviceFitness):
    """EQE (percent) of a population of compositions from the exciton path: photons per decayed exciton
    times the outcoupling, for ``trajectories_per_candidate`` single-exciton trajectories of ``steps`` KMC
    events on each candidate's own lattice (``excitons.ExcitonEnsemble``'s model, DESIGN.md section 9 — this
    package's own: the reference has no exciton dynamics and therefore no EQE)."""

    def __init__(self, dims, parameters=None, cfg: GAConfig | None = None, seed=0, device=0, rank=0, world=1):
        super().__init__(dims, 1, 0.1, cfg, seed, device, rank, world)
        self.parameters = parameters or {}

    def __call__(self, counts):
        from . import _native as nat
        counts = np.asarray(counts, dtype=np.int64)
        if len(counts) != self.cfg.population:
            raise ValueError("population size mismatch")
        sim = self._handle()
        mine = counts[self.rank * self.local:(self.rank + 1) * self.local]
        sim.generate_lattice_mixed(self.comps.proportions(mine))
        params = nat.x_params_from_dict(self.parameters)
        sim.x_configure(params)
        sim.x_spawn()
        sim.x_run(self.cfg.steps)
        t = sim.x_realisation_tallies()
        self.last_tallies = t
        rad = (t["rad_host"] + t["rad_tadf"] + t["rad_fluo"]).astype(np.float64)
        decayed = rad + sum(t[k] for k in t if k.startswith("nonrad_")).astype(np.float64)
        local = 100.0 * params.outcoupling * rad / np.maximum(decayed, 1.0)
        self.generation += 1
        return all_gather_fitness(local, self.world)


class DeviceEQEFitness(DeviceFitness):
    """EQE (percent) of a population of compositions from the INTEGRATED device: carriers are injected, drift, meet
    and bind (the reference's First Reaction Method), and the exciton they form hops, crosses and decays in the same
    event loop (``Lattice(..., excitons=...)``), so the fitness sees WHERE recombination
    happens.  EQE = outcoupling x IQE with the reference's IQE = 100 * 2 * emission / injection (reseau.py:595).
    The exciton part is this package's own model (parity unpinned)."""

    def __init__(self, dims, charges=4, parameters=None, electric_field=0.1, cfg: GAConfig | None = None, seed=0,
                 device=0, rank=0, world=1):
        super().__init__(dims, charges, electric_field, cfg, seed, device, rank, world)
        self.parameters = parameters or {}

    def __call__(self, counts):
        from . import _native as nat
        counts = np.asarray(counts, dtype=np.int64)
        if len(counts) != self.cfg.population:
            raise ValueError("population size mismatch")
        sim = self._handle()
        mine = counts[self.rank * self.local:(self.rank + 1) * self.local]
        sim.generate_lattice_mixed(self.comps.proportions(mine))
        params = nat.x_params_from_dict(self.parameters)
        sim.x_configure(params)                                # tables and Boltzmann weights of the new lattices
        sim.p_excitons(2)
        sim.inject()
        sim.run(self.cfg.steps)
        t = sim.realisation_tallies()
        self.last_tallies = t
        local = params.outcoupling * 100.0 * 2.0 * t["emission"].astype(np.float64) / t["injection"].astype(np.float64)
        self.generation += 1
        return all_gather_fitness(local, self.world)


def all_gather_fitness(local, world):
    """Concatenate the per-rank fitness blocks (identity for one rank)."""
    local = np.ascontiguousarray(local, dtype=np.float64)
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(local)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return torch.cat(out).cpu().numpy()


class CompositionGA:
    """The optimiser: ``run(generations)`` returns the best composition found and its history."""

    def __init__(self, dims, fitness, cfg: GAConfig | None = None, seed=0):
        self.cfg = cfg or GAConfig()
        self.comps = Compositions(dims, self.cfg.max_dopant_fraction)
        self.fitness_fn = fitness
        self.rng = np.random.default_rng(seed)        # same seed on every rank => same population everywhere
        self.population = self.comps.random(self.rng, self.cfg.population)
        self.history = []

    def step(self):
        fit = np.asarray(self.fitness_fn(self.population), dtype=np.float64)
        best = int(np.argmax(fit))
        self.history.append(dict(best_fitness=float(fit[best]), mean_fitness=float(fit.mean()),
                                 best_counts=tuple(int(v) for v in self.population[best])))
        self.population = next_generation(self.population, fit, self.comps, self.cfg, self.rng)
        return fit

    def run(self, generations):
        for _ in range(generations):
            self.step()
        best = max(self.history, key=lambda h: h["best_fitness"])
        props = self.comps.proportions(np.array([best["best_counts"]]))[0]
        return dict(best_counts=best["best_counts"], best_proportions=tuple(float(p) for p in props),
                    best_fitness=best["best_fitness"], history=self.history)
