#!/usr/bin/env python
"""Where a GA fitness sweep (256 candidates x 1e4 trajectories, 10x10x5, charges 4) spends its time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hyperfluorescence_b200 import _native as nat
from hyperfluorescence_b200.evolution import Compositions
dims, P, T = (10, 10, 5), 256, 10_000
sim = nat.Sim(dims, (0.84, 0.15, 0.01), 0.1, 4, 1, n_trajectories=P * T, n_realisations=P, seed=1)
comps = Compositions(dims)
pop = comps.random(np.random.default_rng(7), P)
for rep in range(2):
    t = [time.perf_counter()]
    sim.generate_lattice_mixed(comps.proportions(pop)); sim.sync(); t.append(time.perf_counter())
    sim.inject(); sim.sync(); t.append(time.perf_counter())
    sim.run(500); sim.sync(); t.append(time.perf_counter())
    sim.realisation_tallies(); t.append(time.perf_counter())
    print("generate %.3f s, inject %.3f s, run %.3f s, tallies %.3f s" % tuple(b - a for a, b in zip(t, t[1:])))
sim.close()
