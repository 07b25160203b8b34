#!/usr/bin/env python
"""Wall time of each C-ABI call of one GA fitness sweep on the exciton path (ExcitonFitness), eight sweeps in a row."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hyperfluorescence_b200 import _native as nat
from hyperfluorescence_b200.evolution import Compositions
dims, P, T = (32, 32, 16), 256, 10_000
sim = nat.Sim(dims, (0.84, 0.15, 0.01), 0.1, 1, 1, n_trajectories=P * T, n_realisations=P, seed=1)
comps = Compositions(dims)
pop = comps.random(np.random.default_rng(7), P)
params = nat.x_params_from_dict({})
for rep in range(8):
    out = []
    def t(name, f):
        t0 = time.perf_counter(); f(); sim.sync(); out.append((name, round((time.perf_counter() - t0) * 1e3, 1)))
    t("generate", lambda: sim.generate_lattice_mixed(comps.proportions(pop)))
    t("x_configure", lambda: sim.x_configure(params))
    t("x_spawn", lambda: sim.x_spawn())
    t("x_run", lambda: sim.x_run(200))
    t("tallies", lambda: sim.x_realisation_tallies())
    print(out)
sim.close()
