#!/usr/bin/env python
"""Throughput of the high-density exciton devices: B devices x n excitons on one L^3 lattice."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperfluorescence_b200 import _native as nat

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=128)
ap.add_argument("--B", type=int, default=4096)
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--events", type=int, default=200)
ap.add_argument("--cutoff", type=float, default=4.0)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
sim = nat.Sim((a.L, a.L, a.L), (0.84, 0.15, 0.01), 0.1, 1, 1, n_trajectories=a.B, seed=1)
sim.generate_lattice()
sim.x_configure(nat.x_params_from_dict(dict(cutoff=a.cutoff)))
sim.xd_configure(a.n)
sim.xd_spawn()
sim.xd_run(a.events); sim.sync(); sim.run_timing()
for _ in range(a.steps):
    sim.xd_run(a.events)
ms, n = sim.run_timing()
t = sim.xd_tallies()
ev = a.B * a.events * n
lost = sum(v for k, v in t.items() if k.startswith(("rad_", "nonrad_", "ann_")))
print(json.dumps(dict(L=a.L, devices=a.B, excitons_per_device=a.n, density=a.n / a.L ** 3, cutoff=a.cutoff,
                      events_per_s=ev / (ms * 1e-3), ms_per_launch=ms / n,
                      annihilation_fraction=(t["ann_ss"] + t["ann_st"] + t["ann_ts"] + t["ann_tt"]) / max(1, lost),
                      iqe=(t["rad_tadf"] + t["rad_fluo"]) / max(1, lost))))
sim.close()
