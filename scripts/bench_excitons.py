#!/usr/bin/env python
"""Throughput of the exciton path: B trajectories on one L^3 lattice, n events per launch."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hyperfluorescence_b200 import _native as nat

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=128)
ap.add_argument("--B", type=int, default=1_000_000)
ap.add_argument("--events", type=int, default=100)
ap.add_argument("--cutoff", type=float, default=4.0)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--real", type=int, default=1, help="disorder realisations (trajectories are split evenly over them)")
a = ap.parse_args()
sim = nat.Sim((a.L, a.L, a.L), (0.84, 0.15, 0.01), 0.1, 1, 1, n_trajectories=a.B, n_realisations=a.real, seed=1)
sim.generate_lattice()
sim.x_configure(nat.x_params_from_dict(dict(cutoff=a.cutoff)))
sim.x_spawn()
sim.x_run(a.events); sim.sync(); sim.run_timing()
for _ in range(a.steps):
    sim.x_run(a.events)
ms, n = sim.run_timing()
t = sim.x_tallies()
M = sim.x_tables()["g6"].size
ev = a.B * a.events * n
bytes_per_event = M * 9 + (M + 7) // 8 + 2 * 21 + 16
print(json.dumps(dict(launch=sim.x_launch_info(), L=a.L, B=a.B, realisations=a.real, cutoff=a.cutoff, M=int(M), events_per_s=ev / (ms * 1e-3), ms_per_launch=ms / n,
                      bytes_per_event=bytes_per_event, algorithmic_GBs=ev * bytes_per_event / (ms * 1e-3) / 1e9,
                      iqe=(t["rad_tadf"] + t["rad_fluo"]) / max(1, t["generated"] - a.B))))
sim.close()
