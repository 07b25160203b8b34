#!/usr/bin/env python
"""Path P (the reference's charge First-Reaction-Method) at high carrier density: B devices with C electrons + C holes."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperfluorescence_b200 import _native as nat
ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=128)
ap.add_argument("--B", type=int, default=2048)
ap.add_argument("--C", type=int, default=164)
ap.add_argument("--events", type=int, default=100)
a = ap.parse_args()
sim = nat.Sim((a.L, a.L, a.L), (0.84, 0.15, 0.01), 0.1, a.C, 1, n_trajectories=a.B, seed=2)
sim.generate_lattice(); sim.inject()
sim.run(a.events); sim.sync(); sim.run_timing()
for _ in range(3):
    sim.run(a.events)
ms, n = sim.run_timing()
t = sim.tallies()
print(json.dumps(dict(L=a.L, devices=a.B, charges=a.C, events_per_s=a.B * a.events * n / (ms * 1e-3), ms_per_launch=ms / n,
                      fired=t["fired"], stopped=t["stopped"], recombination=t["recombination"], emission=t["emission"])))
sim.close()
