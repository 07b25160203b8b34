#!/usr/bin/env python
"""Path P at small `charges`: thread-per-trajectory kernel (default) vs the warp-per-trajectory kernel."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperfluorescence_b200 import _native as nat
ap = argparse.ArgumentParser()
ap.add_argument("--dims", type=int, nargs=3, default=[10, 10, 5])
ap.add_argument("--B", type=int, default=1_000_000)
ap.add_argument("--C", type=int, default=4)
ap.add_argument("--events", type=int, default=200)
ap.add_argument("--warp", action="store_true")
a = ap.parse_args()
sim = nat.Sim(tuple(a.dims), (0.84, 0.15, 0.01), 0.1, a.C, 1, n_trajectories=a.B, seed=2, flags=1 if a.warp else 0)
sim.generate_lattice(); sim.inject()
sim.run(a.events); sim.sync(); sim.run_timing()
for _ in range(3):
    sim.run(a.events)
ms, n = sim.run_timing()
t = sim.tallies()
print(json.dumps(dict(dims=a.dims, B=a.B, charges=a.C, kernel="warp" if a.warp else "thread", events_per_s=t["fired"] * 3 / 4 / (ms * 1e-3),
                      ms_per_launch=ms / n, fired=t["fired"], stopped=t["stopped"], recombination=t["recombination"])))
sim.close()
