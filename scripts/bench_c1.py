#!/usr/bin/env python
"""Throughput of the charges == 1 step kernel alone: bench_c1.py [L [B [events per launch]]] (default: BASELINE configs[1], 64^3 x 1e5)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperfluorescence_b200 import _native as nat
L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
S = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
sim = nat.Sim((L, L, L), (0.84, 0.15, 0.01), 0.1, 1, 1, n_trajectories=B, seed=0x48464B4D43)
sim.generate_lattice(); sim.inject()
for _ in range(2):
    sim.run(S)
sim.sync(); sim.run_timing()
for _ in range(5):
    sim.run(S)
ms, n = sim.run_timing()
print(json.dumps(dict(lib=os.environ.get("HFKMC_LIB", "default"), events_per_s=B * S * n / (ms * 1e-3), ms_per_launch=ms / n, fired=sim.tallies()["fired"])))
sim.close()
