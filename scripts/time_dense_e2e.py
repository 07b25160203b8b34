#!/usr/bin/env python
"""Wall time of each C-ABI call of the dense-device end-to-end step (bench.py high_density e2e)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperfluorescence_b200 import _native as nat
sim = nat.Sim((128,) * 3, (0.84, 0.15, 0.01), 0.1, 1, 1, n_trajectories=4096, seed=1)
sim.generate_lattice(); sim.sync()
host = sim.download_lattice(0)
xp = nat.x_params_from_dict(dict(cutoff=4.0))
for rep in range(3):
    out = []
    def t(name, f):
        t0 = time.perf_counter(); f(); sim.sync(); out.append((name, round((time.perf_counter() - t0) * 1e3, 2)))
    t("upload", lambda: sim.upload_lattice(0, host["species"], host["homo"], host["lumo"], host["s1"], host["t1"]))
    t("x_configure", lambda: sim.x_configure(xp))
    t("xd_configure", lambda: sim.xd_configure(256))
    t("xd_spawn", lambda: sim.xd_spawn())
    t("xd_run", lambda: sim.xd_run(200))
    t("tallies", lambda: sim.xd_tallies())
    print(out)
