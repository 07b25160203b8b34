#!/usr/bin/env python
"""hf_run_c1_kernel: the two launch shapes (384 x 2, 288 x 3) against the batch size (wave quantisation)."""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    from hyperfluorescence_b200 import _native as nat
    B = int(sys.argv[1])
    sim = nat.Sim((128, 128, 128), (0.84, 0.15, 0.01), 0.1, 1, 1, n_trajectories=B, seed=1)
    sim.generate_lattice(); sim.inject(); sim.run(200); sim.sync(); sim.run_timing()
    for _ in range(3):
        sim.run(500)
    ms, n = sim.run_timing()
    print(json.dumps(dict(B=B, shape="A (forced)" if os.environ.get("HF_C1_SHAPE_A") else "auto", events_per_s=B * 500 * n / (ms * 1e-3), ms_per_launch=ms / n)))
    sim.close()
else:
    for B in (100000, 113664, 125000, 200000, 250000, 500000, 1000000):
        for force in (True, False):
            env = dict(os.environ)
            if force:
                env["HF_C1_SHAPE_A"] = "1"
            else:
                env.pop("HF_C1_SHAPE_A", None)
            subprocess.run([sys.executable, __file__, str(B)], env=env)
