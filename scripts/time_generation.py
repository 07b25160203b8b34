import time, sys
sys.path.insert(0, '/root/repo')
from hyperfluorescence_b200 import _native as nat
sim = nat.Sim((256,256,256), (0.84,0.15,0.01), 0.1, 1, 1, n_trajectories=64, n_realisations=64, seed=1)
sim.sync(); t0=time.perf_counter(); sim.generate_lattice(); sim.sync(); print("generate 64 x 256^3:", time.perf_counter()-t0, "s")
t0=time.perf_counter(); sim.x_configure(); sim.sync(); print("x_configure (weights):", time.perf_counter()-t0, "s")
sim.close()
