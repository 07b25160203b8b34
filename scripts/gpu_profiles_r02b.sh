#!/bin/bash
# Round-2 evidence, second pass (after the brick layout and the dense-device records): one full capture per changed
# kernel, reduced to text on the box.
set -u
mkdir -p gpurun_out
KERNEL=hfx_run SKIP=1 TAG=xb bash scripts/gpu_ncu_any.sh python scripts/bench_excitons.py --L 128 --B 1000000 --events 50 --steps 2
HF_X_BLOCK=32 KERNEL=hfx_run SKIP=1 TAG=xb256 bash scripts/gpu_ncu_any.sh python scripts/bench_excitons.py --L 256 --real 64 --B 1000000 --events 50 --steps 2
HF_X_BLOCK=32 HF_X_BRICK=0 KERNEL=hfx_run SKIP=1 TAG=xr256 bash scripts/gpu_ncu_any.sh python scripts/bench_excitons.py --L 256 --real 64 --B 1000000 --events 50 --steps 2
HF_X_BLOCK=6 KERNEL=hfx_run SKIP=1 TAG=xb256a bash scripts/gpu_ncu_any.sh python scripts/bench_excitons.py --L 256 --real 64 --B 1000000 --events 50 --steps 2
KERNEL=hfxd_run SKIP=2 TAG=xd bash scripts/gpu_ncu_any.sh python scripts/bench_dense.py --steps 2
rm -f gpurun_out/xr256_sass.csv gpurun_out/xb256a_sass.csv
ls -la gpurun_out
