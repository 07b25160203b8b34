#!/bin/bash
# Round-2 evidence: launch list of the bench command, then one full capture per kernel, reduced to text on the box.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-ga"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
KERNEL=hf_run_c1 SKIP=1 TAG=c1 bash scripts/gpu_ncu_any.sh $CMD
KERNEL=hfx_run SKIP=1 TAG=x bash scripts/gpu_ncu_any.sh python scripts/bench_excitons.py --L 128 --B 1000000 --events 50 --steps 2
KERNEL=hf_run_kernel SKIP=2 TAG=p164 bash scripts/gpu_ncu_any.sh python scripts/bench_p_dense.py --C 164
KERNEL=hfxd_run SKIP=2 TAG=xd bash scripts/gpu_ncu_any.sh python scripts/bench_dense.py
KERNEL=hf_run_small SKIP=1 TAG=small bash scripts/gpu_ncu_any.sh python scripts/bench_small.py --C 4
ls -la gpurun_out
