#!/bin/bash
# ncu evidence for the exciton kernel (hfx_run_kernel): plain run first, then launch list + full capture.
set -u
mkdir -p gpurun_out
CMD="python scripts/bench_excitons.py --L 128 --B 1000000 --events 50 --steps 2"
$CMD > gpurun_out/x_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/x_launches.csv $CMD > gpurun_out/x_ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/x_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hfx_run -s 1 -c 1 -f -o gpurun_out/x_prof $CMD > gpurun_out/x_ncu_full.log 2>&1
echo "full capture exit $?"
cat gpurun_out/x_plain.log
