#!/bin/bash
# ncu evidence for the exciton kernel (hfx_run_kernel): plain run first, then launch list + full capture.
# X_ARGS overrides the workload (default: the north-star configuration), X_TAG the output prefix.
set -u
mkdir -p gpurun_out
ARGS="${X_ARGS:---L 128 --B 1000000 --events 50 --steps 2}"
TAG="${X_TAG:-x}"
CMD="python scripts/bench_excitons.py $ARGS"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hfx_run -s 1 -c 1 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture exit $?"
cat gpurun_out/${TAG}_plain.log
