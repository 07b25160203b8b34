#!/bin/bash
# ncu --set full of one kernel of one command:  KERNEL=regex SKIP=n TAG=name bash scripts/gpu_ncu_any.sh <command...>
set -u
mkdir -p gpurun_out
TAG="${TAG:-any}"
"$@" > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KERNEL} -s ${SKIP:-1} -c 1 -f -o gpurun_out/${TAG}_prof "$@" > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/${TAG}_ncu.log; cat gpurun_out/${TAG}_plain.log | cut -c1-300
