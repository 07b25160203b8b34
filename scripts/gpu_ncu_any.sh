#!/bin/bash
# ncu --set full of one kernel of one command, reduced ON THE BOX to text (the .ncu-rep files are 10-20 MB each and
# gpurun brings back at most 64 MiB):  KERNEL=regex SKIP=n TAG=name bash scripts/gpu_ncu_any.sh <command...>
set -u
mkdir -p gpurun_out
TAG="${TAG:-any}"
"$@" > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KERNEL} -s ${SKIP:-1} -c 1 -f -o gpurun_out/${TAG}_prof "$@" > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/${TAG}_ncu.log; cat gpurun_out/${TAG}_plain.log | cut -c1-300
if [ -f gpurun_out/${TAG}_prof.ncu-rep ]; then
    ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
    ncu -i gpurun_out/${TAG}_prof.ncu-rep --page details > gpurun_out/${TAG}_details.txt 2>/dev/null
    ncu -i gpurun_out/${TAG}_prof.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_sass.csv 2>/dev/null
    rm -f gpurun_out/${TAG}_prof.ncu-rep
fi
