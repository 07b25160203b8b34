#!/bin/bash
# End-of-round pass on one GPU: parity suite, smoke, both bench arms, launch list and the ncu captures that
# profiles/traffic.json is stamped from (headline kernel, exciton kernel).
set -u
mkdir -p gpurun_out
bash scripts/gpu_round2.sh tests bench > gpurun_out/final_round2.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-ga"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
KERNEL=hf_run_c1 SKIP=1 TAG=c1 bash scripts/gpu_ncu_any.sh $CMD
KERNEL=hfx_run SKIP=1 TAG=xb bash scripts/gpu_ncu_any.sh python scripts/bench_excitons.py --L 128 --B 1000000 --events 50 --steps 2
python -c "
import json; d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])
for k in ['configs1','exciton_path','disorder_sweep','high_density','device_eqe']: print(k,d[k]['value'],d[k].get('e2e',{}).get('value'))
print('ga',d['ga']['value'],d['ga']['eqe_sweep']['value'])
r=json.load(open('gpurun_out/bench_ref.json')); print('ref',r['value'])
"
