#!/bin/bash
# Round-2 GPU pass: parity suite, smoke, bench (both arms), then the ncu evidence for the headline kernel.
# Output under gpurun_out/.  Usage: bash scripts/gpu_round2.sh [tests|bench|ncu ...]
set -u
mkdir -p gpurun_out
WHAT="${*:-tests bench ncu}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.csv 2>&1
if [[ " $WHAT " == *" tests "* ]]; then
    timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/pytest_gpu.log 2>&1
    echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
    tail -25 gpurun_out/pytest_gpu.log
    timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
    echo "smoke exit $?" | tee -a gpurun_out/smoke.log
    tail -3 gpurun_out/smoke.log
fi
if [[ " $WHAT " == *" bench "* ]]; then
    timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
    echo "bench exit $?"
    cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
    timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
    echo "reference arm exit $?"
    cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
fi
if [[ " $WHAT " == *" ncu "* ]]; then
    CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-ga"
    $CMD > gpurun_out/plain.log 2>&1 &&
    ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
    echo "launch list exit $?"
    $CMD > gpurun_out/plain2.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:hf_run_c1 -s 1 -c 1 -f -o gpurun_out/prof_c1 $CMD > gpurun_out/ncu_full.log 2>&1
    echo "full capture exit $?"
    tail -3 gpurun_out/ncu_full.log
fi
ls -la gpurun_out | head -40
