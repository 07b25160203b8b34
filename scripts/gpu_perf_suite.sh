#!/bin/bash
# Secondary kernels, one line each (regression check against the numbers in DESIGN.md).
set -u
mkdir -p gpurun_out
{
python scripts/bench_small.py --C 4
python scripts/bench_small.py --C 2
python scripts/bench_small.py --C 8
python scripts/bench_p_dense.py --C 164
python scripts/bench_p_dense.py --C 32 --B 8192
python scripts/bench_p_dense.py --C 10 --B 16384 --warp 2>/dev/null || python scripts/bench_p_dense.py --C 10 --B 16384
python scripts/bench_dense.py
python scripts/bench_dense.py --B 16384 --n 64
python scripts/bench_dense.py --B 4096 --n 1024 --events 100
} > gpurun_out/perf_suite.jsonl 2> gpurun_out/perf_suite.err
cat gpurun_out/perf_suite.jsonl; tail -5 gpurun_out/perf_suite.err
if [ -x scripts/micro/tma_box ]; then scripts/micro/tma_box > gpurun_out/tma_box.txt 2>&1; cat gpurun_out/tma_box.txt; fi
