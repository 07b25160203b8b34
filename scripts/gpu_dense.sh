#!/bin/bash
# dense devices: parity tests first, then the throughput lines
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_exciton_dense.py tests/test_exciton_reference_limits.py -m gpu -q > gpurun_out/pytest_dense.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_dense.log
{
python scripts/bench_dense.py
python scripts/bench_dense.py --B 16384 --n 64
python scripts/bench_dense.py --B 4096 --n 1024 --events 100
python scripts/bench_dense.py --B 8192 --n 128
} > gpurun_out/dense.jsonl 2> gpurun_out/dense.err
cat gpurun_out/dense.jsonl; tail -3 gpurun_out/dense.err
