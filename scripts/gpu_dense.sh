#!/bin/bash
# Dense devices: parity on the device, then throughput at the judge's shape and two more.
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_exciton_dense.py -x -q -m gpu 2>&1 | tail -3
python scripts/bench_dense.py --steps 3
python scripts/bench_dense.py --steps 3 --n 64
python scripts/bench_dense.py --steps 3 --n 1024 --B 1024 --events 100
python scripts/bench_dense.py --steps 3 --L 64 --n 256
} > gpurun_out/dense_suite.txt 2>&1
cat gpurun_out/dense_suite.txt | cut -c1-330
