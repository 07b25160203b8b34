#!/bin/bash
# bench.py at N GPUs of one node, launched exactly as the driver does.
set -u
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit $?"
cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
echo "reference arm N=$N exit $?"
cat gpurun_out/bench_ref_n$N.json
