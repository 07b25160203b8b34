#!/bin/bash
# Brick layout of the weights in the product: parity on the device, then the 256^3 x 64 disorder sweep both ways.
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_exciton_path.py -x -q -m gpu 2>&1 | tail -3
for real in 16 64; do
for brick in 0 1; do
  for blk in "" 32 8 4; do
    echo "== 256^3 x $real HF_X_BRICK=$brick HF_X_BLOCK=${blk:-auto}"
    HF_X_BRICK=$brick HF_X_BLOCK=$blk timeout 600 python scripts/bench_excitons.py --L 256 --real $real --B 1000000 --events 50 --steps 3
  done
done
done
echo "== 128^3 x 1 (L2-resident) brick forced"
HF_X_BRICK=1 python scripts/bench_excitons.py --L 128 --B 1000000 --events 50 --steps 3
HF_X_BRICK=0 python scripts/bench_excitons.py --L 128 --B 1000000 --events 50 --steps 3
} > gpurun_out/brick_product.txt 2>&1
tail -40 gpurun_out/brick_product.txt
