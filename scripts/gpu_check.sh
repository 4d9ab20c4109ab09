#!/bin/bash
# Run on the GPU box via gpurun: GPU parity tests, tail to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q -s 2>&1 | tee gpurun_out/pytest_gpu.log | tail -60
