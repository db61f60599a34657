#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_dit.py -m gpu -q --tb=short -p no:cacheprovider -x -k "chunked or 384" 2>&1 | tail -15 | tee gpurun_out/t.log
