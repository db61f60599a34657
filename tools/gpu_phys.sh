#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_physics.py tests/test_gpu_sampler.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -25
timeout 120 python tools/prof_physics.py
} 2>&1 | grep -v "^\[W\|Warning" | tee gpurun_out/phys.log
