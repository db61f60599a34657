#!/bin/bash
# A/B of two library builds on ONE GPU box (box-to-box spread is ~1.5 %, so variants must share a box):
#   build the variant with a -D switch into build/dbg/libpdk_<NAME>.so here (see tools/gemm_variants.sh), then
#   gpurun -- 'VARIANTS="NAME1 NAME2" bash tools/gpu_ab.sh'
# Measurement switches (PDK_NO_PAIR, PDK_FORCE_PAIR, PDK_NO_WIDE, PDK_WIDE_ALL, PDK_NO_ATTN_BALANCE, PDK_NO_FUSED_TRANSITION,
# PDK_NO_PDL) are read from the environment ONLY by the debug variants (built with -DPDK_MEASURE by tools/gemm_variants.sh);
# the release library ignores them.  Graph vs eager launches: DiffusionSampler(use_cuda_graph=...).
mkdir -p gpurun_out
{
for rep in 1 2; do
  echo "### product"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"ms_per_step[^,]*' | head -1
  for v in $VARIANTS; do
    echo "### $v"; PHYSDOCK_B200_LIB=/root/repo/build/dbg/libpdk_$v.so timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"ms_per_step[^,]*' | head -1
  done
done
} 2>&1 | tee gpurun_out/ab.log
