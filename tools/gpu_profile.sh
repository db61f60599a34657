#!/bin/bash
# ncu captures of the kernels the north star names (one GPU): full-set reports into gpurun_out/, summarised by
# tools/ncu_summary.py into profiles/ afterwards.  Also the per-launch duration list of one bench step.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
N=2 timeout 300 $NCU -k regex:attention_umma -s 1 -c 1 -o gpurun_out/attn_atom python tools/prof_attention.py > gpurun_out/ncu_attn.log 2>&1
B=16 H=16 S=256 N=2 timeout 300 $NCU -k regex:attention_umma -s 1 -c 1 -o gpurun_out/attn_token python tools/prof_attention.py >> gpurun_out/ncu_attn.log 2>&1
timeout 300 $NCU -k regex:gemm_umma -s 2 -c 2 -o gpurun_out/gemm python tools/prof_gemm.py > gpurun_out/ncu_gemm.log 2>&1
timeout 300 $NCU -k regex:gemm_umma -s 2 -c 2 -o gpurun_out/qkv python tools/prof_qkv.py > gpurun_out/ncu_qkv.log 2>&1
timeout 300 $NCU -k regex:transition_umma -s 1 -c 1 -o gpurun_out/transition python tools/prof_transition.py > gpurun_out/ncu_trans.log 2>&1
timeout 300 $NCU -k regex:pair_energy_grad -s 3 -c 1 -o gpurun_out/physics python tools/prof_physics.py > gpurun_out/ncu_phys.log 2>&1
timeout 600 $NCU -k "regex:centre_augment|euler_kernel|rigid_align|template_eps|template_pick|descent_update" -s 6 -c 8 -o gpurun_out/coords python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-extras --physics > gpurun_out/ncu_coords.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extras > gpurun_out/ncu_bench.log 2>&1
timeout 200 python tools/time_gemm.py > gpurun_out/time_gemm.log 2>&1
timeout 200 python tools/time_attention.py >> gpurun_out/time_gemm.log 2>&1
PHYSDOCK_B200_LIB=/root/repo/build/dbg/libpdk_SKIP.so timeout 120 python tools/trace_attention.py > gpurun_out/attention_timeline.txt 2>&1
timeout 600 python tools/time_step_skip.py > gpurun_out/step_attribution.txt 2>&1
timeout 900 python tools/time_trunk.py > gpurun_out/trunk_vs_sampling.txt 2>&1
ls -la gpurun_out/*.ncu-rep
for f in gpurun_out/ncu_attn.log gpurun_out/ncu_gemm.log gpurun_out/ncu_trans.log gpurun_out/ncu_phys.log gpurun_out/ncu_coords.log; do tail -n 2 $f; done
