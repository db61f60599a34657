#!/bin/bash
# Builds debug variants of the library as separate .so files under build/dbg -- run HERE (needs nvcc).
#   tools/gemm_variants.sh NO_EPI NO_STORE        -> -DPDK_DBG_NO_EPI, -DPDK_DBG_NO_STORE   (gemm_umma.cu)
#   tools/gemm_variants.sh T_NOLN T_NOHID T_NOFINAL                                         (transition_umma.cu)
# Every variant is also compiled with -DPDK_MEASURE: the run-time A/B switches (PDK_NO_PAIR, PDK_NO_PDL, ...) and the
# attention trace hook exist only in these builds, never in the release library.
cd "$(dirname "$0")/../physdock_b200/csrc"
SRC="gemm_umma.cu transition_umma.cu tmap.cu attention_umma.cu pairbias.cu glue.cu coords.cu physics.cu capi.cu"
mkdir -p ../../build/dbg
for v in "$@"; do
  case $v in T_*|ATTN_*) def=PDK_$v ;; *) def=PDK_DBG_$v ;; esac      # SKIP -> -DPDK_DBG_SKIP (capi.cu: PDK_SKIP=<labels>, tools/time_step_skip.py)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DPDK_MEASURE -D$def -o ../../build/dbg/libpdk_$v.so $SRC &
done
wait; ls -la ../../build/dbg/
