#!/bin/bash
# Builds debug variants of the library (gemm_umma.cu compile switches) as separate .so files -- run HERE (needs nvcc).
cd "$(dirname "$0")/../physdock_b200/csrc"
SRC="gemm_umma.cu transition_umma.cu tmap.cu attention_umma.cu pairbias.cu glue.cu coords.cu physics.cu capi.cu"
for v in NO_STORE NO_EPI NO_TMAWAIT; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DPDK_DBG_$v -o /root/repo/build/dbg/libpdk_$v.so $SRC &
done
wait; ls -la /root/repo/build/dbg/libpdk_*.so
