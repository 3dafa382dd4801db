#!/bin/bash
# A/B of libphoenix_b200*.so variants on the fused collector (same box, interleaved twice)
for rep in 1 2; do
for v in "" "$@"; do
  lib=phoenix_drone_simulation_b200/libphoenix_b200$v.so
  echo -n "variant '$v': "
  PDX_LIB=$lib timeout 200 python tools/bench_collect3.py DroneHoverBulletEnv-v0 ${KERNEL:-tc_tf32} 65536 2>&1 | grep -E "k_collect" | cut -c1-60
done; done
