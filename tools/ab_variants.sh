#!/bin/bash
# A/B of libphoenix_b200<suffix>.so variants on the headline bench (same box, interleaved twice):  tools/ab_variants.sh _a _b
for rep in 1 2; do
for v in "" "$@"; do
  lib=phoenix_drone_simulation_b200/libphoenix_b200$v.so
  echo -n "variant '$v': "
  PDX_LIB=$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ppo-rollout --large-envs 0 2>&1 | python -c "import sys, json; d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']/1e9, 3), 'G  fixed', round(d['fixed_policy']['value']/1e9, 3))"
done; done
