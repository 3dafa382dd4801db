for b in 64 96 128 160 192 224 448; do
  echo -n "block $b: "
  PDX_BLOCK=$b python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ppo-rollout --large-envs 0 2>&1 | python -c "import sys, json; d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']/1e9, 3), 'G  fixed', round(d['fixed_policy']['value']/1e9, 3))"
done
