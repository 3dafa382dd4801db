#!/bin/bash
# The multi-GPU measurement pass: bench.py and the sharded lines of bench_configs.py on N GPUs of one box.
#   tools/multi_gpu_run.sh N  ->  gpurun_out/r2_bench_${N}gpu.json, gpurun_out/r2_configs_${N}gpu.jsonl
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
NCCL_DEBUG=INFO $TR bench.py --gpus $N > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -c 400 gpurun_out/r2_bench_${N}gpu.json; echo
grep -m3 -i "NVLS\|via P2P\|Connected all" gpurun_out/r2_bench_${N}gpu.err | cut -c1-160
: > gpurun_out/r2_configs_${N}gpu.jsonl
for only in "configs[2] H=2" "configs[2] H=8" "configs[3]"; do
  $TR bench_configs.py --only "$only" >> gpurun_out/r2_configs_${N}gpu.jsonl 2>> gpurun_out/r2_configs_${N}gpu.err
done
$TR bench_configs.py --only "configs[4]" --kernels tc_tf32,tc >> gpurun_out/r2_configs_${N}gpu.jsonl 2>> gpurun_out/r2_configs_${N}gpu.err
cut -c1-330 gpurun_out/r2_configs_${N}gpu.jsonl
