#!/bin/bash
# N-GPU box (N = $1): inference shard (c2) with the single-poller clock sampler
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2u_bench_c2_n$N.json 2> gpurun_out/r2u_bench_c2_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2u_bench_c2_n$N.json").read().strip().splitlines()[-1])
print("c2", d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"])
PY
