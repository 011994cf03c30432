#!/bin/bash
# N-GPU box (N = $1): train-side data-parallel evidence — train-tail (gradient all-reduce) and c3 (whole train step with graph replay)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
timeout 600 $TR bench.py --gpus $N --config train-tail --steps 20 > gpurun_out/r2t_bench_tail_n$N.json 2> gpurun_out/r2t_bench_tail_n$N.err
timeout 900 $TR bench.py --gpus $N --config c3 --steps 20 --no-cpu-baseline > gpurun_out/r2t_bench_c3_n$N.json 2> gpurun_out/r2t_bench_c3_n$N.err
python - <<PY
import json
for f in ("tail", "c3"):
    try:
        d = json.loads(open(f"gpurun_out/r2t_bench_{f}_n$N.json").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], d["value"], d["ms_per_step"], d.get("collective"), d["e2e"]["value"])
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/r2t_bench_{f}_n$N.err").read()[-800:])
PY
