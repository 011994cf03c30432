#!/bin/bash
# 2-GPU box: train path regression + the data-parallel evidence (gradient all-reduce through the library's communicator)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_model_gpu.py tests/test_train_step_gpu.py -m gpu -q --tb=short 2>&1 | grep -v "UserWarning\|warnings.warn" | tail -30 > gpurun_out/r2g_pytest.txt
python tools/train_step_once.py 8 3 2>&1 | grep "^step" > gpurun_out/r2g_train_once.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus 2 --config train-tail --steps 20 > gpurun_out/r2g_bench_tail_n2.json 2> gpurun_out/r2g_bench_tail_n2.err
timeout 900 $TR bench.py --gpus 2 --config c3 --steps 5 --no-cpu-baseline > gpurun_out/r2g_bench_c3_n2.json 2> gpurun_out/r2g_bench_c3_n2.err
timeout 600 $TR bench.py --gpus 2 --steps 10 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2g_bench_c2_n2.json 2> gpurun_out/r2g_bench_c2_n2.err
timeout 600 python bench.py --config c3 --steps 5 > gpurun_out/r2g_bench_c3_n1.json 2> gpurun_out/r2g_bench_c3_n1.err
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2g_pytest.txt; cat gpurun_out/r2g_train_once.txt
python - <<'PY'
import json
for f in ("tail_n2", "c3_n2", "c2_n2", "c3_n1"):
    try:
        d = json.loads(open(f"gpurun_out/r2g_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], d["value"], d["ms_per_step"], d.get("collective"), d["e2e"]["value"])
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/r2g_bench_{f}.err").read()[-600:])
PY
