#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "UserWarning\|warnings.warn" | tail -60 > gpurun_out/r2f_pytest.txt
python tools/train_step_once.py 8 3 > gpurun_out/r2f_train_once.txt 2>&1
timeout 600 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2f_bench_c4.json 2> gpurun_out/r2f_bench_c4.err
timeout 900 python bench.py --config c3 --steps 5 --warmup 3 > gpurun_out/r2f_bench_c3.json 2> gpurun_out/r2f_bench_c3.err
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2f_pytest.txt; cat gpurun_out/r2f_train_once.txt | grep step
python - <<'PY'
import json
for f in ("c4", "c3"):
    try:
        d = json.loads(open(f"gpurun_out/r2f_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d.get("stage_ms"))
    except Exception as e:
        print(f, "ERR", e)
PY
