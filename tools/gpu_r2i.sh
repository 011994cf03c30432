#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | grep -v "UserWarning\|warnings.warn" | tail -40 > gpurun_out/r2i_pytest.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2i_bench_c2.json 2> gpurun_out/r2i_bench_c2.err
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/r2i_pytest.txt | head
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2i_bench_c2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["stage_ms"], d["e2e"]["value"])
PY
