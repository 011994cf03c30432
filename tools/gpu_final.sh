#!/bin/bash
# final verification pass: smoke, the whole GPU suite, every bench configuration
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/final_smoke.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "UserWarning\|warnings.warn" | tail -30 > gpurun_out/final_pytest.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench_c2.json 2> gpurun_out/final_bench_c2.err
python bench.py --config c1 --steps 20 --warmup 3 > gpurun_out/final_bench_c1.json 2> gpurun_out/final_bench_c1.err
timeout 600 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_c4.json 2> gpurun_out/final_bench_c4.err
timeout 900 python bench.py --config c3 --steps 20 --warmup 3 > gpurun_out/final_bench_c3.json 2> gpurun_out/final_bench_c3.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
tail -2 gpurun_out/final_smoke.txt; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/final_pytest.txt
python - <<'PY'
import json
for f in ("c2", "c1", "c4", "c3", "ref"):
    try:
        d = json.loads(open(f"gpurun_out/final_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 3), d.get("stage_ms"), round(d["e2e"]["value"]), (d.get("roofline") or {}).get("frac"),
              (d.get("gpu_eager_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
