#!/bin/bash
# decode3 LL-exchange bring-up: debug tool first (bounded), then the parity suite and a bench line
mkdir -p gpurun_out
timeout 300 python tools/dec3_debug.py > gpurun_out/r2b_dec3_debug.txt 2>&1
echo "dec3_debug rc=$?" >> gpurun_out/r2b_dec3_debug.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "UserWarning\|warnings.warn" | tail -150 > gpurun_out/r2b_pytest.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err
cat gpurun_out/r2b_dec3_debug.txt; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2b_pytest.txt
