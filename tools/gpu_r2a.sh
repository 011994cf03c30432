#!/bin/bash
# round-2 first GPU pass: parity suite + the four bench configs
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2a_pytest.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err
python bench.py --config c1 --steps 10 --warmup 3 > gpurun_out/r2a_bench_c1.json 2> gpurun_out/r2a_bench_c1.err
timeout 600 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_c4.json 2> gpurun_out/r2a_bench_c4.err
python bench.py --config train-tail --steps 10 > gpurun_out/r2a_bench_tail.json 2> gpurun_out/r2a_bench_tail.err
tail -3 gpurun_out/r2a_pytest.txt
