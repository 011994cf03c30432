#!/bin/bash
# round-2 measurement pass: parity suite, the bench configs, the ncu launch list and --set full captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "UserWarning\|warnings.warn" | tail -60 > gpurun_out/r2d_pytest.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench_c2.json 2> gpurun_out/r2d_bench_c2.err
python bench.py --config c1 --steps 20 --warmup 3 > gpurun_out/r2d_bench_c1.json 2> gpurun_out/r2d_bench_c1.err
timeout 600 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_c4.json 2> gpurun_out/r2d_bench_c4.err
timeout 900 python bench.py --config c3 --steps 5 --warmup 3 > gpurun_out/r2d_bench_c3.json 2> gpurun_out/r2d_bench_c3.err
python bench.py --config train-tail --steps 10 > gpurun_out/r2d_bench_tail.json 2> gpurun_out/r2d_bench_tail.err
# every launch of one bench step (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2d_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2d_ncu_bench.log 2>&1
# --set full: the decode kernel, the bf16 stem (tc_gemm_kernel<32,1,1>), one trunk kernel of each kind
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode3_kernel -s 1 -c 1 -o gpurun_out/r2d_decode3 \
    python tools/ncu_decode.py 32 > gpurun_out/r2d_ncu_dec.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:tc_gemm_kernel|pw_mma_kernel" -c 6 -o gpurun_out/r2d_video \
    python tools/ncu_video.py 1 bf16 > gpurun_out/r2d_ncu_video.log 2>&1
for r in r2d_decode3 r2d_video; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null
done
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2d_pytest.txt; ls -la gpurun_out | tail -20
