#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_audio_gpu.py -m gpu -q --tb=short 2>&1 | grep -v "UserWarning\|warnings.warn" | tail -60 > gpurun_out/r2h_pytest.txt
cat gpurun_out/r2h_pytest.txt | tail -40
