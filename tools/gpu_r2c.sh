#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_model_gpu.py tests/test_patch_reference_gpu.py tests/test_train_step_gpu.py "tests/test_gpu_parity.py::test_batch_70_chunked_decode_vs_oracle" -m gpu -q --tb=short -s 2>&1 | grep -v "^  warnings.warn\|UserWarning" > gpurun_out/r2c_pytest.txt
tail -5 gpurun_out/r2c_pytest.txt
