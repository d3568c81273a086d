#!/bin/bash
bash tools/launcher_check.sh
timeout 200 python -m pytest tests/test_api_gpu.py -q -x -k "n_proc_2" 2>&1 | tail -3
timeout 120 python -m pytest tests/test_kernels_gpu.py -q -x -k "nan_returns" 2>&1 | tail -3
