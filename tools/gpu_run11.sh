#!/bin/bash
out=gpurun_out; mkdir -p $out; L=estorch_b200/lib
for v in "" _g3; do ESTK_LIBRARY=$L/libestk$v.so timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "f16_tensor" 2>&1 | tail -1 | cut -c1-200; ESTK_LIBRARY=$L/libestk$v.so timeout 120 python tools/eval_time.py 2048 f16 2>&1 | tail -1 | sed "s/^/v3i$v: /" | tee -a $out/r02k_eval_time.txt; done
