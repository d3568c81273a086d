#!/bin/bash
out=gpurun_out; mkdir -p $out; L=estorch_b200/lib
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "f16_tensor" 2>&1 | tail -3 | cut -c1-300 | tee $out/r02j_tests.log
for v in "" _pf0; do ESTK_LIBRARY=$L/libestk$v.so timeout 120 python tools/eval_time.py 2048 f16 2>&1 | tail -1 | sed "s/^/v3h$v: /" | tee -a $out/r02j_eval_time.txt; done
ESTK_LIBRARY=$L/libestk_prof.so ESTK_TC_PROFILE=1 timeout 120 python tools/f16_profile.py 2>&1 | tail -16 | tee $out/r02j_f16_profile.txt
