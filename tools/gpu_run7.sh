#!/bin/bash
out=gpurun_out; mkdir -p $out; L=estorch_b200/lib
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "f16_tensor" 2>&1 | tail -3 | cut -c1-300 | tee $out/r02g_tests.log
for v in "" _d1; do ESTK_LIBRARY=$L/libestk$v.so timeout 120 python tools/eval_time.py 2048 f16 2>&1 | tail -1 | sed "s/^/v3f$v: /" | tee -a $out/r02g_eval_time.txt; done
ESTK_LIBRARY=$L/libestk_prof.so ESTK_TC_PROFILE=1 timeout 120 python tools/f16_profile.py 2>&1 | tail -16 | tee $out/r02g_f16_profile.txt
timeout 600 python tools/step_time.py 28 6 2>&1 | tail -4 | tee $out/r02g_step_time.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | cut -c1-300 | tee $out/r02g_gpu_tests.log
