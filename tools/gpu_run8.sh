#!/bin/bash
out=gpurun_out; mkdir -p $out; L=estorch_b200/lib
for v in "" _wi1 _wi2 _wi4 _wi7; do ESTK_LIBRARY=$L/libestk$v.so timeout 120 python tools/eval_time.py 2048 f16 2>&1 | tail -1 | sed "s/^/whatif$v: /" | tee -a $out/r02h_eval_time.txt; done
timeout 600 python tools/step_time.py 28 6 2>&1 | tail -4 | tee $out/r02h_step_time.txt
timeout 300 python tools/step_time.py 26 3 2>&1 | tail -4 | tee -a $out/r02h_step_time.txt
