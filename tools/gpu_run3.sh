#!/bin/bash
# GPU pass 3 of round 2: evaluate v3 (theta by TMA, in-place conversion) -- correctness, timing, role counters.
out=gpurun_out; mkdir -p $out; L=estorch_b200/lib
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -s -k "f16" 2>&1 | tail -14 | cut -c1-400 | tee $out/r02c_f16_tests.log
for v in "" _g4; do ESTK_LIBRARY=$L/libestk$v.so timeout 120 python tools/eval_time.py 2048 f16 2>&1 | tail -1 | sed "s/^/v3$v: /" | tee -a $out/r02c_eval_time.txt; done
ESTK_LIBRARY=$L/libestk_prof.so ESTK_TC_PROFILE=1 timeout 120 python tools/f16_profile.py 2>&1 | tail -16 | tee $out/r02c_f16_profile.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $out/r02c_gpu_tests.log
timeout 300 python bench.py --steps 100 > $out/r02c_bench_n1.json 2> $out/r02c_bench_n1.err; tail -c 300 $out/r02c_bench_n1.err; head -c 300 $out/r02c_bench_n1.json
