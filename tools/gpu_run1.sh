#!/bin/bash
# GPU pass 1 of round 2: parity of the fp16 tensor-core evaluate + first timings.
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee $out/r02a_smi.txt
python -c "import os; print('cpus', os.cpu_count())" | tee -a $out/r02a_smi.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -s -k "f16" 2>&1 | tail -25 | tee $out/r02a_f16_tests.log
for m in f16 bf16 bf16s; do timeout 120 python tools/eval_time.py 2048 $m 2>&1 | tail -1 | tee -a $out/r02a_eval_time.txt; done
TC_MODE=f16 timeout 120 python tools/tc_profile.py 2>&1 | tail -20 | tee $out/r02a_tc_profile_f16.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $out/r02a_gpu_tests.log
timeout 300 python bench.py --steps 30 > $out/r02a_bench_n1.json 2> $out/r02a_bench_n1.err; tail -c 300 $out/r02a_bench_n1.err
