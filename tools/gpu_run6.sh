#!/bin/bash
out=gpurun_out; mkdir -p $out; L=estorch_b200/lib
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "f16_tensor or fp16_table" 2>&1 | tail -3 | cut -c1-300 | tee $out/r02f_tests.log
for v in "" _prev; do ESTK_LIBRARY=$L/libestk$v.so timeout 120 python tools/eval_time.py 2048 f16 2>&1 | tail -1 | sed "s/^/v3e$v: /" | tee -a $out/r02f_eval_time.txt; done
ESTK_LIBRARY=$L/libestk_prof.so ESTK_TC_PROFILE=1 timeout 120 python tools/f16_profile.py 2>&1 | tail -16 | tee $out/r02f_f16_profile.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | cut -c1-300 | tee $out/r02f_gpu_tests.log
timeout 600 python bench.py --steps 100 > $out/r02f_bench_n1.json 2> $out/r02f_bench_n1.err; tail -c 300 $out/r02f_bench_n1.err; python -c "
import json; d=json.load(open('$out/r02f_bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], [(k['kernel'],round(k['ms'],3),round(k['frac'],3)) for k in d['kernels']]); print(json.dumps(d.get('extra'))[:1500]); print(d['cpu_baseline'])"
