#!/bin/bash
out=gpurun_out; mkdir -p $out; L=estorch_b200/lib
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "f16_tensor" 2>&1 | tail -3 | cut -c1-300 | tee $out/r02i_tests.log
for v in "" _pf0 _pf24; do ESTK_LIBRARY=$L/libestk$v.so timeout 120 python tools/eval_time.py 2048 f16 2>&1 | tail -1 | sed "s/^/v3g$v: /" | tee -a $out/r02i_eval_time.txt; done
ESTK_LIBRARY=$L/libestk_prof.so ESTK_TC_PROFILE=1 timeout 120 python tools/f16_profile.py 2>&1 | tail -16 | tee $out/r02i_f16_profile.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | cut -c1-300 | tee $out/r02i_gpu_tests.log
timeout 600 python bench.py --steps 100 > $out/r02i_bench_n1.json 2> $out/r02i_bench_n1.err; tail -c 300 $out/r02i_bench_n1.err; python -c "
import json; d=json.load(open('$out/r02i_bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['cuda_graph'], [(k['kernel'],round(k['ms'],3),round(k['frac'],3)) for k in d['kernels']]); print({k:(round(v.get('value',0),1), round(v.get('e2e',0),1)) if 'value' in v else v for k,v in d['extra'].items()})"
