#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | cut -c1-300 | tee $out/r02n_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $out/r02n_smoke.log
timeout 200 python tools/step_timeline.py 2>/dev/null | tee $out/r02n_timeline_n1.txt
timeout 900 python bench.py --no-extras > $out/r02n_bench_n1.json 2> $out/r02n_bench_n1.err; tail -c 300 $out/r02n_bench_n1.err
python - <<PY
import json
d = json.load(open("$out/r02n_bench_n1.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], [(k["kernel"], round(k["ms"], 4), round(k["frac"], 3)) for k in d["kernels"]], d["clocks"])
PY
