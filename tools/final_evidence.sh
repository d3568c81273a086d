#!/bin/bash
# One GPU-box pass (1 GPU) that produces the round's evidence: GPU tests, smoke, the N=1 bench line (ours and the
# reference arm), compute-sanitizer over the hot kernels, the ncu launch list of the bench command and one
# `--set full` capture of the two hot kernels.  Usage (repo root, on the GPU box):  bash tools/final_evidence.sh r02
tag="${1:-r02}"; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | cut -c1-300 | tee $out/${tag}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $out/${tag}_smoke.log
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; tail -c 300 $out/${tag}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $out/${tag}_bench_reference_n1.json 2> $out/${tag}_bench_reference_n1.err
bash tools/gpu_evidence.sh $tag
python - <<PY
import json
d = json.load(open("$out/${tag}_bench_n1.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], [(k["kernel"], round(k["ms"], 4), round(k["frac"], 3)) for k in d["kernels"]], d["clocks"])
r = json.load(open("$out/${tag}_bench_reference_n1.json")); print("reference arm:", r["value"], r["cpu_baseline"]["kind"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["sample_P"])
PY
