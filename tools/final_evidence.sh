#!/bin/bash
# One GPU-box pass that produces the round's evidence: GPU tests, smoke, the N=1 bench line,
# the ncu launch list of the bench command and one `--set full` capture of the two hot kernels.
# Usage (from the repo root, on the GPU box):  bash tools/final_evidence.sh <tag>
tag="${1:-r01}"
out=gpurun_out
mkdir -p $out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/gpu_tests_$tag.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $out/smoke_$tag.log
timeout 300 python bench.py > $out/bench_${tag}_n1.json 2> $out/bench_${tag}_n1.err
tail -c 400 $out/bench_${tag}_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_under_ncu_$tag.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:eval_mlp_tc|rank_grad_kernel' -s 2 -c 2 -f \
    -o $out/prof_$tag python tools/profile_kernels.py --pairs 2048 --iters 2 > $out/ncu_full_$tag.log 2>&1
tail -2 $out/ncu_full_$tag.log
python - <<PY
import json
d = json.load(open("$out/bench_${tag}_n1.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], [(k["kernel"], round(k["ms"], 4)) for k in d["kernels"]], d["clocks"])
PY
