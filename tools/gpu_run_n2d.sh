#!/bin/bash
out=gpurun_out; mkdir -p $out
N=$(python -c "import torch; print(torch.cuda.device_count())")
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/xr_check.py > $out/r02k_xr_check_n$N.txt 2>&1; echo "exit $?"
grep "^rank\|^W=\|Traceback\|Error\|error" -A3 $out/r02k_xr_check_n$N.txt | grep -v "^\[W\|^W09\|iteration [1-9]" | head -12
timeout 300 $TR tools/multi_gpu_check.py > $out/r02k_n${N}_check.log 2>&1; echo "check exit $?"
grep "bit-identical\|peer-memory\|AssertionError\|graph replay ==" $out/r02k_n${N}_check.log | head -12
timeout 200 $TR tools/step_timeline.py 2>/dev/null | tee $out/r02k_timeline_n$N.txt
timeout 300 $TR bench.py --gpus $N --steps 200 --no-extras > $out/r02k_bench_n$N.json 2> $out/r02k_bench_n$N.err
python - <<PY
import json
try:
    d = json.load(open("$out/r02k_bench_n$N.json")); print(round(d["value"], 1), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), d["gpu_launches"], d.get("cuda_graph"), [(k["kernel"], round(k["ms"], 4)) for k in d["kernels"]])
except Exception as e:
    print("failed", e, open("$out/r02k_bench_n$N.err").read()[-800:])
PY
