#!/bin/bash
# 2-GPU pass for the in-kernel cross-GPU reduction: correctness vs NCCL, graph replay, timings with and without it.
out=gpurun_out; mkdir -p $out
N=$(python -c "import torch; print(torch.cuda.device_count())")
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR tools/multi_gpu_check.py 2>&1 | grep -v "^W\|^\[W\|OMP_NUM\|^\*\*\*" | tail -16 | tee $out/r02k_n${N}_check.log
timeout 200 $TR tools/step_timeline.py > $out/r02k_timeline_n$N.txt 2> $out/r02k_timeline_n$N.err; echo "timeline exit $?"; cat $out/r02k_timeline_n$N.txt; grep -i "error\|Traceback" -A5 $out/r02k_timeline_n$N.err | tail -12
ESTORCH_B200_PEER=0 timeout 200 $TR tools/step_timeline.py > $out/r02k_timeline_n${N}_nccl.txt 2> /dev/null; cat $out/r02k_timeline_n${N}_nccl.txt
timeout 300 $TR bench.py --gpus $N --steps 200 --no-extras > $out/r02k_bench_n$N.json 2> $out/r02k_bench_n$N.err
python - <<PY
import json
for f in ("r02k_bench_n$N.json",):
    try:
        d = json.load(open("$out/" + f)); print(f, round(d["value"], 1), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), d["gpu_launches"], d.get("cuda_graph"), [(k["kernel"], round(k["ms"], 4)) for k in d["kernels"]])
    except Exception as e:
        print(f, "failed", e, open("$out/" + f.replace(".json", ".err")).read()[-800:])
PY
