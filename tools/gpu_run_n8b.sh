#!/bin/bash
# final N-GPU pass: scaling line + per-phase timelines with the returns all-gather in the kernel and through NCCL
out=gpurun_out; mkdir -p $out
N=$(python -c "import torch; print(torch.cuda.device_count())")
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
timeout 200 $TR bench.py --gpus $N --steps 200 --no-extras > $out/r02q_bench_n$N.json 2> $out/r02q_bench_n$N.err; echo "exit $?"
timeout 120 $TR tools/step_timeline.py 2>/dev/null | tee $out/r02q_timeline_n$N.txt
ESTORCH_B200_PEER_GATHER=0 timeout 120 $TR tools/step_timeline.py 2>/dev/null | tee $out/r02q_timeline_n${N}_nccl_gather.txt
python - <<PY
import json
try:
    d = json.load(open("$out/r02q_bench_n$N.json")); print(round(d["value"], 1), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), d["gpu_launches"], d.get("cuda_graph"), [(k["kernel"], round(k["ms"], 4)) for k in d["kernels"]], d["clocks"])
except Exception as e:
    print("failed", e, open("$out/r02q_bench_n$N.err").read()[-800:])
PY
