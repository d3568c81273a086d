#!/bin/bash
# 2-GPU pass: NCCL + peer-memory + CUDA-graph correctness, the train(n_proc=2) launcher, kernel-level comparison,
# per-phase timeline and the scaling line at N=2.  Usage on a 2-GPU box:  bash tools/gpu_run_n2.sh
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR tools/multi_gpu_check.py > $out/n2_check.log 2>&1; echo "check exit $?"
grep "bit-identical\|peer-memory\|AssertionError\|graph replay ==" $out/n2_check.log | head -14
timeout 300 $TR tools/xr_check.py 2>&1 | grep "^rank\|^W=\|FAILED" | grep -v "iteration [1-9]" | tee $out/n2_xr_check.txt
bash tools/launcher_check.sh
timeout 600 python -m pytest tests/test_api_gpu.py -m gpu -x -q -k "n_proc_2 or peer_memory" 2>&1 | tail -3 | cut -c1-400
timeout 200 $TR tools/step_timeline.py 2>/dev/null | tee $out/n2_timeline.txt
timeout 300 $TR bench.py --gpus 2 --steps 200 --no-extras > $out/n2_bench.json 2> $out/n2_bench.err
python - <<PY
import json
try:
    d = json.load(open("$out/n2_bench.json")); print(round(d["value"], 1), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), d["gpu_launches"], d.get("cuda_graph"), [(k["kernel"], round(k["ms"], 4)) for k in d["kernels"]])
except Exception as e:
    print("failed", e, open("$out/n2_bench.err").read()[-800:])
PY
