#!/bin/bash
# 2-GPU pass: NCCL + CUDA-graph correctness, the train(n_proc=2) launcher test, scaling at N=2.
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/multi_gpu_check.py 2>&1 | grep -v "^W\|^\[W\|OMP_NUM" | tail -8 | tee $out/r02_n2_check.log
LOG_INTERVAL=3 timeout 600 $TR tools/multi_gpu_check.py 2>&1 | grep -v "^W\|^\[W\|OMP_NUM" | tail -6 | tee -a $out/r02_n2_check.log
timeout 600 python -m pytest tests/test_api_gpu.py -m gpu -x -q -k "n_proc_2" 2>&1 | tail -3 | cut -c1-400 | tee $out/r02_n2_launcher_test.log
timeout 900 $TR bench.py --gpus 2 --steps 100 --no-extras > $out/r02_bench_n2.json 2> $out/r02_bench_n2.err; tail -c 400 $out/r02_bench_n2.err
ESTORCH_B200_GRAPH=0 timeout 900 $TR bench.py --gpus 2 --steps 100 --no-extras > $out/r02_bench_n2_nograph.json 2> $out/r02_bench_n2_nograph.err
python - <<PY
import json
for f in ("r02_bench_n2.json", "r02_bench_n2_nograph.json"):
    try:
        d = json.load(open("$out/" + f)); print(f, round(d["value"], 1), round(d["ms_per_step"], 3), round(d["e2e"]["value"], 1), d["gpu_launches"], d.get("cuda_graph"), [(k["kernel"], round(k["ms"], 3)) for k in d["kernels"]])
    except Exception as e:
        print(f, "failed", e)
PY
