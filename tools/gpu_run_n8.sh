#!/bin/bash
# N-GPU pass (N = number of visible GPUs): NCCL + CUDA-graph correctness, then the scaling line(s).
out=gpurun_out; mkdir -p $out
N=$(python -c "import torch; print(torch.cuda.device_count())")
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
LOG_INTERVAL=3 timeout 400 $TR tools/multi_gpu_check.py 2>&1 | grep -v "^W\|^\[W\|OMP_NUM\|^\*\*\*" | tail -6 | tee $out/r02_n${N}_check.log
timeout 600 $TR bench.py --gpus $N --steps 200 --no-extras > $out/r02_bench_n$N.json 2> $out/r02_bench_n$N.err; tail -c 300 $out/r02_bench_n$N.err
timeout 400 $TR bench.py --gpus $N --steps 100 --no-extras --workload config3 > $out/r02_bench_config3_n$N.json 2> $out/r02_bench_config3_n$N.err
python - <<PY
import json
for f in ("r02_bench_n$N.json", "r02_bench_config3_n$N.json"):
    try:
        d = json.load(open("$out/" + f)); print(f, round(d["value"], 1), round(d["ms_per_step"], 3), round(d["e2e"]["value"], 1), d["gpu_launches"], d.get("cuda_graph"), [(k["kernel"], round(k["ms"], 3)) for k in d.get("kernels", [])], d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
