#!/bin/bash
# N-GPU pass (N = number of visible GPUs): the scaling line and the BASELINE configs named for this GPU count.
out=gpurun_out; mkdir -p $out
N=$(python -c "import torch; print(torch.cuda.device_count())")
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
timeout 500 $TR bench.py --gpus $N --steps 100 --no-extras > $out/r02_bench_n$N.json 2> $out/r02_bench_n$N.err; tail -c 300 $out/r02_bench_n$N.err
if [ "$N" = "8" ]; then W="config3 atari_vbn"; else W="nsra_bipedal"; fi
for w in $W; do
  st=100; [ $w = atari_vbn ] && st=10
  timeout 400 $TR bench.py --gpus $N --steps $st --no-extras --workload $w > $out/r02_bench_${w}_n$N.json 2> $out/r02_bench_${w}_n$N.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$out/r02_bench_*_n$N.json") + ["$out/r02_bench_n$N.json"]):
    try:
        d = json.load(open(f)); print(f, round(d["value"], 1), round(d["ms_per_step"], 3), round(d["e2e"]["value"], 1), d["gpu_launches"], d.get("cuda_graph"), [(k["kernel"], round(k["ms"], 3)) for k in d.get("kernels", [])], d["clocks"])
    except Exception as e:
        print(f, "failed", e, open(f.replace(".json", ".err")).read()[-600:])
PY
