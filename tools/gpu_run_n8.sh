#!/bin/bash
# N-GPU pass (N = number of visible GPUs), ONE process start: the scaling line plus the other BASELINE configs.
out=gpurun_out; mkdir -p $out
N=$(python -c "import torch; print(torch.cuda.device_count())")
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
if [ "$N" = "8" ]; then X="config3,atari_vbn,nsra_bipedal"; else X="config3,nsra_bipedal"; fi
timeout ${LIMIT:-300} $TR bench.py --gpus $N --steps 200 --extras $X > $out/r02_bench_n$N.json 2> $out/r02_bench_n$N.err
echo "exit $?"
timeout 120 $TR tools/xr_check.py 2>&1 | grep "^W=\|FAILED\|iteration 0:" | sort | uniq | head -6 | tee $out/r02_xr_check_n$N.txt
[ "$N" = "8" ] && ESTORCH_B200_PEER=0 timeout 150 $TR tools/step_timeline.py > $out/r02_timeline_n${N}_nccl.txt 2> /dev/null; cat $out/r02_timeline_n${N}_nccl.txt
timeout 150 $TR tools/step_timeline.py > $out/r02_timeline_n$N.txt 2> $out/r02_timeline_n$N.err; echo "timeline exit $?"; cat $out/r02_timeline_n$N.txt; tail -3 $out/r02_timeline_n$N.err
grep -v "^W\|^\[W\|OMP_NUM\|^\*\*\*" $out/r02_bench_n$N.err | tail -5
python - <<PY
import json
try:
    d = json.load(open("$out/r02_bench_n$N.json"))
    print(round(d["value"], 1), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), d["gpu_launches"], d.get("cuda_graph"),
          [(k["kernel"], round(k["ms"], 4)) for k in d.get("kernels", [])], d["clocks"])
    for k, v in (d.get("extra") or {}).items():
        print(k, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items()})
except Exception as e:
    print("failed", e)
PY
