#!/bin/bash
# Evidence pass (1 GPU): compute-sanitizer over the hot kernels, ncu launch list of the bench command, one
# `ncu --set full` capture of the two hot kernels.  Usage on the GPU box: bash tools/gpu_evidence.sh r02
tag="${1:-r02}"; out=gpurun_out; mkdir -p $out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tests/sanitizer_pass.py > $out/${tag}_sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitizer_pass ok' $out/${tag}_sanitizer_$tool.log | tr '\n' ' ')"
done
ESTORCH_B200_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench_n1.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $out/${tag}_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:eval_mlp_f16_kernel|rank_grad_kernel' -s 2 -c 2 -f \
    -o $out/prof_$tag python tools/profile_kernels.py --pairs 2048 --iters 2 > $out/${tag}_ncu_full.log 2>&1
tail -2 $out/${tag}_ncu_full.log
