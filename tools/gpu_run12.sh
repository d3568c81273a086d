#!/bin/bash
out=gpurun_out; mkdir -p $out
ESTK_LIBRARY=estorch_b200/lib/libestk_rcbar.so timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tests/sanitizer_pass.py > $out/r02_sanitizer_racecheck_groupbarrier.log 2>&1
grep -E "RACECHECK SUMMARY|sanitizer_pass ok" $out/r02_sanitizer_racecheck_groupbarrier.log
