#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "rank or sharded or adam or grad or ns or novelty or north_star" 2>&1 | tail -3
timeout 200 python tools/rank_grad_time.py new 2>&1 | grep -v Warn | tee $out/r02j_rank_grad_time.txt
