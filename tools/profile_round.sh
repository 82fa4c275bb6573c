#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list of three eager steps, ncu --set full of the hot kernels.
# Usage (from the repo root): bash tools/profile_round.sh <tag>     -> gpurun_out/*_<tag>.*
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $out/pytest_gpu_$tag.log
timeout 400 python bench.py --steps 40 --warmup 5 > $out/bench_$tag.json 2> $out/bench_err_$tag.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_step_$tag.csv \
  python tools/prof_step.py 4096 3 > $out/prof_step_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'skh_persist2|skh_final_tile|gemm_tf32|procr_pose|match_from_best|prep_operand' \
  -s 9 -c 9 -f -o $out/prof_step_full_$tag python tools/prof_step.py 4096 2 > $out/prof_full_$tag.log 2>&1
tail -3 $out/pytest_gpu_$tag.log
cat $out/bench_$tag.json
tail -2 $out/prof_full_$tag.log
