#!/bin/bash
# profiles/run_ncu.sh -- the profiling recipe of /opt/skills/guides/B200_PROFILING.md for this repo.
# Run on the GPU box: gpurun -- 'bash profiles/run_ncu.sh <tag>'.  Outputs land in gpurun_out/.
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --sweeps 50 --em-n 200000 --em-maxit 3 --em-steps 1"
mkdir -p gpurun_out
# 1. the same command without ncu must exit 0 first
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$TAG.log; exit 1; }
# 2. every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
# 3. the two dominant kernels, full set, with source correlation
ncu --set full --clock-control none --import-source on -k regex:rj_sweep_kernel -s 2 -c 1 \
    -o gpurun_out/prof_rj_$TAG -f $CMD > gpurun_out/ncu_rj_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:em_fit_kernel -s 1 -c 1 \
    -o gpurun_out/prof_em_$TAG -f $CMD > gpurun_out/ncu_em_$TAG.log 2>&1
ls -la gpurun_out/
