#!/bin/bash
# profiles/run_ncu.sh -- the profiling recipe of /opt/skills/guides/B200_PROFILING.md for this repo.
# Run on the GPU box: gpurun -- 'bash profiles/run_ncu.sh <tag>'.  Outputs land in gpurun_out/.
# Each profiled command is first run plain (must exit 0), then under ncu.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu --sweeps 50 --em-n 200000 --em-maxit 3 --em-steps 1"
RJ="python profiles/rj_only.py 1048576 50 toy1"
EM="python profiles/em_only.py 1000000 2 10"
# 1. launch list of the bench command (cold-cache, serialised: compare SHARES, not absolutes)
$BENCH > gpurun_out/plain_bench_$TAG.log 2>&1 || { echo "plain bench failed"; tail -20 gpurun_out/plain_bench_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv $BENCH > gpurun_out/ncu_launches_$TAG.log 2>&1
# 2. the two dominant kernels, full set, with source correlation
$RJ > gpurun_out/plain_rj_$TAG.log 2>&1 || { echo "plain rj failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:rj_sweep_kernel -s 2 -c 1 \
    -o gpurun_out/prof_rj_$TAG -f $RJ > gpurun_out/ncu_rj_$TAG.log 2>&1
$EM > gpurun_out/plain_em_$TAG.log 2>&1 || { echo "plain em failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:em_fit -s 1 -c 1 \
    -o gpurun_out/prof_em_$TAG -f $EM > gpurun_out/ncu_em_$TAG.log 2>&1
cat gpurun_out/plain_rj_$TAG.log gpurun_out/plain_em_$TAG.log
ls -la gpurun_out/ | grep $TAG
