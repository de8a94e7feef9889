#!/bin/bash
# Round-2 ncu captures (one gpurun call; each ncu run directly after the same command exited 0 without ncu).
#   bash tools/ncu_round2.sh [lv|generic|dmma|all]
set -u
what=${1:-all}
mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --no-ess --no-e2e --no-cpu --no-configs"
if [ "$what" = lv ] || [ "$what" = all ]; then
  $BENCH > gpurun_out/plain_bench.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
  $BENCH > gpurun_out/plain_bench2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:lv_mh_kernel -s 4 -c 2 -f -o gpurun_out/r02_lv $BENCH > gpurun_out/ncu_lv.log 2>&1
fi
if [ "$what" = generic ] || [ "$what" = all ]; then
  python tools/prof_generic.py > gpurun_out/plain_generic.log 2>&1 &&
  ncu --metrics sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
      --clock-control none -k regex:generic_mh_kernel -c 6 --csv --log-file gpurun_out/r02_generic_metrics.csv python tools/prof_generic.py > gpurun_out/ncu_generic.log 2>&1
fi
if [ "$what" = dmma ] || [ "$what" = all ]; then
  python tools/prof_big_linear.py > gpurun_out/plain_dmma.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:linear_dmma_mh_kernel -s 1 -c 2 -f -o gpurun_out/r02_dmma python tools/prof_big_linear.py > gpurun_out/ncu_dmma.log 2>&1
fi
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv 2>/dev/null
for f in gpurun_out/ncu_*.log; do echo "== $f"; tail -n 3 $f; done
du -sh gpurun_out
