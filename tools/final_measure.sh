#!/bin/bash
# The measurement sequence whose outputs are committed under profiles/ (run under gpurun on one B200).
set -x
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_reference.json 2>> gpurun_out/bench_final.err
# launch list of the default workload (2M columns), two timed steps
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_launches.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_groups|k_canflux|k_init" -s 24 -c 16 --csv \
    --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_launches.log 2>&1
# counters + full sections of one step at 524288 columns
CMD="python bench.py --ncols 524288 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_512k.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:"k_groups|k_canflux|k_init" -s 24 -c 8 --csv --log-file gpurun_out/counters_final.csv $CMD > gpurun_out/ncu_counters.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_groups|k_canflux|k_init" -s 24 -c 8 -o gpurun_out/prof_final -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python tools/show_bench.py < gpurun_out/bench_final.json
