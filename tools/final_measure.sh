#!/bin/bash
# The measurement sequence whose outputs are committed under profiles/ (run under gpurun on one B200): tools/final_measure.sh [tag]
TAG=${1:-r2}
set -x
if [ -z "$ONLY_NCU" ]; then
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || exit 1
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_reference_arm.json 2>> gpurun_out/${TAG}_bench.err
for c in 2 3 4; do
  python bench.py --config $c --steps 20 > gpurun_out/${TAG}_bench_config$c.json 2>> gpurun_out/${TAG}_bench.err
done
fi
K='regex:k_groups|k_canflux|k_init_timestep|k_snicar|k_coszen|k_phenology|k_atm|k_bareground'
# launch list of the default workload (2M columns), two timed steps
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-verify"
$CMD > gpurun_out/${TAG}_plain_launches.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 45 -c 30 --csv \
    --log-file gpurun_out/${TAG}_launches_2Mcols.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
# counters + full sections of one step at 524288 columns
CMD="python bench.py --ncols 524288 --steps 1 --warmup 3 --no-cpu-baseline --no-verify"
$CMD > gpurun_out/${TAG}_plain_512k.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k "$K" -s 45 -c 15 --csv --log-file gpurun_out/${TAG}_counters_512k.csv $CMD > gpurun_out/${TAG}_ncu_counters.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "$K" -s 45 -c 15 -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
python tools/show_bench.py < gpurun_out/${TAG}_bench.json
