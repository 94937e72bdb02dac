# Round-2 profile pass (run on the GPU box): each program first runs WITHOUT ncu and must exit 0.
set -x
mkdir -p gpurun_out
timeout -s KILL 200 python scripts/ncu_step.py 2 > gpurun_out/ncu_step_plain.log 2>&1 || { tail -5 gpurun_out/ncu_step_plain.log; exit 1; }
tail -1 gpurun_out/ncu_step_plain.log
# launch list of exactly one train step, backward overlap on (its default, also under the profiler)
timeout -s KILL 420 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_r2_overlap.csv python scripts/ncu_step.py 2 > gpurun_out/ncu_step_overlap.log 2>&1; echo "ncu overlap-on rc=$?"
tail -2 gpurun_out/ncu_step_overlap.log
# the same with the default (overlap switched off under an attached profiler): the serial schedule
timeout -s KILL 420 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_r2_serial.csv python scripts/ncu_step.py 2 > gpurun_out/ncu_step_serial.log 2>&1; echo "ncu serial rc=$?"
tail -2 gpurun_out/ncu_step_serial.log
timeout -s KILL 200 python scripts/ncu_targets.py > gpurun_out/ncu_targets_plain.log 2>&1 && timeout -s KILL 900 ncu --set full --clock-control none --import-source on \
    -k regex:"gemm_bf16_tc_kernel|attn_step_split_kernel|lstm_rec_fwd_dsm_kernel|lstm_rec_bwd_dsm_kernel|dec_persist_fwd_kernel" \
    -o gpurun_out/prof_r2 python scripts/ncu_targets.py > gpurun_out/ncu_full_r2.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full_r2.log
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r2_*.csv
