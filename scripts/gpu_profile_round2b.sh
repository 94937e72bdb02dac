# Round-2 final profile pass (run on the GPU box): each program first runs WITHOUT ncu and must exit 0.
set -x
mkdir -p gpurun_out
timeout -s KILL 200 python scripts/ncu_step.py 2 > gpurun_out/ncu_step_plain.log 2>&1 || { tail -5 gpurun_out/ncu_step_plain.log; exit 1; }
tail -1 gpurun_out/ncu_step_plain.log
# launch list of exactly one train step with the shipped schedule (backward overlap, forward / backward pipelining behind progress counters)
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_r2_final.csv python scripts/ncu_step.py 2 > gpurun_out/ncu_step_final.log 2>&1; echo "ncu launch list rc=$?"
tail -2 gpurun_out/ncu_step_final.log
timeout -s KILL 300 python scripts/ncu_targets.py > gpurun_out/ncu_targets_plain.log 2>&1 || { tail -5 gpurun_out/ncu_targets_plain.log; exit 1; }
NCU_CAPTURE=1 timeout -s KILL 1200 ncu --set full --clock-control none --import-source on \
    -k regex:"gemm_bf16_tc_kernel|attn_step_split_kernel|lstm_rec_fwd_dsm_kernel|lstm_rec_bwd_dsm_kernel|dec_persist_fwd_kernel" \
    -o gpurun_out/prof_r2_final python scripts/ncu_targets.py > gpurun_out/ncu_full_r2_final.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full_r2_final.log
# the report itself is larger than what gpurun copies back: extract the raw page here, keep the CSVs
ncu -i gpurun_out/prof_r2_final.ncu-rep --page raw --csv > gpurun_out/ncu_raw_r2_final.csv 2> gpurun_out/ncu_raw_err.log
python scripts/extract_ncu_full.py "ncu --set full --clock-control none, scripts/ncu_targets.py (round 2 final): gate GEMM fwd / dgrad / wgrad, attention step fwd (train + greedy shape), DSMEM recurrence fwd / BPTT, persistent decoder kernel (fp16 K/V, tensor-core attention passes), decoder backward step (attn_step_split_kernel<1,0,1> = attention backward + dq.Wq + cell-1 backward; its small tcgen05 GEMMs)" < gpurun_out/ncu_raw_r2_final.csv > gpurun_out/ncu_full_r2_final_kernels.csv
ncu -i gpurun_out/prof_r2_final.ncu-rep --page source --csv --kernel-name regex:dec_persist_fwd_kernel > gpurun_out/ncu_source_dec_persist.csv 2>> gpurun_out/ncu_raw_err.log
rm -f gpurun_out/prof_r2_final.ncu-rep
ls -la gpurun_out/ | tail -12; du -sh gpurun_out
