export LAS_BENCH_MIN_WARMUP=1
timeout -s KILL 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-greedy > gpurun_out/plain.log 2>&1 || exit 1
timeout -s KILL 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 6300 -c 6400 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-greedy > gpurun_out/ncu.log 2>&1; echo "ncu list rc=$?"
timeout -s KILL 120 python scripts/ncu_targets.py > gpurun_out/ncu_targets_plain.log 2>&1 && timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_tc_kernel|attn_step_split_kernel|lstm_rec_fwd_dsm_kernel|lstm_rec_bwd_dsm_kernel|lstm_rec_fwd_tc_kernel|lstm_rec_bwd_tc2_kernel" -o gpurun_out/prof_r1_final python scripts/ncu_targets.py > gpurun_out/ncu_full3.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full3.log
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r1_final.csv
