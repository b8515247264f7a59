mkdir -p gpurun_out/r2c
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__icc_request_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,launch__registers_per_thread,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active
for cfg in "planar_20R baseline_cfg2.json" "baxter baseline_cfg4.json"; do
  tag=$(echo $cfg | cut -d' ' -f1)
  ncu --metrics $M --clock-control none -k regex:k_edges -c 8 --csv --log-file gpurun_out/r2c/ncu_${tag}_exact.csv python tools/time_cfg.py $cfg > gpurun_out/r2c/${tag}_exact.log 2>&1
  ncu --metrics $M --clock-control none -k regex:k_edges -c 8 --csv --log-file gpurun_out/r2c/ncu_${tag}_plain.csv python tools/time_cfg.py $cfg exact=0 > gpurun_out/r2c/${tag}_plain.log 2>&1
done
ls -la gpurun_out/r2c
