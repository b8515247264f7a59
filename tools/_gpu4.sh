mkdir -p gpurun_out/r2d
python tools/time_cfg.py planar_20R baseline_cfg2.json Nw=1500 > gpurun_out/r2d/plain_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_edges_build -s 1 -c 1 -o gpurun_out/r2d/flag_cfg2 python tools/time_cfg.py planar_20R baseline_cfg2.json Nw=1500 > gpurun_out/r2d/ncu_flag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_edges_build -s 1 -c 1 -o gpurun_out/r2d/plain_cfg2 python tools/time_cfg.py planar_20R baseline_cfg2.json Nw=1500 exact=0 > gpurun_out/r2d/ncu_plain.log 2>&1
python tools/exact_report.py planar_20R baseline_cfg2.json scales=8 > gpurun_out/r2d/exact_cfg2.jsonl 2> gpurun_out/r2d/exact_cfg2.err
cut -c1-900 gpurun_out/r2d/exact_cfg2.jsonl; ls -la gpurun_out/r2d; tail -3 gpurun_out/r2d/plain_run.log
