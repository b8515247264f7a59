mkdir -p gpurun_out/r2b
python tools/exact_report.py planar_20R baseline_cfg2.json scales=8 > gpurun_out/r2b/exact_cfg2.jsonl 2> gpurun_out/r2b/exact_cfg2.err
python tools/exact_report.py jaco baseline_cfg3.json scales=8 > gpurun_out/r2b/exact_cfg3.jsonl 2> gpurun_out/r2b/exact_cfg3.err
python tools/exact_report.py baxter baseline_cfg4.json scales=8 > gpurun_out/r2b/exact_cfg4.jsonl 2> gpurun_out/r2b/exact_cfg4.err
python tools/exact_report.py robonaut2 baseline_cfg5.json scales=8 Nw=30000 > gpurun_out/r2b/exact_cfg5.jsonl 2> gpurun_out/r2b/exact_cfg5.err
cut -c1-900 gpurun_out/r2b/*.jsonl; tail -2 gpurun_out/r2b/*.err
