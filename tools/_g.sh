mkdir -p gpurun_out/r2p
timeout 600 python tools/exact_report.py planar_20R baseline_cfg2.json scales=8 > gpurun_out/r2p/exact.jsonl 2> gpurun_out/r2p/exact.err; cut -c1-900 gpurun_out/r2p/exact.jsonl; tail -3 gpurun_out/r2p/exact.err
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_exact.py -x -q 2>&1 | tail -4
