mkdir -p gpurun_out/r2a
python tools/exact_report.py planar_20R baseline_cfg2.json scales=1,4,8 > gpurun_out/r2a/exact_cfg2.jsonl 2> gpurun_out/r2a/exact_cfg2.err
python tools/exact_report.py planar_3R baseline_cfg1.json scales=1,8 > gpurun_out/r2a/exact_cfg1.jsonl 2> gpurun_out/r2a/exact_cfg1.err
python tools/exact_report.py jaco baseline_cfg3.json scales=1,8 > gpurun_out/r2a/exact_cfg3.jsonl 2> gpurun_out/r2a/exact_cfg3.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a/smoke.log 2>&1
python bench.py > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err
tail -3 gpurun_out/r2a/pytest.log; cat gpurun_out/r2a/exact_cfg2.jsonl | cut -c1-600
