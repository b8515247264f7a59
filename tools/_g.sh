mkdir -p gpurun_out/r2m
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2m/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m/pytest.log
tail -8 gpurun_out/r2m/pytest.log
timeout 900 python bench.py > gpurun_out/r2m/bench.json 2> gpurun_out/r2m/bench.err; tail -c 300 gpurun_out/r2m/bench.json; tail -3 gpurun_out/r2m/bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m/smoke.log 2>&1; tail -3 gpurun_out/r2m/smoke.log
