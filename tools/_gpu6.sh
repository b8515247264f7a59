mkdir -p gpurun_out/r2f
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2f/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f/pytest.log
tail -15 gpurun_out/r2f/pytest.log
timeout 900 python bench.py > gpurun_out/r2f/bench.json 2> gpurun_out/r2f/bench.err; tail -c 4000 gpurun_out/r2f/bench.json; tail -5 gpurun_out/r2f/bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f/smoke.log 2>&1; tail -3 gpurun_out/r2f/smoke.log
