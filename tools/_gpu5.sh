mkdir -p gpurun_out/r2e
for cfg in "planar_20R baseline_cfg2.json" "jaco baseline_cfg3.json" "baxter baseline_cfg4.json" "robonaut2 baseline_cfg5.json Nw=30000"; do
  tag=$(echo $cfg | cut -d' ' -f1)
  timeout 300 python tools/exact_report.py $cfg scales=8 > gpurun_out/r2e/exact_$tag.jsonl 2> gpurun_out/r2e/exact_$tag.err
done
cut -c1-800 gpurun_out/r2e/*.jsonl
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e/pytest.log
tail -15 gpurun_out/r2e/pytest.log
timeout 600 python bench.py > gpurun_out/r2e/bench.json 2> gpurun_out/r2e/bench.err; tail -c 3000 gpurun_out/r2e/bench.json; tail -5 gpurun_out/r2e/bench.err
