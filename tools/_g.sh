mkdir -p gpurun_out/r2n
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2n/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n/pytest.log
tail -6 gpurun_out/r2n/pytest.log
# launch list of the default bench command (reduced: the profiler serialises and replays)
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-extras"
$CMD > gpurun_out/r2n/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n/launches_fullsize.csv $CMD > gpurun_out/r2n/ncu_launches.log 2>&1
tail -2 gpurun_out/r2n/ncu_launches.log
# full-size capture of the dominant kernel (traffic for the roofline entry)
ncu --set full --clock-control none --import-source on -k regex:k_edges_build -s 1 -c 1 -o gpurun_out/r2n/edge_exact_fullsize $CMD > gpurun_out/r2n/ncu_full.log 2>&1
tail -2 gpurun_out/r2n/ncu_full.log
