#!/bin/bash
# One ncu capture of the edge kernel on a reduced workload (the full-size launch would be replayed ~40x).
set -e
CMD="python bench.py --steps 1 --warmup 1 --nw ${NW:-600} --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_edges_build -s 1 -c 1 -o gpurun_out/edge_prof $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/plain.log
