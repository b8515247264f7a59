#!/bin/bash
# One ncu capture of the edge kernel on a reduced workload (the full-size launch would be replayed ~40x).
# usage: [NW=600] [EXTRA="--problem robonaut2/baseline_cfg5.json --nq 100"] [OUT=edge_prof] tools/ncu_edge.sh
set -e
CMD="python bench.py --steps 1 --warmup 1 --nw ${NW:-600} --no-cpu ${EXTRA:-}"
OUT=${OUT:-edge_prof}
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${OUT}.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_edges_build -s 1 -c 1 -o gpurun_out/${OUT} $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/plain.log
