mkdir -p gpurun_out/r2k
export GRR_DUAL=1
CMD="python bench.py --steps 1 --warmup 1 --nw 1500 --no-cpu --no-extras"
$CMD > gpurun_out/r2k/plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_edges_build2 -s 1 -c 1 -o gpurun_out/r2k/dual_nw1500 $CMD > gpurun_out/r2k/ncu_full.log 2>&1
tail -2 gpurun_out/r2k/ncu_full.log
