export GRR_LIB=$PWD/globalredundancyresolution_b200/libgrr_var1.so
for cfg in "jaco baseline_cfg3.json" "baxter baseline_cfg4.json" "robonaut2 baseline_cfg5.json Nw=20000"; do
  timeout 300 python tools/time_cfg.py $cfg 2>&1 | tail -1 | cut -c1-260
done
unset GRR_LIB
timeout 300 python tools/time_cfg.py robonaut2 baseline_cfg5.json Nw=20000 2>&1 | tail -1 | cut -c1-260
