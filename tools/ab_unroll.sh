# A/B of builds of the library with different GRR_F64_PLANAR_UNROLL (libgrr_b200_old.so = 2, libgrr_b200_uN.so = N): exact build of config 2, seeded sweep, config 1
t() { lbl=$1; shift; "$@" 2>&1 | tail -2 | python tools/ab_fmt.py "$lbl"; }
D=$PWD/globalredundancyresolution_b200
for rep in 1 2; do
for L in old u4 u5 u10; do
  case $L in old) LIB=$D/libgrr_b200_old.so;; u4) LIB=$D/libgrr_b200.so;; *) LIB=$D/libgrr_b200_$L.so;; esac
  t "$L exact" env GRR_LIB=$LIB REPS=3 timeout 300 python tools/time_cfg.py planar_20R baseline_cfg2.json
done; done
for L in old u4 u5 u10; do
  case $L in old) LIB=$D/libgrr_b200_old.so;; u4) LIB=$D/libgrr_b200.so;; *) LIB=$D/libgrr_b200_$L.so;; esac
  t "$L seeded" env GRR_LIB=$LIB timeout 300 python tools/time_seeded.py planar_20R baseline_cfg2.json
  t "$L cfg1" env GRR_LIB=$LIB REPS=3 timeout 300 python tools/time_cfg.py planar_3R baseline_cfg1.json
done
