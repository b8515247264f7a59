# A/B of two builds of the library (OLD = globalredundancyresolution_b200/libgrr_b200_old.so, built from another checkout) on the BASELINE configs: edge + sampling times, device events
run() { # label lib name json extra...
  lbl=$1; lib=$2; shift 2
  GRR_LIB=$lib REPS=3 timeout 600 python tools/time_cfg.py "$@" 2>&1 | tail -2 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print('$lbl', l.strip()[:200]); continue
    print('$lbl', d['cfg'], 't_sample', round(d['t_sample'],4), 't_connect', round(d['t_connect'],4), 'pairs', d['pairs'])"
}
OLD=$PWD/globalredundancyresolution_b200/libgrr_b200_old.so
NEW=$PWD/globalredundancyresolution_b200/libgrr_b200.so
for rep in 1 2; do
run old $OLD planar_20R baseline_cfg2.json
run new $NEW planar_20R baseline_cfg2.json
run old $OLD jaco baseline_cfg3.json
run new $NEW jaco baseline_cfg3.json
done
run old $OLD baxter baseline_cfg4.json
run new $NEW baxter baseline_cfg4.json
run old $OLD robonaut2 baseline_cfg5.json Nw=20000
run new $NEW robonaut2 baseline_cfg5.json Nw=20000
