import sys, json
lbl = sys.argv[1]
for l in sys.stdin:
    try:
        d = json.loads(l)
        print(lbl, d.get("mode", ""), d["cfg"], "t_sample", round(d["t_sample"], 4), "t_connect", round(d["t_connect"], 4), d.get("configs_bit_equal_to_levels", ""))
    except Exception:
        print(lbl, l.strip()[:400])
