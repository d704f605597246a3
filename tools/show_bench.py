"""Readable digest of a bench.py JSON line:  python tools/show_bench.py gpurun_out/bench.json"""
import json, sys
l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: l[k] for k in ("n_gpus", "value", "ms_per_step", "ms_step_min_median_max", "gpu_launches") if k in l})
r = l["roofline"]
print("roofline", {k: r[k] for k in ("achieved", "frac", "t_min_over_t", "kernel")})
print("parity", {k: (l.get("parity") or {}).get(k) for k in ("ok", "max_rel_fro", "masks_equal", "users")})
e = l["e2e"]
print("e2e", {k: e.get(k) for k in ("value", "ceiling", "frac", "d2h_gb_per_s", "ceiling_gb_per_s")}, "default", (l.get("e2e_default") or {}).get("frac_of_pinned"))
if "cpu_baseline" in l:
    print("cpu", l["cpu_baseline"]["value"])
for k, v in (l.get("workloads") or {}).items():
    if "error" in v:
        print(k, v); continue
    keys = ("ms_per_step", "gb_per_s_per_gpu", "hbm_frac", "t_min_over_t", "t_min_bound", "mean_active_paths", "kernel", "coef_per_s")
    print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a in keys},
          "parity", {a: (v.get("parity") or {}).get(a) for a in ("ok", "max_rel_fro")})
