import json, sys
d = json.load(open(sys.argv[1]))
r = d["roofline"]
print(f"value {d['value']:.4g} {d['unit']}  ms/step {d['ms_per_step']:.3f}  hbm {r['achieved']:.0f} GB/s ({r['frac']:.3f})  "
      f"fp32 {r['fp32']['achieved_tflops']:.1f} TF ({r['fp32']['frac']:.3f})  t_min/t {r['t_min_over_t']:.3f} [{r['t_min_bound']}]  {r['kernel']}")
print("e2e", f"{d['e2e']['value']:.4g}", "clocks", d["clocks"])
if "cpu_baseline" in d:
    print("cpu", d["cpu_baseline"])
for k, v in d.get("workloads", {}).items():
    print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
