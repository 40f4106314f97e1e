import json, sys
d = json.load(open(sys.argv[1]))
print("ms/step %.3f  utt/s %.0f  e2e utt/s %.0f (%.3f ms)  launches/step %.0f  clocks %s" % (
    d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"] / d["steps"], d["clocks"]))
if d.get("roofline"): print("roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["roofline"].items()})
if d.get("cpu_baseline"): print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
for k, v in (d.get("kernel_profile") or {}).items():
    print("  %-22s n=%3d %8.1fus share %5.1f%%" % (k, v["launches_per_step"], v["us_per_launch"], v["share"] * 100))
