"""Print the key figures of bench.py JSON lines: python tools/show.py gpurun_out/x.json ..."""
import json, sys
for f in sys.argv[1:]:
    for l in open(f):
        if not l.startswith("{"):
            continue
        d = json.loads(l)
        if "value" not in d:
            print(f, d); continue
        r = d.get("roofline", {})
        print(f, round(d["value"], 1), "pairs/s", round(d["ms_per_step"], 2), "ms/step frac", round(r.get("frac", 0), 3), "solver ms",
              round(r.get("avg_launch_ms", 0), 2), "e2e", round(d["e2e"]["value"], 1), d.get("clocks", {}).get("sm_mhz"), d.get("clocks", {}).get("reasons"))
        if "ms_per_pair_at_scale" in d:
            print("   scale", [round(x, 2) for x in d["ms_per_pair_at_scale"]])
            print("   wc   ", [round(x, 2) for x in d.get("ms_per_pair_warp_constants_at_scale", [])])
            print("   iter ", [round(x, 2) for x in d.get("ms_per_pair_iterations_at_scale", [])])
