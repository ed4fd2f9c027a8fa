import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d["roofline"]
print("%.4g p-steps/s  %.3f ms/step  adv %.0f GB/s (%.3f) %.3f ms/launch  win %s" % (
    d["value"], d["ms_per_step"], r["achieved"], r["frac"], r["avg_launch_ms"],
    r.get("window_stats(gather_miss,deposit_miss,moves,rounds)")))
