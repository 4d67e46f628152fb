"""Print the key numbers of a bench.py JSON line read from stdin (development aid)."""
import json
import sys

d = json.loads(sys.stdin.read().strip().split("\n")[-1])
tag = sys.argv[1] if len(sys.argv) > 1 else ""
r = d.get("roofline") or {}
print(tag, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "conv TF/s", round(r.get("achieved", 0), 1),
      "frac", round(r.get("frac", 0), 3), "by_bn", {k: round(v["tflops"]) for k, v in (r.get("by_tile_width") or {}).items()},
      "hbm", [(h["kernel"][:12], round(h["achieved"]), round(h["frac"], 3)) for h in d.get("roofline_hbm") or []],
      "post_us", round((d.get("postprocess") or {}).get("us_per_step", 0), 1), "clk", (d.get("clocks") or {}).get("sm_mhz"))
