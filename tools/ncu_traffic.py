"""profiles/r2_ncu_full_conv.csv + r2_ncu_full_hbm.csv (summaries written by tools/ncu_summary.py from `ncu --set full` of
tools/prof_step.py 12) -> profiles/r2_ncu_traffic.json: DRAM bytes per launch / per step and tensor-pipe activity per
kernel family, read by bench.py for its `traffic` fields.  Usage: python tools/ncu_traffic.py"""
import collections
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONV_PER_STEP = 38


def load(name):
    rows = list(csv.DictReader(open(os.path.join(ROOT, "profiles", name))))
    for r in rows:
        r["kernel"] = re.sub(r"^void\s+", "", r["kernel"]).strip()
    return rows


def family(k):
    m = re.match(r"(conv_gemm_\w+kernel)(<\d+)?", k)
    return (m.group(1) + (m.group(2) + ">" if m.group(2) else "")) if m else k


conv = load("r2_ncu_full_conv.csv")[:CONV_PER_STEP]                    # the first step of the capture
fam = collections.OrderedDict()
for r in conv:
    f = fam.setdefault(family(r["kernel"]), {"launches": 0, "time_us": 0.0, "tp": 0.0, "dram_MB": 0.0})
    t = float(r["time_us"])
    f["launches"] += 1
    f["time_us"] += t
    f["tp"] += t * float(r["tensor_pipe_active_pct"])
    f["dram_MB"] += float(r["dram_read_MB"]) + float(r["dram_write_MB"])
families = {k: {"launches": v["launches"], "time_us": round(v["time_us"], 1),
                "tensor_pipe_active_pct": round(v["tp"] / v["time_us"], 1), "dram_MB": round(v["dram_MB"], 1)}
            for k, v in fam.items()}
conv_bytes = sum(v["dram_MB"] for v in families.values()) * 1e6
tot_t = sum(v["time_us"] for v in families.values())
hbm = load("r2_ncu_full_hbm.csv")
# one step = the kernels from one canvas_clear to the next
first = [i for i, r in enumerate(hbm) if r["kernel"].startswith("canvas_clear")]
step = hbm[first[1]:first[2]] if len(first) > 2 else hbm[first[0]:first[1]]       # a steady-state step (sparse clear active)
fuse = sum(float(r["dram_read_MB"]) + float(r["dram_write_MB"]) for r in step if "warp_att" in r["kernel"]) * 1e6
front = sum(float(r["dram_read_MB"]) + float(r["dram_write_MB"]) for r in step if "warp_att" not in r["kernel"]) * 1e6
out = {
    "source": "profiles/r2_ncu_full_conv.csv, r2_ncu_full_hbm.csv (ncu --set full, tools/prof_step.py 12: 12 scenes x 5 agents "
              "per launch sequence, final kernels; regenerate with tools/ncu_traffic.py)",
    "scenes_per_step": 12,
    "conv_launches_per_step": CONV_PER_STEP,
    "conv_avg_dram_bytes_per_launch": conv_bytes / CONV_PER_STEP,
    "conv_dram_bytes_per_step": conv_bytes,
    "conv_time_us_per_step_under_ncu": round(tot_t, 1),
    "conv_tensor_pipe_active_pct_time_weighted": round(sum(v["time_us"] * v["tensor_pipe_active_pct"] for v in families.values()) / tot_t, 1),
    "fuse_dram_bytes_per_step": fuse,
    "front_dram_bytes_per_step": front,
    "hbm_kernels_of_the_step": [{"kernel": r["kernel"][:48], "time_us": float(r["time_us"]),
                                 "dram_MB": round(float(r["dram_read_MB"]) + float(r["dram_write_MB"]), 1)} for r in step],
    "conv_families": families,
}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "hbm_kernels_of_the_step"}, indent=1))
