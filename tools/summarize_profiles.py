#!/usr/bin/env python
"""Condense what a GPU visit left in gpurun_out/ into small tracked files under profiles/ (round tag as argv[1])."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(P, exist_ok=True)
md = ["# %s profile summary\n" % tag]

bj = os.path.join(G, "bench.json")
if os.path.exists(bj):
    b = json.load(open(bj))
    json.dump(b, open(os.path.join(P, "%s_bench.json" % tag), "w"), indent=1)
    md.append("## bench.py (N=1, CUDA-event timing, not under a profiler)\n")
    md.append("* value %.1f meshes/s (%.2f ms/step, batch %d), e2e %.1f meshes/s, launches %d, clocks %s\n" % (
        b["value"], b["ms_per_step"], b["config"]["meshes_per_gpu"], b["e2e"]["value"], b["gpu_launches"], b.get("clocks")))
    for c in ("c2", "c5"):
        fn = os.path.join(G, "bench_%s.json" % c)
        if os.path.exists(fn):
            try:
                bc = json.load(open(fn))
            except Exception:
                continue
            json.dump(bc, open(os.path.join(P, "%s_bench_%s.json" % (tag, c)), "w"), indent=1)
            rc = bc.get("roofline", {})
            md.append("* `--config %s` (%s): %.1f meshes/s (%.2f ms/step, batch %d), e2e %.1f; dominant family %.3f of the HBM "
                      "roofline, whole step %.3f\n" % (c, bc["config"]["workload"][:60], bc["value"], bc["ms_per_step"],
                                                        bc["config"]["meshes_per_gpu"], bc["e2e"]["value"], rc.get("frac", 0),
                                                        rc.get("whole_step_frac_of_hbm_roofline", 0)))
    r = b.get("roofline")
    if r:
        md.append("* dominant family `%s`: %.0f GB/s algorithmic = %.3f of %s %.0f GB/s; whole step = %.3f of the HBM roofline\n" % (
            r["kernel"], r["achieved"], r["frac"], r["peak_source"], r["peak"], r["whole_step_frac_of_hbm_roofline"]))
        md.append("\n| family | launches/step | ms/step | algorithmic GB/step | GB/s |\n|---|---:|---:|---:|---:|\n")
        for k, v in r["families"].items():
            md.append("| %s | %d | %.2f | %.2f | %.0f |\n" % (k, v["launches"], v["ms"], v["alg_GB"], v["GBps"]))
    if "cpu_baseline" in b:
        md.append("* cpu_baseline: %s\n" % b["cpu_baseline"])

lp = os.path.join(G, "launch_profile.json")
if os.path.exists(lp):
    rows = json.load(open(lp))
    rows.sort(key=lambda r: -r["us"])
    tot = sum(r["us"] for r in rows)
    md.append("\n## per-launch CUDA-event times of one eager step (top 25 of %d tagged launches, %.1f ms)\n\n" % (len(rows), tot / 1e3))
    md.append("| family | launch | us | algorithmic GB/s | share |\n|---|---|---:|---:|---:|\n")
    for r in rows[:25]:
        md.append("| %s | %s | %.0f | %.0f | %.1f%% |\n" % (r["family"], r["launch"], r["us"], r["GBps"] or 0, 100 * r["us"] / tot))
    with open(os.path.join(P, "%s_launch_profile.csv" % tag), "w") as f:
        w = csv.writer(f)
        w.writerow(["family", "launch", "alg_bytes", "us", "GBps"])
        for r in rows:
            w.writerow([r["family"], r["launch"], int(r["alg_bytes"]), "%.1f" % r["us"], "%.1f" % (r["GBps"] or 0)])

lc = os.path.join(G, "launches.csv")
if os.path.exists(lc):
    rows = list(csv.reader(open(lc)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv, mn, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("cape::", "").replace("<unnamed>::", "")
        v = float(r[mv].replace(",", ""))
        v = v / 1e3 if r[mu] == "ns" else (v * 1e3 if r[mu] == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    md.append("\n## ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`, eager steps; cold-cache, serialised: compare shares)\n\n")
    md.append("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
    with open(os.path.join(P, "%s_ncu_launches_by_kernel.csv" % tag), "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "share"])
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, v[0], "%.1f" % v[1], "%.4f" % (v[1] / tot)])
            if v[1] / tot > 0.003:
                md.append("| `%s` | %d | %.0f | %.1f%% |\n" % (k[:70], v[0], v[1], 100 * v[1] / tot))

WANT = ["Kernel Name", "launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
traffic = {}
for rep in sorted(f for f in os.listdir(G) if f.endswith(".ncu-rep")):
    out = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    md.append("\n## ncu --set full: %s\n\n| " % rep + " | ".join(hdr[i].split(".")[0].replace("__", " ") for i in idx) + " |\n|" + "---|" * len(idx) + "\n")
    with open(os.path.join(P, "%s_%s.csv" % (tag, rep.replace(".ncu-rep", ""))), "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
            try:      # DRAM bytes per launch of the conv kernels -> bench.py's roofline.traffic
                if any(k in r[hdr.index("Kernel Name")] for k in ("ellconv_tc", "gemm_tc", "apply_kernel")):   # the conv family
                    b = 0.0
                    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                        k = hdr.index(key)
                        b += float(r[k].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[k]]
                    traffic.setdefault("ellconv", []).append(b)
            except (ValueError, KeyError):
                pass
            md.append("| " + " | ".join(re.sub(r"\(.*", "", r[i])[-48:] + (" " + units[i] if units[i] else "") for i in idx) + " |\n")
if traffic:
    json.dump({"ellconv": sum(traffic["ellconv"]) / len(traffic["ellconv"]),
               "source": "profiles/%s_prof_*.csv" % tag,
               "note": "mean dram__bytes_read.sum + dram__bytes_write.sum per launch over the %d conv-family launches (fused, "
                       "plain-operand and apply kernels) captured with ncu --set full; bench.py reports it as roofline.traffic"
                       % len(traffic["ellconv"])},
              open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
open(os.path.join(P, "%s_summary.md" % tag), "w").write("".join(md))
print("".join(md)[:3000])
