#!/usr/bin/env python
"""Per-layer roofline table of the nz64 train step from a launch profile (profiles/<tag>_launch_profile.csv):
measured launch time against the layer's HBM floor (algorithmic bytes) and its tensor floor (3xTF32).
    python tools/layer_table.py r02b   ->  profiles/r02b_layer_table.md"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02b"
rows = list(csv.DictReader(open(os.path.join(ROOT, "profiles", "%s_launch_profile.csv" % tag))))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
    else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
PEAK_HBM = peaks["hbm_gbs"] * 1e9
TF32 = 0.5 * peaks["bf16_tflops_sustained"] * 1e12       # tf32 runs at half the bf16 rate; three passes for fp32 accuracy

# layer geometry of the nz64 model: rows per mesh, Fin incl. condition channels, Fout, K, affine
geo = {}
M, F = [6890, 6890, 3445, 3445, 1723, 1723, 862, 862, 862], [64, 64, 128, 128, 256, 256, 512, 512]
fin = 3
for i in range(8):
    geo["enc/conv%d" % (i + 1)] = (M[i], fin, F[i], 2, False)
    fin = F[i]
geo["enc/1x1"], geo["dec/1x1"] = (862, 512, 64, 1, False), (862, 64, 512, 1, False)
fin, lev = 512, [862, 862, 1723, 1723, 3445, 3445, 6890, 6890]
for i in range(8):
    geo["dec/aff%d" % (i + 1)] = (lev[i], fin + 64, F[-i - 1] // 2, 2, True)
    fin = F[-i - 1] // 2
geo["dec/outputs"] = (6890, fin + 64, 3, 2, False)
fin, Md = 67, [6890, 3445, 1723, 862, 431]
for i in range(4):
    geo["disc/conv%d" % (i + 1)] = (Md[i], fin, F[i], 3, False)
    fin = F[i]
geo["disc/pred_map"] = (431, 128, 1, 2, False)

agg = collections.OrderedDict()
for r in rows:
    if ":" not in r["launch"]:
        continue
    layer, p = r["launch"].split(":")
    a = agg.setdefault((layer, p.split("/")[0]), [0.0, 0.0, 0])
    a[0] += float(r["us"]); a[1] += float(r["alg_bytes"]); a[2] += 1
out = ["# Per-layer roofline table, nz64 train step at batch 64 (from `%s_launch_profile.csv`; eager step, CUDA events per launch)\n\n" % tag,
       "HBM floor = algorithmic bytes (SURVEY.md 8d) / %.0f GB/s (measured copy bandwidth).  Tensor floor = 3 x FLOPs / %.0f "
       "TFLOP/s (3xTF32 for fp32 accuracy, tf32 at half the measured sustained bf16 rate): the path is co-limited, `frac` = "
       "max(floor) / measured.  dW rows: all weight-gradient launches of the layer.  Discriminator rows: 2N meshes in the "
       "forward and the D-loss backward, N more in the G-loss backward (summed).\n\n" % (PEAK_HBM / 1e9, TF32 / 1e12),
       "| layer | pass | launches | us | alg MB | GB/s | HBM floor us | tensor floor us | frac |\n|---|---|---:|---:|---:|---:|---:|---:|---:|\n"]
tot = [0.0, 0.0, 0.0]
for (layer, p), (us, b, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    fl = 0.0
    if layer in geo:
        Mo, Fi, Fo, K, aff = geo[layer]
        Nn = 64
        if layer.startswith("disc"):
            Nn = 128 if p in ("fwd", "dW") else 192
        fl = 2.0 * Nn * Mo * Fi * K * Fo * ((K + 1.0) / K if aff else 1.0)
    hb, tf = b / PEAK_HBM * 1e6, 3 * fl / TF32 * 1e6
    out.append("| %s | %s | %d | %.0f | %.0f | %.0f | %.0f | %.0f | %.2f |\n" % (layer, p, n, us, b / 1e6, b / us / 1e3 if us else 0,
                                                                               hb, tf, max(hb, tf) / us if us else 0))
    tot[0] += us; tot[1] += hb; tot[2] += max(hb, tf)
out.append("\nSum of the measured launch times %.2f ms; sum of the HBM floors %.2f ms; sum of the per-layer max(HBM, tensor) floors "
           "%.2f ms: the step runs at %.2f of its co-limited floor (%.2f of the HBM-only floor).\n"
           % (tot[0] / 1e3, tot[1] / 1e3, tot[2] / 1e3, tot[2] / tot[0], tot[1] / tot[0]))
open(os.path.join(ROOT, "profiles", "%s_layer_table.md" % tag), "w").write("".join(out))
print(out[-1])
