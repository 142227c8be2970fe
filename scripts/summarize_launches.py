"""ncu launch-list CSV (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch) -> per-kernel table
of ONE step (the last `--per-step` launches of the capture) as JSON + markdown.

  python scripts/summarize_launches.py profiles/r01_launches_X.csv --per-step 166 --out profiles/r01_launch_summary.json
"""
import argparse, collections, csv, json, re

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--per-step", type=int, default=0)
ap.add_argument("--out")
ap.add_argument("--source", default="")
a = ap.parse_args()

rows = []
with open(a.csv, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
per = collections.OrderedDict()
for r in rd:
    kid = int(r["ID"])
    d = per.setdefault(kid, {"kernel": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    name = r["Metric Name"]
    if name == "gpu__time_duration.sum":
        d["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    elif name.startswith("dram__bytes"):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        d["dram"] = d.get("dram", 0.0) + v * mult
allk = list(per.values())
if a.per_step > 0:
    launches = allk[-a.per_step:]
else:  # one forward + scoring: the library's own kernels from the last embedding kernel on (the visual encoder of the same step
    # may have been enqueued just before it on the side stream)
    i0 = max(i for i, l in enumerate(allk) if "embed_ln_kernel" in l["kernel"])
    if i0 > 0 and "visn_embed_kernel" in allk[i0 - 1]["kernel"]:
        i0 -= 1
    launches = [l for l in allk[i0:] if "rgqa::" in l["kernel"]]
tot = sum(l["us"] for l in launches)
agg = collections.OrderedDict()
for l in launches:
    k = re.sub(r"^void\s+", "", l["kernel"])
    k = re.sub(r"\(.*$", "", k).replace("rgqa::<unnamed>::", "").replace("(anonymous namespace)::", "")
    g = agg.setdefault(k, {"kernel": k, "launches": 0, "us": 0.0, "dram_bytes": 0.0})
    g["launches"] += 1
    g["us"] += l["us"]
    g["dram_bytes"] += l.get("dram", 0.0)
ks = sorted(agg.values(), key=lambda g: -g["us"])
for g in ks:
    g["share"] = round(g["us"] / tot, 4)
    g["us"] = round(g["us"], 1)
gem = [g for g in ks if any(t in g["kernel"] for t in ("gemm_bf16", "qkv_attn_pair", "visn_embed"))]  # every tcgen05 GEMM-class launch
out = {"source": a.source, "step_kernel_time_us": round(tot, 1), "launches": len(launches), "kernels": ks,
       "tcgen05_gemm": {"launches": sum(g["launches"] for g in gem), "us": round(sum(g["us"] for g in gem), 1),
                        "share": round(sum(g["us"] for g in gem) / tot, 4),
                        "dram_bytes_per_launch": sum(g["dram_bytes"] for g in gem) / max(1, sum(g["launches"] for g in gem))},
       "dram_bytes_per_step": sum(g["dram_bytes"] for g in ks)}
if a.out:
    json.dump(out, open(a.out, "w"), indent=1)
print(f"one step: {tot:.1f} us of kernel time, {len(launches)} launches, {out['dram_bytes_per_step']/1e9:.1f} GB DRAM")
print("| kernel | launches | us | share | DRAM MB / launch | DRAM GB/s |\n|---|---|---|---|---|---|")
for g in ks:
    print(f"| `{g['kernel']}` | {g['launches']} | {g['us']:.0f} | {100*g['share']:.1f} % | {g['dram_bytes']/g['launches']/1e6:.0f} | {g['dram_bytes']/g['us']/1e3:.0f} |")
