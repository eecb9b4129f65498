"""ncu CSV (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum over one C3 prefill step) ->
profiles/r02_traffic.json: DRAM bytes per launch of the roofline kernels (what bench.py's roofline objects carry as `traffic`)
and a per-kernel share table.  usage: python scripts/make_traffic.py gpurun_out/r02_dram_step.csv profiles/r02_traffic.json"""
import collections, csv, json, sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mn, mv, mu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
idc = h.index("ID")
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    unit = r[mu]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
    d = per.setdefault(r[idc], {"name": r[kn]})
    d[r[mn]] = v * scale
agg = collections.defaultdict(lambda: {"launches": 0, "dram_bytes": 0.0, "ns": 0.0})
for d in per.values():
    name = d["name"]
    key = ("gemm" if "gemm_bf16_tcgen05" in name else "attn_causal" if ("attn_tc_kernel" in name and "true" in name.lower()) else
           "attn_causal" if ("attn_tc_kernel" in name and ", 1>" in name.replace("(bool)", "")) else "attn" if "attn_tc_kernel" in name else
           "hfre" if "hfre_sweep_mma" in name else "chan_gram" if "chan_gram" in name else "chan_apply" if "chan_apply" in name else
           name.split("(")[0].split("<")[0].replace("void ", "").replace("fo1::", ""))
    a = agg[key]
    a["launches"] += 1
    a["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a["ns"] += d.get("gpu__time_duration.sum", 0.0)
tot = sum(a["ns"] for a in agg.values())
out = {"source": sys.argv[1], "total_ms": tot / 1e6,
       "kernels": {k: {"launches": a["launches"], "ms": a["ns"] / 1e6, "share": a["ns"] / tot, "dram_bytes_per_launch": a["dram_bytes"] / a["launches"]}
                   for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"])}}
for k in ("gemm", "attn", "attn_causal", "hfre"):
    if k in agg:
        out[f"{k}_bytes_per_launch"] = agg[k]["dram_bytes"] / agg[k]["launches"]
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "kernels"}))
