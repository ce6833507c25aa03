"""profiles/traffic.json from committed ncu csv launch lists (dram__bytes_read.sum + dram__bytes_write.sum per launch).
bench.py reads the json for `roofline.traffic` instead of carrying literals.

    python tools/ncu_traffic.py --stage2 profiles/r02_stage2_ncu.csv --attention profiles/r02_attention_ncu.csv
"""
import argparse, collections, csv, json, os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    d = collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault((int(r[ii]), r[ki]), {})[r[mi]] = float(r[vi].replace(",", ""))
    return d


p = argparse.ArgumentParser()
p.add_argument("--stage2")
p.add_argument("--attention")
a = p.parse_args()
out_path = os.path.join(ROOT, "profiles", "traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
if a.stage2:
    d = launches(a.stage2)
    # one iteration = the launches from one uvt_gather to the next
    keys = list(d.keys())
    starts = [i for i, k in enumerate(keys) if "uvt_gather" in k[1]]
    it = keys[starts[0]:starts[1]] if len(starts) > 1 else keys[starts[0]:]
    # the Adam of that iteration is launched right after level0 / the loss assembly: include kernels up to the next gather
    per = [{"kernel": k[1].split("(")[0], "us": d[k]["gpu__time_duration.sum"] / 1e3,
            "dram_bytes": d[k]["dram__bytes_read.sum"] + d[k]["dram__bytes_write.sum"]} for k in it]
    out["stage2_iteration"] = {"dram_bytes_per_launch": sum(x["dram_bytes"] for x in per), "kernel_us_sum": sum(x["us"] for x in per),
                               "kernels": per, "source": os.path.relpath(a.stage2, ROOT)}
if a.attention:
    d = launches(a.attention)
    best = max(d.items(), key=lambda kv: kv[1]["gpu__time_duration.sum"])
    out["attention_ds1"] = {"dram_bytes_per_launch": best[1]["dram__bytes_read.sum"] + best[1]["dram__bytes_write.sum"],
                            "us": best[1]["gpu__time_duration.sum"] / 1e3, "kernel": best[0][1].split("(")[0],
                            "source": os.path.relpath(a.attention, ROOT)}
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1)[:1500])
