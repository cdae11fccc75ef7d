"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel for the LAST step
(from the last launch of `anchor`).   python scripts/summarize_launches.py file.csv [anchor]"""
import collections
import csv
import sys

path = sys.argv[1]
anchor = sys.argv[2] if len(sys.argv) > 2 else "embed_drop_kernel"
rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
idx = [i for i, r in enumerate(rows) if anchor in r[4]]
last = rows[idx[-1]:] if idx else rows
agg = collections.OrderedDict()
for r in last:
    name = r[4].split("(")[0].replace("void ", "").replace("roko::", "")
    if "at::" in name:
        name = "torch:" + name.split("<")[0][-40:]
    agg.setdefault(name, [0, 0.0])
    agg[name][0] += 1
    agg[name][1] += float(r[-1]) / 1e3
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{v[1]:9.1f} us {100 * v[1] / tot:5.1f}% {v[0]:3d}x  {k[:90]}")
print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
