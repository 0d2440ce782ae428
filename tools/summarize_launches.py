#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) of a
bench.py run into one step's kernel table: launches per step, time share, DRAM bytes.  Writes a markdown table and the
JSON bench.py reads for roofline.traffic."""
import csv
import json
import sys
from collections import OrderedDict

src, out_md, out_json = sys.argv[1], sys.argv[2], sys.argv[3]
slots = int(sys.argv[4]) if len(sys.argv) > 4 else 512
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
iid, ik, im, iv = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
launch = OrderedDict()
for r in rows:
    d = launch.setdefault(int(r[iid]), {"name": r[ik]})
    d[r[im]] = float(r[iv])
seq = list(launch.values())


def short(n):
    n = n.replace("void ", "").replace("unnamed>::", "")
    return n.split("(")[0]


# one step = the launches between two consecutive k_pfb launches
idx = [i for i, l in enumerate(seq) if short(l["name"]).startswith("k_pfb<")]
assert len(idx) >= 2, "need at least two steps in the capture"
step = seq[idx[0]:idx[1]]
agg = OrderedDict()
for l in step:
    a = agg.setdefault(short(l["name"]), {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0})
    a["n"] += 1
    a["ns"] += l.get("gpu__time_duration.sum", 0)
    a["rd"] += l.get("dram__bytes_read.sum", 0)
    a["wr"] += l.get("dram__bytes_write.sum", 0)
tot_ns = sum(a["ns"] for a in agg.values())
tot_b = sum(a["rd"] + a["wr"] for a in agg.values())
with open(out_md, "w") as f:
    f.write("| kernel | launches | time (ms, under ncu) | share | DRAM read MB | DRAM write MB |\n|---|---|---|---|---|---|\n")
    for k, a in agg.items():
        f.write("| `%s` | %d | %.3f | %.1f %% | %.1f | %.1f |\n" % (k, a["n"], a["ns"] / 1e6, 100 * a["ns"] / tot_ns, a["rd"] / 1e6, a["wr"] / 1e6))
    f.write("| **step** | %d | %.3f | 100 %% | %.1f | %.1f |\n" % (sum(a["n"] for a in agg.values()), tot_ns / 1e6,
                                                                    sum(a["rd"] for a in agg.values()) / 1e6, sum(a["wr"] for a in agg.values()) / 1e6))
json.dump({"slots": slots, "ddc": "poly", "workload": "sniffer", "dram_bytes_per_step": int(tot_b),
           "kernels": {k: {"launches": a["n"], "ms": a["ns"] / 1e6, "dram_bytes": int(a["rd"] + a["wr"])} for k, a in agg.items()},
           "source": src}, open(out_json, "w"), indent=1)
print(open(out_md).read())
