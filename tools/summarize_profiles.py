#!/usr/bin/env python
"""Turns gpurun_out/*.ncu-rep + launch list into profiles/r1_ncu.md and profiles/r1_traffic.json."""
import csv, io, json, subprocess, sys
from collections import defaultdict

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"
rows = list(csv.reader(open("profiles/%s_launches.csv" % rnd)))
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
H = rows[hdr]; ki = H.index("Kernel Name"); vi = H.index("Metric Value")
d = defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > vi:
        d[r[ki]].append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in d.values())
md = ["# Round %s profiles (B200, sm_100a)\n" % rnd[1:], "## Launch list\n",
      "Command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file %s_launches.csv python bench.py --steps 2 --warmup 1 --no-alt --no-cpu`" % rnd,
      "(512 slots = 32 M samples per step, exact snr mode, stateless + lazy squelch; cold-cache, serialised: compare SHARES).",
      "Raw list: `profiles/%s_launches.csv`.  The channel FIR (`k_fir_packed<16, ...>`) appears once per batch on the device-"
      "resident path and in 4 parts (following the 4 parts of the input copy) on the host path, hence its launch count.\n" % rnd,
      "| kernel | launches | mean ms | share of GPU time |", "|---|---|---|---|"]
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    md.append("| `%s` | %d | %.3f | %.1f %% |" % (k[:90], len(v), sum(v) / len(v) / 1e6, 100 * sum(v) / tot))

keys = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg"]
traffic = {}
md += ["\n## `ncu --set full` summaries\n",
       "Commands: `ncu --set full --clock-control none --import-source on -k regex:\"k_fir_packed|k_fir_dl\" -s 4 -c 2 python tools/trace_collect.py exact` "
       "(whole-batch launches, 512 slots) and `... -k regex:\"k_demod|k_mm_|k_search|k_energy|k_gather\" -s 7 -c 7 python bench.py --steps 1 --warmup 1 --no-alt --no-cpu`.\n"]
for rep in sys.argv[2:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(out)))
    H, U, body = rr[0], rr[1], rr[2:]
    for r in body:
        name = r[H.index("Kernel Name")]
        md.append("### `%s`\n" % name)
        md.append("| metric | value | unit |\n|---|---|---|")
        for k in keys:
            if k in H:
                md.append("| %s | %s | %s |" % (k, r[H.index(k)], U[H.index(k)]))
        md.append("")
        def b(k):
            i = H.index(k)
            return float(r[i]) * UNIT.get(U[i], 1)
        t = int(b("dram__bytes_read.sum") + b("dram__bytes_write.sum"))
        for key, tag in (("k_fir_packed<16", "chan_fir"), ("k_fir_packed<2", "noise_fir"), ("k_fir_tiled<2", "noise_fir"), ("k_fir_dl", "noise_fir"),
                         ("k_mm_stateless", "demod_mm")):
            if key in name.replace("(int)", ""):
                traffic[tag] = t
json.dump({"slots": 512, "dram_bytes_per_launch": traffic,
           "source": "ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum per launch, see profiles/%s_ncu.md" % rnd},
          open("profiles/%s_traffic.json" % rnd, "w"), indent=1)
open("profiles/%s_ncu.md" % rnd, "w").write("\n".join(md))
print(traffic)
