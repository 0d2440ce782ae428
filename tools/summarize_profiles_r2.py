#!/usr/bin/env python
"""profiles/r2_ncu.md from ncu --set full reports (raw page: headline metrics and stall ratios; source page: where the
warp samples fall between the block-wide barriers of each kernel = its phases)."""
import csv, io, subprocess, sys
from collections import Counter

out_md, reps = sys.argv[1], sys.argv[2:]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum"]
md = ["# Round 2 profiles (B200, sm_100a): `ncu --set full --clock-control none --import-source on`\n",
      "Command: `ncu --set full --clock-control none --import-source on -k regex:\"^k_pfb$|^k_nest$|k_mm_stateless_v2|k_search_warp\" -s 6 -c 4 "
      "python tools/poly_timing.py --slots 512 --iters 2` (throughput mode, 512 slots = 32 M samples per batch; per-launch times under ncu are "
      "cold-cache and serialised).  Launch list of the bench command itself: `profiles/r2_launches.csv`, one step of it: `profiles/r2_launches_step.md`.\n"]
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    H, U, body = rr[0], rr[1], rr[2:]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    secs, cur = [], None
    for r in csv.reader(io.StringIO(src)):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            secs.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    for li, r in enumerate(body):
        name = r[H.index("Kernel Name")]
        md.append("## `%s`\n" % name)
        md.append("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in H:
                md.append("| %s | %s | %s |" % (k, r[H.index(k)], U[H.index(k)]))
        st = [(float(r[i]), h.split("issue_stalled_")[1].split("_per_issue")[0]) for i, h in enumerate(H)
              if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and r[i]]
        st = [s for s in sorted(st, reverse=True) if s[0] >= 0.15 and s[1] not in ("selected",)]
        md.append("\nWarp stall reasons (warps per issue-active cycle): " + ", ".join("%s %.2f" % (n, v) for v, n in st) + ".\n")
        if li < len(secs):
            rows = secs[li]["rows"]
            h2, data = rows[0], rows[1:]
            iS, iI, iSrc = h2.index("# Samples"), h2.index("Instructions Executed"), h2.index("Source")
            tot, ti = sum(int(x[iS]) for x in data), sum(int(x[iI]) for x in data)
            ops = Counter()
            for x in data:
                tok = x[iSrc].split()
                op = (tok[1] if tok and tok[0].startswith("@") else tok[0]) if tok else "?"
                ops[op.split(".")[0]] += int(x[iI])
            md.append("Instruction mix (executed warp instructions): " + ", ".join("%s %.1f %%" % (k, 100 * v / ti) for k, v in ops.most_common(8)) + ".\n")
            bars = [i for i, x in enumerate(data) if "BAR.SYNC" in x[iSrc]]
            if bars:
                md.append("Phases between block-wide barriers (share of the warp samples / of the executed instructions / floating-point share of the phase):\n")
                md.append("| SASS range | samples | instructions | fp share |\n|---|---|---|---|")
                prev = 0
                for b in bars + [len(data) - 1]:
                    seg = data[prev:b + 1]
                    sm = sum(int(x[iS]) for x in seg); e = sum(int(x[iI]) for x in seg)
                    fp = sum(int(x[iI]) for x in seg if any(t in x[iSrc] for t in ("FFMA", "FMUL", "FADD")))
                    md.append("| %d..%d | %.1f %% | %.1f %% | %.0f %% |" % (prev, b, 100 * sm / max(tot, 1), 100 * e / max(ti, 1), 100 * fp / max(e, 1)))
                    prev = b + 1
                md.append("")
open(out_md, "w").write("\n".join(md))
print("wrote", out_md)
