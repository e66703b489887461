"""Developer tool: FP64 arithmetic thread-instructions (DADD + DMUL + DFMA, predicated-on) of every render_bvh /
render_exact launch in an .ncu-rep, summed -> one entry of profiles/fp64_ops.json.

usage: python tools/ncu_fp64_ops.py <key, e.g. c2:exact> report.ncu-rep [profiles/fp64_ops.json]

The report must come from ONE render of the workload on one GPU (tools/sweep.py ... 1) captured with
`ncu --set full` (the per-opcode instanced metric sass__thread_inst_executed_true_per_opcode is part of that set).
bench.py divides the sum by the measured kernel time: roofline.achieved."""
import csv
import json
import os
import re
import subprocess
import sys

key, rep = sys.argv[1], sys.argv[2]
dst = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "profiles", "fp64_ops.json")
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-metric-instances", "details", "--metrics",
                      "sass__thread_inst_executed_true_per_opcode,gpu__time_duration.sum"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[0]
ik, io = hdr.index("Kernel Name"), hdr.index("sass__thread_inst_executed_true_per_opcode")
total = {}
launches = []
for vals in rows[2:]:
    name = vals[ik]
    if "render_bvh" not in name and "render_exact" not in name:
        continue
    per = {m.group(1): int(m.group(2)) for m in re.finditer(r"([A-Z0-9_.]+): (\d+)", vals[io])}
    launches.append({"kernel": name.split("(")[0], "thread_inst": sum(per.values())})
    for k, v in per.items():
        total[k] = total.get(k, 0) + v
arith = {k: total.get(k, 0) for k in ("DADD", "DMUL", "DFMA")}
entry = {"fp64_arith_thread_inst": sum(arith.values()), "per_opcode": {**arith, "DSETP": total.get("DSETP", 0),
                                                                      "MUFU": total.get("MUFU", 0)},
         "all_thread_inst": sum(total.values()), "launches": launches,
         "source": f"profiles/{os.path.basename(rep).replace('.ncu-rep', '')} (ncu --set full, "
                   "sass__thread_inst_executed_true_per_opcode; captured under ncu, not a timing)"}
data = json.load(open(dst)) if os.path.exists(dst) else {}
data[key] = entry
json.dump(data, open(dst, "w"), indent=1)
print(json.dumps({key: entry}, indent=1))
