"""Developer tool: DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum, bytes per launch) of the kernels in an
.ncu-rep -> JSON {kernel base name: bytes}.  usage: python tools/ncu_traffic.py report.ncu-rep [more.ncu-rep ...]
bench.py reads profiles/ncu_traffic.json for roofline.traffic."""
import csv
import json
import subprocess
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")].split("<")[0].split("(")[0].replace("tor::", "").replace("void ", "").strip()
        total = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(k)
            total += float(vals[i].replace(",", "")) * UNIT[units[i]]
        out[name] = int(total)
        out[name + "_source"] = rep.split("/")[-1]
print(json.dumps(out, indent=1))
