#!/usr/bin/env python
"""Turn the ncu artefacts of a round (gpurun_out/) into the tracked summaries under profiles/.

    python tools/make_profiles.py <tag> <pages> <compress.ncu-rep> <decompress.ncu-rep> <launches.csv> <bench.json>
"""
import csv
import json
import shutil
import subprocess
import sys

tag, pages, rep_c, rep_d, launches, bench = sys.argv[1:7]
pages = int(pages)


def dram(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    d, un = dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))
    mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return int(sum(float(d[k]) * mul[un[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum")))


for name, rep in (("compress", rep_c), ("decompress", rep_d)):
    with open(f"profiles/{tag}_{name}_kernel.txt", "w") as f:
        f.write(subprocess.run([sys.executable, "tools/ncu_summary.py", rep], capture_output=True, text=True).stdout)
        f.write(subprocess.run([sys.executable, "tools/ncu_hot.py", rep, str(pages), "1e9", "1.2"], capture_output=True,
                               text=True).stdout)
json.dump({"how": "ncu --set full --clock-control none, one launch each of `python bench.py --pages %d` (default workload); "
                  "dram__bytes_read.sum + dram__bytes_write.sum" % pages,
           "pages": pages, "compress": dram(rep_c), "decompress": dram(rep_d)}, open("profiles/traffic.json", "w"), indent=1)
rows = list(csv.reader(open(launches)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ik, ib, ig, iv = h.index("Kernel Name"), h.index("Block Size"), h.index("Grid Size"), h.index("Metric Value")
out, tot = [["id", "kernel", "block", "grid", "gpu__time_duration.sum [ns]"]], {}
for r in rows[hdr + 1:]:
    if len(r) > iv and "csb::" in r[ik]:
        out.append([r[0], r[ik], r[ib], r[ig], r[iv]])
        k = r[ik].split("(")[0]
        tot[k] = tot.get(k, 0) + float(r[iv])
csv.writer(open(f"profiles/{tag}_launches.csv", "w")).writerows(out)
with open(f"profiles/{tag}_launches_summary.txt", "w") as f:
    f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none -c 80 python bench.py --pages {pages} --steps 2 --warmup 3 --no-cpu --no-alt\n"
            "(cold-cache, serialised launch times: compare SHARES with bench.py's live CUDA-event times; the small launches are "
            "the e2e leg's 8192-page chunks)\n")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        f.write(f"{k:70s} {v / 1e6:10.3f} ms  {100 * v / sum(tot.values()):5.1f}%\n")
shutil.copy(bench, f"profiles/{tag}_bench_1Mi_pages.json")
print(open(f"profiles/{tag}_launches_summary.txt").read())
print(open("profiles/traffic.json").read())
