"""Executed-instruction mix per object-model pair (or per distance) from the source page of an `ncu --set full
--import-source on` report: thread-level instructions of every SASS opcode, divided by the units the launch processed.
Usage: python tools/ncu_opmix.py report.ncu-rep units_per_launch[,units...] > table.md
(one `units` per profiled launch, in launch order; 0 skips a launch)"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
units = [float(u) for u in sys.argv[2].split(",")]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
secs, cur = [], None
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
# the source page lists every launch twice: keep one section of each pair
secs = secs[::2] if len(secs) % 2 == 0 and all(secs[i]["name"] == secs[i + 1]["name"] for i in range(0, len(secs), 2)) else secs
for s, u in zip(secs, units):
    if u <= 0:
        continue
    hdr = s["rows"][0]
    ia, it = hdr.index("Source"), hdr.index("Thread Instructions Executed")
    cls = collections.Counter()
    for r in s["rows"][1:]:
        if len(r) > it and r[it].isdigit():
            toks = r[ia].split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            cls[op.split(".")[0]] += int(r[it])
    tot = sum(cls.values())
    print("## `%s`\n" % s["name"].replace("<unnamed>::", "").replace("(int)", "").replace("(bool)", "")[:150])
    print("%.4g units per launch, %.2f thread instructions per unit, %d SASS lines\n" % (u, tot / u, len(s["rows"]) - 1))
    print("| opcode | thread instructions per unit |\n|---|---:|")
    for k, v in cls.most_common():
        if v / u >= 0.05:
            print("| %s | %.2f |" % (k, v / u))
    print()
