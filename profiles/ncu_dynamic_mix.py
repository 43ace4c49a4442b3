#!/usr/bin/env python3
"""Dynamic SASS opcode mix of one kernel from `ncu --page source --csv`: warp-instructions executed per opcode."""
import collections, csv, re, subprocess, sys
rep, kregex = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kregex}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
start = hi[0]; end = hi[1] - 1 if len(hi) > 1 else len(rows)
hdr = rows[start]; data = [r for r in rows[start + 1:end] if len(r) == len(hdr)]
ie = hdr.index("Instructions Executed")
ops = collections.Counter(); tot = 0
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[1])
    n = int(r[ie] or 0); tot += n
    if m: ops[m.group(2)] += n
print(f"# {rep} kernel~{kregex}: {tot} warp-instructions executed")
for k, v in ops.most_common(30): print(f"{k:16s} {v:14d} {100*v/tot:5.1f}%")
