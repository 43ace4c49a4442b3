#!/usr/bin/env python3
"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv` (one kernel)."""
import csv
import subprocess
import sys

rep, kregex = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kregex}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if not hi:
    sys.exit("no source page")
start = hi[0]
end = hi[1] - 1 if len(hi) > 1 else len(rows)   # first launch only
hdr = rows[start]
data = [r for r in rows[start + 1:end] if len(r) == len(hdr)]
si = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in data)
agg = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in stall_cols}
print(f"# {rep} kernel~{kregex}: {len(data)} SASS instructions, {tot} samples")
print("# stall totals:", ", ".join(f"{k[6:]}={v} ({100*v/max(1,tot):.0f}%)" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
for r in sorted(data, key=lambda r: -int(r[si] or 0))[:top]:
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{int(r[si]):6d} {100*int(r[si])/max(1,tot):5.1f}%  {r[1].strip():60s} {st}")
