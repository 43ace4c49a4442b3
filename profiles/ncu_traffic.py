#!/usr/bin/env python3
"""DRAM traffic per CG step of the persistent PCG kernel from an `ncu --set full` capture of profiles/kernel_times.py (the captured
launch runs exactly `steps` CG steps, rtol 0) -> profiles/ncu_traffic.json, which bench.py reports as roofline.traffic.

  python profiles/ncu_traffic.py <workload>:<n_gpus> <file.ncu-rep> <steps> [more triples ...]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
data = json.load(open(out_path)) if os.path.exists(out_path) else {}
args = sys.argv[1:]
for key, rep, steps in zip(args[0::3], args[1::3], args[2::3]):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    r = [x for x in rows[2:] if "k_pcg_persistent" in x[hdr.index("Kernel Name")]][0]

    def val(name):
        v, u = float(r[hdr.index(name)].replace(",", "")), units[hdr.index(name)]
        return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    data[key] = {"dram_bytes_per_cg_step": (rd + wr) / int(steps), "dram_read_bytes": rd, "dram_write_bytes": wr, "cg_steps_in_launch": int(steps),
                 "lts_hit_rate_pct": float(r[hdr.index("lts__t_sector_hit_rate.pct")]),
                 "source": f"profiles: ncu --set full of {os.path.basename(rep)} ({steps} CG steps in the captured launch)"}
    print(key, data[key])
json.dump(data, open(out_path, "w"), indent=1)
