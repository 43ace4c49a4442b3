#!/usr/bin/env python3
"""Per-kernel SASS evidence from the built library (no GPU needed): bulk-copy engine (UBLKCP = cp.async.bulk, the 1-D TMA path),
mbarrier transactions (SYNCS), fp64 arithmetic, peer / exchange stores.   python profiles/sass_evidence.py > profiles/rNN_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "globalsfmpy_b200", "csrc", "libgsfm_ra.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}   (arch: {re.search(r'arch = (sm_[0-9a-z]+)', sass).group(1)})")
print("# kernel | instructions | UBLKCP (cp.async.bulk) | SYNCS (mbarrier) | DFMA+DMUL+DADD | STG | LDG | SHFL | BAR | registers are in -Xptxas -v")
for block in sass.split("Function : ")[1:]:
    name = block.split("\n", 1)[0]
    short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    short = re.sub(r"\(anonymous namespace\)::", "", short).split("(")[0].replace("void ", "")
    ops = collections.Counter()
    for line in block.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            ops[m.group(2)] += 1
    n = sum(ops.values())
    if not any(k in short for k in ("k_edges", "k_pcg_persistent", "k_spmv", "k_stream_probe", "k_node_finalize", "k_dense_cholesky")):
        continue
    print(f"{short:60s} {n:6d} | UBLKCP {ops['UBLKCP']:3d} | SYNCS {ops['SYNCS']:3d} | fp64 {ops['DFMA'] + ops['DMUL'] + ops['DADD']:5d} | "
          f"STG {ops['STG']:3d} | LDG {ops['LDG']:3d} | SHFL {ops['SHFL']:3d} | BAR {ops['BAR']:2d}")
