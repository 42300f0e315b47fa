#!/usr/bin/env python3
"""SASS instruction counts of the field / curve primitives ("no SASS listing or instruction count is
committed").  Disassembles tools/microbench (cuobjdump -sass), counts per single-operation kernel (k_one_*) the
IMAD.WIDE (the half-rate 32x32->64 multiply-add), single-half multiplies (IMAD / IMAD.HI), move-like IMAD forms
(IMAD.MOV, IMAD.X, IMAD.SHL, IMAD.IADD), IADD3-family and everything else, and subtracts the load/store skeleton (k_one_ldst).

    make -C tools && python tools/sass_counts.py > profiles/sass_counts_r2.json
"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    exe = os.path.join(ROOT, "tools", "microbench")
    sass = subprocess.run(["cuobjdump", "-sass", exe], capture_output=True, text=True, check=True).stdout
    kernels, name = {}, None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kernels[name] = {"imad_wide": 0, "imad_mul": 0, "imad_movlike": 0, "iadd3": 0, "other": 0, "total": 0}
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if not m or name is None:
            continue
        op = m.group(1)
        c = kernels[name]
        c["total"] += 1
        if op.startswith("IMAD.WIDE"):
            c["imad_wide"] += 1
        elif op in ("IMAD", "IMAD.HI.U32", "IMAD.HI", "IMAD.U32"):
            c["imad_mul"] += 1          # a real 32x32 multiply (lo or hi half): fmaheavy pipe
        elif op.startswith("IMAD"):
            c["imad_movlike"] += 1      # IMAD.MOV / IMAD.X / IMAD.SHL / IMAD.IADD: moves and adds ptxas puts on an FMA pipe
        elif op.startswith("IADD3") or op.startswith("IADD"):
            c["iadd3"] += 1
        else:
            c["other"] += 1
    out = {}
    for nm, c in kernels.items():
        m = re.match(r"void k_one_(\w+)<b2p::Field<b2p::(\w+?)Params>\s*>", nm)
        if not m:
            continue
        out.setdefault(m.group(2), {})[m.group(1)] = c
    res = {"source": "cuobjdump -sass tools/microbench (nvcc 12.9, sm_100a, -O3)", "fields": {}}
    for fld, ops in out.items():
        base = ops.get("ldst", {"imad_wide": 0, "imad_mul": 0, "imad_movlike": 0, "iadd3": 0, "other": 0, "total": 0})
        res["fields"][fld] = {op: {"raw": c, "minus_ldst_skeleton": {k: c[k] - base[k] for k in c}}
                              for op, c in ops.items()}
    json.dump(res, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
