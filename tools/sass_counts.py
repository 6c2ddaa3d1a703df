#!/usr/bin/env python
"""SASS opcode census of libvecgpu.so per kernel (cuobjdump -sass): the mnemonics that prove which hardware paths a kernel
uses — UTCHMMA / UTCIMMA (tcgen05.mma kind::tf32 / kind::i8), LDTM (tcgen05.ld), UTMALDG (TMA tensor loads, .MULTICAST),
UBLKCP (cp.async.bulk), SYNCS (mbarrier), POPC, IDP (dp4a), FFMA, LDGSTS (cp.async), REDUX, ATOMS ...
    python tools/sass_counts.py > profiles/r2_sass_opcode_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "sqlite-vec-hnsw_b200", "libvecgpu.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
KEY = ["UTCHMMA", "UTCIMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "POPC", "IDP", "FFMA", "FMUL", "FADD", "DFMA",
       "LDGSTS", "REDUX", "ATOMS", "ATOMG", "ATOM", "RED", "LDG", "LDS", "STS", "STG", "SHFL", "BAR", "MEMBAR", "UCGABAR", "HMMA", "IMMA"]
cur, counts = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        op, mods = m.group(1), m.group(2)
        counts[cur][op] += 1
        if op == "UTMALDG" and "MULTICAST" in mods:
            counts[cur]["UTMALDG.MULTICAST"] += 1
        counts[cur]["_total"] += 1
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)} : instructions per kernel (sm_100a), selected opcodes")
for fn, c in counts.items():
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(.*", "", name)
    sel = [f"{k}={c[k]}" for k in KEY + ["UTMALDG.MULTICAST"] if c.get(k)]
    print(f"{name}\n    total={c['_total']}  " + "  ".join(sel))
