"""Opcode histogram and code size per kernel of libisr.so (cuobjdump -sass) -> profiles/rNN_sass_summary.json.
Runs without a GPU.    python scripts/sass_summary.py [out.json]"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "imagesequenceregistrationfor6dposeestimationlabeling_b200", "csrc", "libisr.so")
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_summary.json")
text = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kernels, cur = {}, None
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
# the opcodes that say what kind of kernel it is
KEY = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FMNMX3", "FMNMX", "DFMA", "DADD", "DMUL", "MUFU", "UBLKCP", "SYNCS", "LDS", "STS",
       "LDG", "STG", "LDL", "STL", "SHFL", "ATOMG", "ATOMS", "RED", "BAR", "WARPSYNC", "VOTE", "CALL"]
res = []
for name, c in sorted(kernels.items(), key=lambda kv: -sum(kv[1].values())):
    n = sum(c.values())
    res.append({"kernel": demangle(name)[:160], "instructions": n, "code_bytes": 16 * n,
                "key_opcodes": {k: c[k] for k in KEY if c.get(k)},
                "top_opcodes": dict(c.most_common(12))})
json.dump({"source": "cuobjdump -sass csrc/libisr.so (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo)",
           "note": "FFMA2 / FADD2 / FMUL2 = Blackwell packed FP32; FMNMX3 = three-input min; UBLKCP = cp.async.bulk "
                   "(TMA engine, 1-D bulk copy) with SYNCS = mbarrier operations; no UTCxMMA / UTMALDG by design "
                   "(the contraction stays on the FP32 CUDA cores, 1-D planes need no tensor map)",
           "kernels": res}, open(out_path, "w"), indent=1)
for r in res[:14]:
    print(f"{r['instructions']:6d} instr  {r['kernel'][:90]}  {r['key_opcodes']}")
