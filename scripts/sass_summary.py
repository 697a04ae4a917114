#!/usr/bin/env python
"""SASS opcode summary of the in-tree libjexrhs.so (cuobjdump -sass): per kernel family and for the default element kernels the
counts of the opcodes that show what the code is made of (DFMA/DMUL/DADD, LDS/STS, LDG/STG, REDG, LDGSTS, UBLKCP/UBLKPF bulk
copies, UTMALDG tensor-map loads, BAR).  Usage: python scripts/sass_summary.py [lib] > profiles/rNN_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "jexpresso_b200", "lib", "libjexrhs.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
OPS = ["DFMA", "DMUL", "DADD", "LDS", "STS", "LDG", "STG", "REDG", "LDGSTS", "UBLKCP", "UBLKPF", "UTMALDG", "BAR", "LDL", "STL", "MUFU"]
kern = None
arch = set()
counts = collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        counts[kern]["_total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                counts[kern][o] += 1
print(f"# SASS opcode summary of {os.path.relpath(lib, ROOT)}\n")
print(f"cubin architectures: {', '.join(sorted(arch))}; {len(counts)} kernels\n")
fam = collections.OrderedDict()
for k, c in counts.items():
    name = re.sub(r"<.*", "", k.replace("void ", "")).replace("jx::", "").replace("(anonymous namespace)::", "")
    f = fam.setdefault(name, [0, collections.Counter()])
    f[0] += 1
    f[1].update(c)
print("## per kernel family (sum over all instantiations)\n")
print("| kernel | instantiations | instr | " + " | ".join(OPS) + " |")
print("|---|---|---|" + "---|" * len(OPS))
for name, (n, c) in fam.items():
    print(f"| {name} | {n} | {c['_total']} | " + " | ".join(str(c[o]) for o in OPS) + " |")
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print(f"| **library** | {len(counts)} | {tot['_total']} | " + " | ".join(str(tot[o]) for o in OPS) + " |")
print("\n## default kernels of the bench configuration (3D theta TOTAL, jx_pow, nop 4)\n")
print("| kernel | instr | " + " | ".join(OPS) + " |")
print("|---|---|" + "---|" * len(OPS))
want = ["k_elem_team<5, jx::EulerTheta<3, false, true>, 2, 2, 2, false>", "k_elem_team<5, jx::EulerTheta<3, false, true>, 2, 2, 2, true>",
        "k_elem_team<5, jx::EulerTheta<3, false, true>, 2, 0, 2, false>", "k_visc_quad<5, jx::EulerTheta<3, false, true>, 2>",
        "k_visc_quad<5, jx::EulerTheta<3, false, true>, 0>", "k_visc_team<5, jx::EulerTheta<3, false, true>, 2>",
        "k_stage_direct<jx::EulerTheta<3, false, true>", "k_node_aux<jx::EulerTheta<3, false, true>",
        "k_bc_dirichlet<jx::EulerTheta<3, false, true>", "k_elem_node<3, 5, jx::EulerTheta<3, false, true>, true",
        "k_elem_node<3, 8, jx::EulerTheta<3, false, true>, false", "k_elem_node<2, 6, jx::EulerTheta<2, false, true>, true",
        "k_elem_tri<8, jx::EulerTheta<3, false, true>, 2>", "k_gather<5>"]
for w in want:
    for k, c in counts.items():
        kk = k.replace("(int)", "").replace("(bool)0", "false").replace("(bool)1", "true")
        if w in kk:
            print(f"| `{kk[:110]}` | {c['_total']} | " + " | ".join(str(c[o]) for o in OPS) + " |")
            break
