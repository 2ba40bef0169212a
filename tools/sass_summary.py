#!/usr/bin/env python3
"""SASS evidence for profiles/: mnemonic histogram of every kernel in csrc/libbfsm_b200.so plus the full
listings of the two gain kernels of the default 64^3 pipeline (encodings stripped).

    python tools/sass_summary.py        ->  profiles/r02_sass_summary.txt, r02_sass_k_plane_gain_ws.txt,
                                            r02_sass_k_pencil_gain_async_tma.txt, r02_sass_k_plane_gain_r32_tmem_32.txt

What to look for: UTMALDG (cp.async.bulk.tensor: the TMA-filled x-stage ring) + SYNCS (mbarrier),
LDTM / STTM / UTCATOMSWS (tcgen05.ld / .st / .alloc: the fhat line of the radix-32 plane kernel in tensor memory),
USETMAXREG (register hand-over between the warpgroups of the pipelined plane kernel), UCGABAR_* (cluster
barriers of the 32^3 DSMEM kernel), LDGSTS (cp.async), SHFL (lane butterflies of the register-resident
x stage), no HMMA/DMMA (fp64 butterflies are not a contraction; DMMA shares the FP64 pipe on sm_100a,
tools/microbench.cu), ATOMG only in k_gain_fused (hand-over counters, never data).
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "boltzmann-fourier-spectral-method_b200", "csrc", "libbfsm_b200.so")
OUT = os.path.join(ROOT, "profiles")
KEEP = {"k_plane_gain_wsILi64E": "r02_sass_k_plane_gain_ws.txt",
        "k_pencil_gain_asyncILi64ELi4ELi3ELi2ELb0ELb1": "r02_sass_k_pencil_gain_async_tma.txt",
        "k_plane_gain_r32ILi32ELi11ELi1ELb1": "r02_sass_k_plane_gain_r32_tmem_32.txt"}
INTERESTING = ["UTMALDG", "UBLKCP", "SYNCS", "LDTM", "STTM", "UTCATOMSWS", "LDGSTS", "USETMAXREG", "UCGABAR_ARV", "UCGABAR_WAIT", "BAR", "DFMA",
               "DADD", "DMUL", "SHFL", "LDS", "STS", "LDG", "STG", "ST", "LD", "ATOMG", "MEMBAR", "HMMA", "DMMA",
               "LDL", "STL"]

txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
tot, per = collections.Counter(), {}
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    ops = collections.Counter()
    for line in f.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(2).split(".")[0]] += 1
    per[name] = ops
    tot.update(ops)
    for key, fn in KEEP.items():
        if key in name:
            body = "\n".join(l for l in f.split("\n") if re.match(r"\s+/\*[0-9a-f]{4}\*/", l))
            body = re.sub(r"\s*/\* 0x[0-9a-f]+ \*/", "", body)
            with open(os.path.join(OUT, fn), "w") as fh:
                fh.write("// cuobjdump -sass csrc/libbfsm_b200.so, function " + name + "\n" + body + "\n")
lines = ["mnemonic histogram of csrc/libbfsm_b200.so (sm_100a cubin, %d kernels); static SASS instruction counts" % len(funcs),
         "", "whole library: " + ", ".join("%s %d" % (k, tot[k]) for k in INTERESTING if tot[k]), ""]
for name, ops in sorted(per.items()):
    short = re.sub(r"^_ZN4bfsm\d+", "", name)
    lines.append(short[:90] + ": " + ", ".join("%s %d" % (k, ops[k]) for k in INTERESTING if ops[k]))
with open(os.path.join(OUT, "r02_sass_summary.txt"), "w") as fh:
    fh.write("\n".join(lines) + "\n")
print("\n".join(lines[:3]))
