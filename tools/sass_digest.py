#!/usr/bin/env python
"""SASS digest of libmpb200.so: per kernel, the opcode histogram of the mnemonics that prove (or disprove) a
Blackwell-native kernel -- UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG/UBLKCP (TMA), SYNCS
(mbarrier), LDS/STS, LDG/STG, HMMA (legacy mma.sync: expected in k_stem3x3_relu_maxpool_f16 only -- the K = 27 RGB stem fused
with its max-pool, csrc/stem_pool.cu -- and nowhere else) -- plus registers / code size from cuobjdump.

    python tools/sass_digest.py [out.txt]          (no GPU needed: cuobjdump reads the cubin inside the .so)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "megaportrait-hack_b200", "libmpb200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA",
        "LDS", "STS", "LDG", "STG", "LD", "ST", "ATOM", "RED", "FFMA", "BAR", "SHFL", "MATCH", "REDUX", "ELECT"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+).*?SHARED:(\d+)", res):
        regs[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
            kernels[cur]["__total__"] += 1
            if m.group(1) in ("UTMALDG", "UTMASTG", "LDTM", "UTCHMMA", "UTCQMMA"):
                kernels[cur][m.group(1) + m.group(2)] += 1
    lines = ["# SASS digest of megaportrait-hack_b200/libmpb200.so (cuobjdump -sass, sm_100a); demangled kernel name, "
             "instructions, registers, then the counts of the mnemonics that matter", ""]
    for name, c in kernels.items():
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(anonymous namespace\)::", "", dem)
        dem = re.sub(r"\(.*", "", dem)
        r = regs.get(name, ("?", "?"))
        lines.append(f"{dem}   [{c['__total__']} instr, {r[0]} regs, {r[1]} B static smem]")
        parts = [f"{k}={c[k]}" for k in KEYS if c.get(k)]
        detail = [f"{k}={v}" for k, v in sorted(c.items()) if "." in k]
        lines.append("    " + " ".join(parts))
        if detail:
            lines.append("    " + " ".join(detail))
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    else:
        sys.stdout.write(text)


if __name__ == "__main__":
    main()
