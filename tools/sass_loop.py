#!/usr/bin/env python3
"""usage: tools/sass_loop.py <kernel-substring> [min_table_loads]
Finds the hot loop of a kernel in libisscabac.so (the innermost backward branch whose body holds at
least `min_table_loads` shared-memory table loads, default 16 = one 16-op block) and prints its
instruction mix by opcode and by issue pipe (alu / fma / xu / lsu / branch ...)."""
import collections
import re
import subprocess
import sys

ALU = {"IADD3", "LOP3", "SHF", "PRMT", "SEL", "ISETP", "VIMNMX", "VIADD", "LEA", "MOV", "IABS", "BMSK", "SGXT", "PLOP3", "IADD", "VIMNMX3", "FSEL", "P2R", "R2P"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD"}
XU = {"FLO", "POPC", "BREV", "MUFU", "I2F", "F2I"}
LSU = {"LDS", "STS", "LDG", "STG", "LD", "ST", "ATOMS", "ATOMG", "RED", "LDC", "LDCU"}


def main():
    name = sys.argv[1]
    need = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    out = subprocess.run(["bash", "tools/sass.sh", name], capture_output=True, text=True).stdout.splitlines()
    ins = []
    for l in out:
        m = re.match(r"^([0-9a-f]+)\s+(.*?)\s*;?\s*$", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\s+(0x[0-9a-f]+)", t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= a:
            continue
        body = [x for x in ins if tgt <= x[0] <= a]
        nl = sum(1 for x in body if re.search(r"\bLDS(\.\w+)*\b", x[1]) and (".128" in x[1] or ".64" in x[1]))
        if nl >= need and (best is None or len(body) < len(best)):
            best = body
    if best is None:
        print("no loop found")
        return
    ops = collections.Counter()
    pipes = collections.Counter()
    for _, t in best:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        op = t.split()[0].split(".")[0]
        ops[op] += 1
        p = "alu" if op in ALU else "fma" if op in FMA else "xu" if op in XU else "lsu" if op in LSU else "ctl"
        pipes[p] += 1
    n = len(best)
    print(f"{name}: loop {best[0][0]:#x}..{best[-1][0]:#x}, {n} instructions")
    print("  pipes:", dict(pipes))
    print("  ops:", dict(ops.most_common()))


if __name__ == "__main__":
    main()
