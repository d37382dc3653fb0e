#!/usr/bin/env python3
"""Innermost-loop report for one kernel in libkgpu.so (works without a GPU):
    python scripts/sass_loops.py <lib.so> <kernel-substr> [--dump FILE]
For every backward branch (a loop) that contains no other backward branch: extent, instruction count split
into IMAD-pipe / ALU-pipe / other, and the issue-slot estimate of DESIGN.md 5 (IMAD 1, ALU 2, other 1).
CREDUX/REDUX inside the loop marks a (K,F) bucket loop of the sparse scorer."""
import collections
import re
import subprocess
import sys

FMA = {"IMAD", "FFMA", "FMUL", "FADD", "IMAD.MOV", "IMAD.SHL", "IMAD.IADD", "IMAD.WIDE", "IMAD.U32", "IMAD.X"}
ALU = {"IADD3", "IADD", "VIMNMX3", "VIMNMX", "VIADDMNMX", "LOP3", "SHF", "SEL", "ISETP", "LEA", "PRMT", "IABS", "POPC", "FLO",
       "VIADD", "MOV", "ISETP", "PLOP3", "SGXT", "BMSK", "IMNMX", "I2FP", "LOP", "R2P", "P2R"}


def functions(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, out = None, collections.OrderedDict()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if cur and m:
            out[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return out


def opcode(ins):
    ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
    return ins.split()[0]


def klass(op):
    base = op.split(".")[0]
    if base in ("IMAD", "FFMA", "FMUL", "FADD"):
        return "fma"
    if base in ALU or base.startswith("U"):
        return "alu" if not base.startswith("U") else "uni"
    return "oth"


def main():
    lib, sub = sys.argv[1], sys.argv[2]
    fns = functions(lib)
    names = [n for n in fns if sub in n]
    if not names:
        sys.exit("no kernel matches %r" % sub)
    rows = fns[names[0]]
    print("#", names[0], len(rows), "instructions")
    if "--dump" in sys.argv:
        with open(sys.argv[sys.argv.index("--dump") + 1], "w") as f:
            for a, ins in rows:
                f.write("%05x %s\n" % (a, ins))
    addr_index = {a: i for i, (a, _) in enumerate(rows)}
    loops = []
    for i, (a, ins) in enumerate(rows):
        m = re.search(r"BRA(?:\.\w+)*\s+(?:U?P\d+,\s*)?0x([0-9a-f]+)", ins)
        if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in addr_index:
            loops.append((addr_index[int(m.group(1), 16)], i))
    inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
    for lo, hi in inner:
        body = rows[lo:hi + 1]
        cnt = collections.Counter(klass(opcode(ins)) for _, ins in body)
        ops = collections.Counter(opcode(ins).split(".")[0] for _, ins in body)
        redux = any("REDUX" in ins for _, ins in body)
        slots = cnt["fma"] + 2 * cnt["alu"] + cnt["oth"] + cnt["uni"]
        print("%05x-%05x n=%3d fma=%3d alu=%3d uni=%2d oth=%2d slots~%3d %s | %s" % (
            rows[lo][0], rows[hi][0], len(body), cnt["fma"], cnt["alu"], cnt["uni"], cnt["oth"], slots,
            "REDUX" if redux else "     ", " ".join("%s:%d" % kv for kv in ops.most_common(12))))


if __name__ == "__main__":
    main()
