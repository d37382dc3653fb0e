#!/usr/bin/env python3
"""SASS helper: python scripts/sass.py <lib.so> <kernel-substr> [grep-regex]  -> compact listing / opcode histogram."""
import re, subprocess, sys, collections
lib, sub = sys.argv[1], sys.argv[2]
pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, rows = None, []
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and sub in cur:
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m:
            rows.append((int(m.group(1), 16), m.group(2).strip()))
if pat is None:
    hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", r[1]).split()[0].split(".")[0] for r in rows)
    print(len(rows), "instructions"); print(", ".join("%s %d" % kv for kv in hist.most_common(40)))
else:
    for a, ins in rows:
        if pat.search(ins):
            print("%05x  %s" % (a, ins[:110]))
