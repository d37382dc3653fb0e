#!/usr/bin/env python3
"""Per-source-line view of an ncu capture: joins the SASS page of the report (stall samples and executed
instructions per instruction address) with nvdisasm's line info of the same cubin.
    python scripts/ncu_lines.py <rep> <lib.so> <kernel-substr> [top]
The .so must be the build that was profiled (compiled with -lineinfo)."""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, lib, sub = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
insts = []
for r in rows[hdr_i + 1:]:
    try:
        insts.append((int(r[0], 16), r[1].strip(), int(r[si]), int(r[ii])))
    except (ValueError, IndexError):
        pass
base = insts[0][0]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = max((os.path.join(tmp, f) for f in os.listdir(tmp)), key=os.path.getsize)
dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
line_of, cur, infn = {}, None, False
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        infn = sub in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)), "inlined" in m.group(3))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
agg = collections.defaultdict(lambda: [0, 0])
for addr, sass, smp, ex in insts:
    key = line_of.get(addr - base, ("?", 0, False))[:2]
    agg[key][0] += smp
    agg[key][1] += ex
tot_s, tot_i = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
print("kernel %s: %d samples, %d warp instructions" % (sub, tot_s, tot_i))
src_cache = {}
def src(f, n):
    if f not in src_cache:
        path = os.path.join(os.path.dirname(os.path.abspath(lib)), "..", "csrc", f)
        try:
            src_cache[f] = open(path).read().splitlines()
        except OSError:
            src_cache[f] = []
    L = src_cache[f]
    return L[n - 1].strip()[:90] if 0 < n <= len(L) else ""
for (f, n), (smp, ex) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100.0 * smp / max(1, tot_s), 100.0 * ex / max(1, tot_i), f, n, src(f, n)))
