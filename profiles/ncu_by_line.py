"""Attribute an ncu --import-source capture to CUDA source lines: joins the per-SASS-instruction rows of the report
(instructions executed, stall samples) with nvdisasm's line table of the same cubin, by instruction order.
usage: python profiles/ncu_by_line.py rep.ncu-rep libxptwarp.so 'k_stripILi4ELb0' [file_filter]
prints: per (file, line) instructions executed / samples, and totals per line range given as extra args a-b:name"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, so, fn_pat = sys.argv[1], sys.argv[2], sys.argv[3]
ranges = []
for a in sys.argv[4:]:
    m = re.match(r"(\d+)-(\d+):(\S+)", a)
    if m:
        ranges.append((int(m.group(1)), int(m.group(2)), m.group(3)))

MINLINE = int(os.environ.get('MINLINE', '130'))      # skip the small helpers at the top of the file
PFX = os.environ.get('FILEPFX', 'xpt_strip')         # the kernel's own source file (prefix)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines = []          # per instruction: (file, line, inline chain)
cur, loc, on, chain, fresh_chain = None, None, False, [], False
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m:
        on = fn_pat in m.group(1)
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh_chain:
            chain, fresh_chain = [], False
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        # innermost frame inside the kernel's own file (intrinsics headers are skipped), tagged with the outermost line
        own = [c for c in chain if c[0].startswith(PFX) and c[1] >= MINLINE] or [c for c in chain if c[0].startswith("xpt_")]
        loc = (own[0][0], own[0][1], chain[-1][1]) if own else (("?", 0, 0) if not chain else (chain[0][0], chain[0][1], chain[-1][1]))
        lines.append(loc)
        fresh_chain = True          # the next annotation starts a new chain; no annotation = same location
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr)]
if len(body) != len(lines):
    print(f"warning: {len(body)} profiled instructions vs {len(lines)} disassembled", file=sys.stderr)
per = collections.defaultdict(lambda: [0, 0, collections.Counter()])
stalls = [h for h in hdr if h.startswith("stall_")]
tot_i = tot_s = 0
for r, loc in zip(body, lines):
    ins = int(r[ix["Instructions Executed"]] or 0)
    smp = int(r[ix["# Samples"]] or 0)
    key = (loc[0], loc[1]) if loc else ("?", 0)
    per[key][0] += ins
    per[key][1] += smp
    for s in stalls:
        v = int(r[ix[s]] or 0)
        if v:
            per[key][2][s] += v
    tot_i += ins
    tot_s += smp
print(f"total warp-inst {tot_i}, samples {tot_s}")
if ranges:
    for lo, hi, name in ranges:
        i = sum(v[0] for k, v in per.items() if k[0].startswith(PFX) and lo <= k[1] <= hi)
        s = sum(v[1] for k, v in per.items() if k[0].startswith(PFX) and lo <= k[1] <= hi)
        st = collections.Counter()
        for k, v in per.items():
            if k[0].startswith(PFX) and lo <= k[1] <= hi:
                st.update(v[2])
        top = ", ".join(f"{n[6:]} {100 * c / max(1, sum(st.values())):.0f}%" for n, c in st.most_common(4))
        print(f"{name:10s} lines {lo}-{hi}: inst {100 * i / tot_i:5.1f}%  samples {100 * s / tot_s:5.1f}%  | {top}")
    other_i = sum(v[0] for k, v in per.items() if not k[0].startswith(PFX))
    other_s = sum(v[1] for k, v in per.items() if not k[0].startswith(PFX))
    print(f"other files: inst {100 * other_i / tot_i:5.1f}%  samples {100 * other_s / tot_s:5.1f}%")
print("--- top lines by instructions")
for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:45]:
    top = ", ".join(f"{n[6:]} {c}" for n, c in v[2].most_common(3))
    print(f"{k[0]}:{k[1]:4d}  inst {100 * v[0] / tot_i:5.2f}%  samples {100 * v[1] / tot_s:5.2f}%  {top}")
if os.environ.get("BYLINE"):
    lo, hi = [int(x) for x in os.environ["BYLINE"].split("-")]
    print(f"--- lines {lo}-{hi} in order")
    for k, v in sorted(per.items()):
        if k[0].startswith(PFX) and lo <= k[1] <= hi:
            top = ", ".join(f"{n[6:]} {c}" for n, c in v[2].most_common(3))
            print(f"{k[0]}:{k[1]:4d}  inst {100 * v[0] / tot_i:5.2f}%  samples {100 * v[1] / tot_s:5.2f}%  {top}")
