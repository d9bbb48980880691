"""Per-role view of an ncu --import-source capture of k_strip: every SASS instruction is attributed to the warp role
whose inlined body it belongs to (the call line inside k_strip), with instructions executed, stall mix and opcode
mix per role, and the hottest source lines of one role.
usage: python profiles/ncu_roles.py rep.ncu-rep libxptwarp.so k_stripILi4ELb0 [role-for-line-detail]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, so, fn_pat = sys.argv[1], sys.argv[2], sys.argv[3]
detail = sys.argv[4] if len(sys.argv) > 4 else None
src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "xpt-mde-2021_b200", "csrc", "xpt_strip.cuh")
src_lines = open(src).read().splitlines()
role_of_line = {}
for i, l in enumerate(src_lines, 1):
    m = re.search(r"strip_role_(\w)<", l)
    if m and ("if (wid" in l or "else" in l or l.strip().startswith("strip_role")):
        role_of_line[i] = m.group(1).upper()
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
locs, on, chain, fresh = [], False, [], False
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m:
        on = fn_pat in m.group(1)
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m:
        role = role_of_line.get(chain[-1][1], "-") if chain else "-"
        own = [c for c in chain if c[0] == "xpt_strip.cuh"]
        locs.append((role, own[0][1] if own else 0, m.group(1)))
        fresh = True
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr)]
assert len(body) == len(locs), (len(body), len(locs))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
R = collections.defaultdict(lambda: dict(inst=0, smp=0, st=collections.Counter(), ops=collections.Counter(), lines=collections.Counter(),
                                     lsmp=collections.Counter(), lst=collections.defaultdict(collections.Counter)))
for r, (role, line, txt) in zip(body, locs):
    n = int(r[ix["Instructions Executed"]] or 0)
    d = R[role]
    d["inst"] += n
    d["smp"] += int(r[ix["# Samples"]] or 0)
    for s in stalls:
        v = int(r[ix[s]] or 0)
        if v:
            d["st"][s[6:]] += v
    op = txt.split()
    o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
    d["ops"][o.split(".")[0]] += n
    d["lines"][line] += n
    d["lsmp"][line] += int(r[ix["# Samples"]] or 0)
    for s_ in stalls:
        v = int(r[ix[s_]] or 0)
        if v:
            d["lst"][line][s_[6:]] += v
tot = sum(d["inst"] for d in R.values())
tots = sum(d["smp"] for d in R.values())
print(f"total warp-inst {tot}, samples {tots}")
for role in "LOYSG-":
    if role not in R:
        continue
    d = R[role]
    st = ", ".join(f"{k} {100 * v / max(1, sum(d['st'].values())):.0f}%" for k, v in d["st"].most_common(5))
    ops = ", ".join(f"{k} {100 * v / max(1, d['inst']):.0f}%" for k, v in d["ops"].most_common(14))
    print(f"{role}: inst {100 * d['inst'] / tot:5.1f}% ({d['inst']})  samples {100 * d['smp'] / max(1, tots):5.1f}%\n    stalls: {st}\n    ops: {ops}")
if detail:
    d = R[detail]
    print(f"--- hottest lines of role {detail}")
    print("line  inst%  samples%  top stalls")
    for line, n in d["lsmp"].most_common(32):
        txt = src_lines[line - 1].strip()[:90] if 0 < line <= len(src_lines) else ""
        st = ", ".join(f"{k} {v}" for k, v in d["lst"][line].most_common(3))
        print(f"{line:4d} {100 * d['lines'][line] / d['inst']:5.1f}% {100 * n / max(1, d['smp']):5.1f}%  [{st}]  {txt}")
