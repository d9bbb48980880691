"""Static per-role view of k_strip's SASS (no GPU): instructions, spill instructions and highest register per warp role.
usage: python profiles/sass_roles.py libxptwarp.so [k_stripILi4ELb0]"""
import collections, os, re, subprocess, sys, tempfile
so = sys.argv[1]; fn = sys.argv[2] if len(sys.argv) > 2 else "k_stripILi4ELb0"
src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "xpt-mde-2021_b200", "csrc", "xpt_strip.cuh")).read().splitlines()
role_of = {}
for i, l in enumerate(src, 1):
    m = re.search(r"strip_role_(\w)<", l)
    if m and ("if (wid" in l or "else" in l or l.strip().startswith("strip_role")):
        role_of[i] = m.group(1).upper()
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
on = False; chain = []; fresh = False
tot = collections.Counter(); spill = collections.Counter(); maxreg = collections.defaultdict(int); ops = collections.defaultdict(collections.Counter)
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m:
        on = fn in m.group(1); continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh: chain = []; fresh = False
        chain.append((os.path.basename(m.group(1)), int(m.group(2)))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        role = role_of.get(chain[-1][1], "-") if chain else "-"
        tot[role] += 1
        if re.search(r"\b(STL|LDL)", m.group(2)): spill[role] += 1
        for r in re.findall(r"\bR(\d+)\b", m.group(2)): maxreg[role] = max(maxreg[role], int(r))
        op = m.group(2).split(); o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
        ops[role][o.split(".")[0]] += 1
        fresh = True
for role in sorted(tot):
    print(f"{role}: {tot[role]:5d} instr, {spill[role]:3d} spill, max R{maxreg[role]:3d} | " + ", ".join(f"{k} {v}" for k, v in ops[role].most_common(10)))
