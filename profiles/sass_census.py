"""Static SASS opcode census of every kernel in libxptwarp.so (cuobjdump -sass): instruction count, registers are in
`make ptxas-info`; here the opcodes that show which hardware paths a kernel uses -- TMA (UTMALDG / UTMASTG), mbarrier
(SYNCS), packed FP32 (FFMA2 / FADD2 / FMUL2), 16-byte gathers (LDG.E.128), vector reductions (REDG.E.ADD.F32x4 ...),
local-memory spills (LDL / STL), barriers.
usage: python profiles/sass_census.py [path/to/libxptwarp.so] > profiles/r02_sass_census.txt"""
import collections, os, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), '..', 'xpt-mde-2021_b200', 'xptwarp', '_lib', 'libxptwarp.so')
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
WATCH = ['UTMALDG', 'UTMASTG', 'UTMAPF', 'SYNCS', 'FFMA2', 'FADD2', 'FMUL2', 'FFMA', 'FADD', 'FMUL', 'LDG.E.128', 'LDG', 'STG', 'LDS.64',
         'LDS.128', 'LDS', 'STS', 'REDG', 'RED', 'ATOMG', 'LDL', 'STL', 'BAR', 'SHFL', 'MUFU', 'LDC', 'LDCU', 'DFMA', 'DADD', 'DMUL']
cur, per = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
        per[cur] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(.*?);', line)
    if m and cur:
        p = m.group(1).split()
        op = p[1] if p[0].startswith('@') and len(p) > 1 else p[0]
        per[cur]['_total'] += 1
        for w in WATCH:
            if op == w or op.startswith(w + '.'):
                per[cur][w] += 1
        if op.startswith('REDG') or op.startswith('RED.'):
            per[cur]['vec:' + op] += 1
for k, c in per.items():
    print(f"{k}: {c['_total']} SASS instructions")
    print('    ' + ', '.join(f"{w} {c[w]}" for w in WATCH if c[w]) + ''.join(f", {o} {n}" for o, n in c.items() if o.startswith('vec:')))
