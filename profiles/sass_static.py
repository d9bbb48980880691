"""Static SASS census of k_fused<GRAD=1,OUT=0,DSRC=0>: instructions between BAR.SYNCs, by opcode.
usage: python profiles/sass_static.py [path/to/libxptwarp.so]"""
import subprocess, sys, collections, re, os
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), '..', 'xpt-mde-2021_b200', 'xptwarp', '_lib', 'libxptwarp.so')
txt = subprocess.run(['cuobjdump', '-sass', '-fun', '_ZN3xpt7k_fusedILb1ELb0ELb0ELi0EEEvNS_9FusedArgsE', so], capture_output=True, text=True).stdout
ins = []
for line in txt.splitlines():
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(.*?);', line)
    if m: ins.append(m.group(1).strip())
segs = [[]]
for i in ins:
    segs[-1].append(i)
    if 'BAR.SYNC' in i: segs.append([])
print('total SASS', len(ins))
def opname(i):
    p = i.split()
    o = p[1] if p[0].startswith('@') else p[0]
    q = o.split('.')
    return q[0] + ('.' + q[1] if q[0] in ('LDS', 'STS', 'LDG', 'STG', 'LDL', 'STL') and len(q) > 1 else '')
for k, s in enumerate(segs):
    c = collections.Counter(opname(i) for i in s)
    print(f"seg {k}: {len(s)} | " + ', '.join(f'{o} {n}' for o, n in c.most_common(22)))
