"""Split an ncu --import-source capture of k_fused into phases by SASS address order:
segments are delimited by BAR.SYNC instructions; prints instructions executed / samples / main stalls per segment.
usage: python profiles/sass_phases.py rep.ncu-rep"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
segs = []; cur = dict(inst=0, smp=0, n=0, st=collections.Counter(), ops=collections.Counter(), first=None)
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix['Source']].strip()
    ins = int(r[ix['Instructions Executed']] or 0); smp = int(r[ix['# Samples']] or 0)
    cur['inst'] += ins; cur['smp'] += smp; cur['n'] += 1
    if cur['first'] is None: cur['first'] = src
    for s in stalls: cur['st'][s] += int(r[ix[s]] or 0)
    op = src.split()
    o = op[1] if op and op[0].startswith('@') and len(op) > 1 else (op[0] if op else '')
    cur['ops'][o.split('.')[0] + ('.' + o.split('.')[1] if o.startswith(('LDS', 'STS', 'LDG', 'STG')) and '.' in o else '')] += ins
    if 'BAR.SYNC' in src or src.startswith('EXIT'):
        segs.append(cur); cur = dict(inst=0, smp=0, n=0, st=collections.Counter(), ops=collections.Counter(), first=None)
segs.append(cur)
ti = sum(s['inst'] for s in segs); ts = sum(s['smp'] for s in segs)
print(f'total warp-inst {ti}  samples {ts}')
for k, s in enumerate(segs):
    if s['inst'] == 0 and s['smp'] == 0: continue
    st = ', '.join(f"{n[6:]} {100*v/max(1,sum(s['st'].values())):.0f}%" for n, v in s['st'].most_common(4))
    ops = ', '.join(f"{n} {v//1000}k" for n, v in s['ops'].most_common(8))
    print(f"seg {k:2d}: sass {s['n']:4d}  inst {100*s['inst']/ti:5.1f}%  samples {100*s['smp']/ts:5.1f}%  | {st}\n        {ops}")
