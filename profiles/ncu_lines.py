import csv,sys,subprocess
rep=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 50
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
out=[]; hdr=None; fname=None
for r in rows:
    if not r: continue
    if r[0]=='File Path': fname=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': continue
    if r[0]=='Line No': hdr=r; idx={}; 
    if r[0]=='Line No':
        for i,h in enumerate(r):
            idx.setdefault(h,i)
        continue
    if hdr is None or len(r)<len(hdr): continue
    if r[0]=='' : continue
    try:
        ln=int(r[0]); n=int(r[idx['Instructions Executed']] or 0); smp=int(r[idx['# Samples']] or 0)
    except: continue
    out.append((fname,ln,n,smp,r[1]))
tot=sum(o[2] for o in out); stot=sum(o[3] for o in out)
print('total inst',tot,'samples',stot)
for f,ln,n,smp,s in sorted(out,key=lambda x:-x[3])[:top]:
    print(f"{f[:14]:14s}:{ln:4d} {100*n/tot:5.1f}% inst {100*smp/max(stot,1):5.1f}% smp | {s.strip()[:100]}")
# phase buckets
def bucket(f,ln):
    if f.startswith('xpt_fused'):
        if ln<295: return 'prologue(tile load, smooth, xstats)'
        if ln<=357: return 'Y phase'
        if ln<=422: return 'S phase'
        if ln<=545: return 'G phase'
        return 'epilogue'
    if f.startswith('xpt_kernels'):
        if 225<=ln<=300: return 'Y phase'
        if ln<=50: return 'reduce'
        return 'kernels.cuh other'
    return 'intrinsics/other'
import collections
bi=collections.Counter(); bs=collections.Counter()
for f,ln,n,smp,s in out:
    b=bucket(f,ln); bi[b]+=n; bs[b]+=smp
print('--- phases')
for b in bi: print(f"{b:40s} inst {100*bi[b]/tot:5.1f}%  samples {100*bs[b]/stot:5.1f}%")
