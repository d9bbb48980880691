import csv,sys,subprocess,collections,re
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; r=rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__warps_eligible.avg.per_cycle_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__cycles_elapsed.max','launch__grid_size','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed_op_shared_ld.sum','sm__inst_executed_pipe_lsu.sum']
for i,h in enumerate(hdr):
    if h in want: print(f'{h} = {r[i]} {rows[1][i]}')
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}
names=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot=collections.Counter(); ops=collections.Counter(); inst=0
for r in rows[2:]:
    if len(r)<len(hdr): continue
    for n in names:
        try: tot[n]+=int(r[idx[n]] or 0)
        except: pass
    try: n=int(r[idx['Instructions Executed']] or 0)
    except: continue
    inst+=n
    op=r[idx['Source']].split()
    if not op: continue
    o=op[0]
    if o.startswith('@') and len(op)>1: o=op[1]
    ops['.'.join(o.split('.')[:2]) if o.startswith(('LDS','STS','LDG','STG')) else o.split('.')[0]]+=n
s=sum(tot.values())
print('--- stalls'); 
for n,v in tot.most_common(9): print(f"{n:26s} {100*v/s:5.1f}%")
print('--- ops, total warp-inst', inst)
for o,v in ops.most_common(24): print(f"{o:12s} {v:10d} {100*v/inst:5.1f}%")
