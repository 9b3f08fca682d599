import csv,sys,subprocess,collections
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
H=rows[0]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','launch__waves_per_multiprocessor','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','SM_A.TriageCompute.sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    d=dict(zip(H,r))
    print(d.get('Kernel Name','')[:70])
    for k in keys:
        if k in d: print('  %-90s %s'%(k,d[k]))
    for k,v in d.items():
        if 'issue_stalled' in k and 'per_issue_active' in k and float(v or 0)>0.15: print('  %-90s %s'%(k.replace('smsp__average_warps_issue_stalled_',''),v))
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
H=None
def flush():
    if not ops: return
    print(kname, 'total warp instr',tot)
    for op,n in ops.most_common(int(sys.argv[2]) if len(sys.argv)>2 else 14): print('  %-18s %12d %5.1f%%  pred-on warps %12d  samples %d'%(op,n,100*n/tot,pon[op]//32,samp[op]))
ops=collections.Counter(); pon=collections.Counter(); samp=collections.Counter(); tot=0; kname=''
for r in rows:
    if len(r)>=1 and r[0]=='Kernel Name':
        flush(); ops=collections.Counter(); pon=collections.Counter(); samp=collections.Counter(); tot=0; kname=r[1][:60]; continue
    if 'Source' in r and 'Instructions Executed' in r:
        H=r; ia=H.index('Source'); ie=H.index('Instructions Executed'); ip=H.index('Predicated-On Thread Instructions Executed'); isamp=H.index('# Samples'); continue
    if H is None or len(r)<=ip: continue
    t=r[ia].strip().split()
    if not t: continue
    op=t[1] if t[0].startswith('@') else t[0]
    try: n=int(r[ie])
    except ValueError: continue
    ops[op]+=n; pon[op]+=int(r[ip]); samp[op]+=int(r[isamp]); tot+=n
flush()
