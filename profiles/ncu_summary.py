import csv,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','launch__registers_per_thread','launch__occupancy_limit','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__throughput.avg','launch__shared_mem_per_block','launch__grid_size','launch__block_size','smsp__issue_active.avg.pct','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__average_warps_issue_stalled','lts__t_sector_hit_rate','l1tex__t_bytes','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','lts__throughput']
for h,u,v in zip(hdr,units,vals):
    if any(w in h for w in want) and 'per_second' not in h and not h.endswith('.max') and '.min' not in h and '.max' not in h: print('%-95s %-12s %s'%(h,u,v))
mix=subprocess.run(['ncu','-i',rep,'--page','source','--print-source','sass,cuda','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(mix.splitlines()))
cur=None; agg=[]; h2=None
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)>5 and r[0]=='Line No': h2=r; continue
    if h2 and len(r)>7 and r[0].isdigit():
        s=int(r[6]) if r[6].isdigit() else 0
        i=int(r[7]) if r[7].isdigit() else 0
        agg.append((s,i,cur,int(r[0]),r[1].strip()[:105]))
tot=sum(a[0] for a in agg)
print('total samples',tot)
for a in sorted(agg,reverse=True)[:int(sys.argv[2]) if len(sys.argv)>2 else 30]:
    print('%6d %5.1f%% inst=%10d %s:%d  %s'%(a[0],100*a[0]/tot,a[1],a[2],a[3],a[4]))
