import csv, sys, subprocess, re, collections
rep = sys.argv[1]
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); hdr=rows[0]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','lts__t_sector_hit_rate.pct','smsp__issue_active.avg.pct_of_peak_sustained_active','dram__bytes.sum.per_second','smsp__inst_executed.sum','launch__grid_size','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__warps_eligible.avg.per_cycle_active']
for w in want:
    if w in hdr:
        i=hdr.index(w); print('%-70s %s %s'%(w, rows[1][i], [r[i] for r in rows[2:]]))
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
h=rows[1]; isrc=h.index('Source'); iex=h.index('Instructions Executed'); ist=h.index('Warp Stall Sampling (All Samples)')
body=[]
for r in rows[2:]:
    if len(r)<=iex:
        if r and r[0]=='Kernel Name': break
        continue
    try: body.append((float(r[iex]), float(r[ist] or 0), r[isrc]))
    except: pass
tot=sum(b[0] for b in body); stot=sum(b[1] for b in body)
print('n sass',len(body),'total warp inst',tot,'stall samples',stot)
op=collections.Counter()
for v,s,t in body:
    m=re.sub(r'^@!?U?P\d+\s+','',t.strip()).split()[0].split('.')[0]; op[m]+=v
print(' '.join('%s:%.1f%%'%(k,100*v/tot) for k,v in op.most_common(16)))
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.012
for i,(v,s,t) in enumerate(body):
    if v/tot>thr or s/stot>0.02:
        print('%4d %5.2f%% st%5.2f%%  %s'%(i,100*v/tot,100*s/stot,t[:90]))
