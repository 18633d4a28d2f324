"""Summarise one `ncu --set full --import-source on` capture: launch/occupancy/stall metrics and the hottest source lines.
usage: python tools/ncu_summary.py <file.ncu-rep> [top_lines]"""
import csv,sys,subprocess,collections
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],stdout=subprocess.PIPE,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,data=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','launch__registers_per_thread','launch__block_size','launch__grid_size','launch__waves','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct','sm__inst_executed.sum ','smsp__inst_executed.sum','dram__bytes_read.sum ','dram__bytes_write.sum ','smsp__average_warps_issue_stalled','smsp__thread_inst_executed_per_inst_executed','smsp__warps_eligible.avg','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__inst_executed.sum.per_cycle_elapsed','launch__occupancy_limit']
for i,h in enumerate(hdr):
    if any(h.startswith(w.strip()) for w in want) and 'Not Issued' not in h:
        v=data[i]
        if 'stalled' in h:
            try:
                if float(v)<0.1: continue
            except: pass
        print(h,'=',v,units[i])
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],stdout=subprocess.PIPE,text=True).stdout
rows=list(csv.reader(src.splitlines()))
cur=None;hdr=None;data={}
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)>2 and r[0]=='Line No': hdr=r; continue
    if hdr and len(r)==len(hdr) and r[0].isdigit():
        d=dict(zip(hdr,r)); key=(cur,int(r[0]))
        if key in data: continue
        data[key]=(int(d['Instructions Executed']),int(d['# Samples']),r[1].strip()[:90],int(d['stall_no_inst']),int(d['stall_long_sb']),int(d['stall_short_sb']),int(d['stall_wait']),int(d['stall_barrier']))
tot=sum(x[0] for x in data.values()); ts=sum(x[1] for x in data.values())
print('total inst',tot,'samples',ts, 'noinst',sum(x[3] for x in data.values()),'lsb',sum(x[4] for x in data.values()),'ssb',sum(x[5] for x in data.values()),'wait',sum(x[6] for x in data.values()),'barrier',sum(x[7] for x in data.values()))
# region aggregation for dm_update.cu by line ranges
for (f,l),(i,s,srcl,ni,lsb,ssb,w,bar) in sorted(data.items(), key=lambda x:-x[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 30]:
    print(f'{f}:{l:4d} inst {100*i/tot:5.1f}% smp {100*s/ts:5.1f}% noinst {ni:6d} lsb {lsb:6d} ssb {ssb:6d} wait {w:6d} bar {bar:6d} | {srcl}')
