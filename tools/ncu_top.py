import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
r=list(csv.reader(raw.splitlines())); hdr=r[0]; vals=r[2] if len(r)>2 else r[1]
want=['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','l1tex__m_xbar2l1tex_read_bytes.sum','l1tex__m_l1tex2xbar_write_bytes.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.per_cycle_active','launch__registers_per_thread','smsp__sass_inst_executed_op_tmem_ldt.sum','lts__t_sector_hit_rate.pct']
for h,v in zip(hdr,vals):
    if h in want: print(f"{h} = {v}")
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
r=list(csv.reader(src.splitlines())); hdr=r[1]; ci={h:i for i,h in enumerate(hdr)}
k=ci['Warp Stall Sampling (All Samples)']
rows=[x for x in r[2:] if len(x)>k and x[k].isdigit()]
tot=sum(int(x[k]) for x in rows)
skip=('EXIT','TRYWAIT','WARPSYNC')
print("total samples",tot)
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg={h:0 for h in stalls}
n=0
for x in sorted(rows,key=lambda x:-int(x[k])):
    s=x[ci['Source']]
    if any(t in s for t in skip) or ('BRA' in s and int(x[k])>50 and 'long_sb' in str(x)): pass
    st={h[6:]:int(x[ci[h]]) for h in stalls if x[ci[h]].isdigit() and int(x[ci[h]])>0}
    if n<int(sys.argv[2] if len(sys.argv)>2 else 30):
        print(x[k], x[ci['Address']][-5:], s[:72], st); n+=1
