import csv, collections, re, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
tot=collections.defaultdict(float); cnt=collections.Counter()
for row in csv.DictReader(lines):
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    if u in('nsecond','ns'): v/=1e3
    elif u in ('msecond','ms'): v*=1e3
    name=re.sub(r'\(.*','',row['Kernel Name']); tot[name]+=v; cnt[name]+=1
T=sum(tot.values()); print(f"total {T/1e3:.1f} ms over {sum(cnt.values())} launches")
for k,v in sorted(tot.items(), key=lambda x:-x[1])[:int(sys.argv[2]) if len(sys.argv)>2 else 16]:
    print(f"{v/1e3:9.2f} ms {100*v/T:5.1f}% n={cnt[k]:5d} avg={v/cnt[k]:8.1f}us  {k[:80]}")
