import csv, gzip, io, sys
path=sys.argv[1]
text=gzip.open(path,'rt').read()
lines=text.splitlines()
rows=list(csv.reader(io.StringIO("\n".join(l for l in lines if not l.startswith("#")))))
h=next(k for k,r in enumerate(rows) if r and r[0]=="Address")
hdr=rows[h]; data=[r for r in rows[h+1:] if len(r)>=len(hdr)-2]
ci={c:i for i,c in enumerate(hdr)}
cs,csrc,cex=ci["Warp Stall Sampling (All Samples)"],ci["Source"],ci["Instructions Executed"]
num=lambda x: float(x.replace(",","")) if x else 0.0
tot=sum(num(r[cs]) for r in data)
stall_cols=[i for i,c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
# group by contiguous regions with similar ex count (ratio within 1.3)
groups=[]
for k,r in enumerate(data):
    ex=num(r[cex])
    if groups and groups[-1]['exs'] and (0.7 < (ex+1)/(groups[-1]['ex']+1) < 1.4):
        g=groups[-1]
    else:
        g={'start':k,'ex':ex,'exs':[], 'samples':0,'n':0,'stalls':{}, 'ops':{}}
        groups.append(g)
    g['exs'].append(ex); g['ex']=sum(g['exs'])/len(g['exs']); g['samples']+=num(r[cs]); g['n']+=1; g['end']=k
    for i in stall_cols:
        g['stalls'][hdr[i]]=g['stalls'].get(hdr[i],0)+num(r[i])
    op=r[csrc].split()
    op=(op[1] if op and op[0].startswith('@') else (op[0] if op else '?')).split('.')[0]
    g['ops'][op]=g['ops'].get(op,0)+1
print(lines[0])
for g in groups:
    if g['samples']/tot<0.004: continue
    st=sorted(((v,k.replace('stall_','')) for k,v in g['stalls'].items() if v),reverse=True)[:4]
    ops=sorted(((v,k) for k,v in g['ops'].items()),reverse=True)[:6]
    print("instr %4d-%4d n=%4d ex/instr=%10.0f samples=%5.1f%%  %s | %s"%(g['start'],g['end'],g['n'],g['ex'],100*g['samples']/tot," ".join("%s=%.0f%%"%(k,100*v/g['samples']) for v,k in st)," ".join("%s:%d"%(k,v) for v,k in ops)))
