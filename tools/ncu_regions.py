import csv, io, subprocess, sys
path=sys.argv[1]
out = subprocess.run(["ncu","-i",path,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
lines=out.splitlines()
start=[i for i,l in enumerate(lines) if l.startswith('"Kernel Name"')][0]
rows=list(csv.reader(io.StringIO("\n".join(lines[start+1:]))))
hdr=rows[0]; si=hdr.index("Source"); ns=hdr.index("# Samples"); ie=hdr.index("Instructions Executed")
stall=[j for j,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data=[]
for i,r in enumerate(rows[1:]):
    try: data.append((i,r[si].strip(),int(r[ns]),int(r[ie] or 0),{hdr[j][6:]:int(r[j] or 0) for j in stall}))
    except Exception: pass
marks=[(i,s) for i,s,n,e,st in data if any(k in s for k in ("LDTM","STTM","UTCHMMA","UTMALDG","UTMASTG","BAR.SYNC","SYNCS.ARRIVE","TRYWAIT"))]
# regions between marks
bounds=[m[0] for m in marks]+[len(data)]
prev=0
for (idx,name),nxt in zip(marks,bounds[1:]):
    seg=[d for d in data if idx<=d[0]<nxt]
    n=sum(d[2] for d in seg); ex=sum(d[3] for d in seg)
    st={}
    for d in seg:
        for k,v in d[4].items(): st[k]=st.get(k,0)+v
    st=dict(sorted(((k,v) for k,v in st.items() if v),key=lambda x:-x[1])[:4])
    print(f"#{idx:5d} {name[:44]:44s} instrs={nxt-idx:4d} samples={n:5d} exec={ex:9d} {st}")
