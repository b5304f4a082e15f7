import csv,sys,subprocess
rep=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 40
out=subprocess.run(["ncu","-i",rep,"--page","source","--print-source","cuda,sass","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[2]
iSamp=hdr.index("# Samples"); iInst=hdr.index("Instructions Executed")
agg={}; src={}; tot_i=0; tot_s=0
for r in rows[3:]:
    if len(r)<=iInst or r[2]!='-': continue
    try: ln=int(r[0]); inst=int(r[iInst]); s=int(r[iSamp])
    except: continue
    a=agg.setdefault(ln,[0,0]); a[0]+=inst; a[1]+=s; src[ln]=r[1]; tot_i+=inst; tot_s+=s
print("total inst",tot_i,"samples",tot_s)
for ln,(i,s) in sorted(agg.items(), key=lambda kv:-kv[1][0])[:top]:
    print(f"{ln:5d} {100*i/tot_i:5.1f}% inst {100*s/tot_s:5.1f}% samp  {src[ln][:105]}")
print("--- by range")
import bisect
rngs=eval(sys.argv[3]) if len(sys.argv)>3 else []
for name,(lo,hi) in rngs:
    i=sum(v[0] for k,v in agg.items() if lo<=k<=hi); s=sum(v[1] for k,v in agg.items() if lo<=k<=hi)
    print(f"{name:24s} {100*i/tot_i:5.1f}% inst {100*s/tot_s:5.1f}% samp")
