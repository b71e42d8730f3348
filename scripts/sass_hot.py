"""Summarise an `ncu --page source --csv` export: opcode mix and the hottest SASS lines (stall samples)."""
import csv, sys, collections
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if len(r)>3 and r[0]=='Address'][0]
hdr=rows[hi]
ia=hdr.index('Source'); isamp=hdr.index('Warp Stall Sampling (All Samples)'); iex=hdr.index('Instructions Executed')
def num(x):
    try: return int(x)
    except: return 0
data=[(num(r[isamp]), num(r[iex]), r[ia].strip()) for r in rows[hi+1:] if len(r)>iex and r[0].startswith('0x')]
tot=sum(d[0] for d in data) or 1; totex=sum(d[1] for d in data) or 1
print("total samples", tot, "total warp-inst", totex, "sass lines", len(data))
ex=collections.Counter(); sm=collections.Counter()
for s,e,src in data:
    t=src.split()
    op=(t[1] if t and t[0].startswith('@') and len(t)>1 else (t[0] if t else '?')).split('.')[0]
    ex[op]+=e; sm[op]+=s
print("opcode mix (executed / stall samples):")
for k,v in ex.most_common(int(sys.argv[2]) if len(sys.argv)>2 else 16): print("  %-10s %5.1f%%  %5.1f%%"%(k,100*v/totex,100*sm[k]/tot))
print("hottest lines:")
for s,e,src in sorted(data,reverse=True)[:int(sys.argv[3]) if len(sys.argv)>3 else 22]: print("  %5.1f%% ex=%9d  %s"%(100*s/tot,e,src[:100]))
