import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
h=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
hdr=rows[h]; kn=hdr.index("Kernel Name"); mv=hdr.index("Metric Value")
from collections import OrderedDict
agg=OrderedDict()
for r in rows[h+1:]:
    if len(r)<=mv: continue
    name=r[kn].split("(")[0][-40:]
    agg.setdefault(name,[0,0.0]); agg[name][0]+=1; agg[name][1]+=float(r[mv].replace(",",""))
tot=sum(v[1] for v in agg.values())
for k,v in agg.items(): print("%-42s n=%4d  %10.1f us  %5.1f%%"%(k,v[0],v[1]/1000,100*v[1]/tot))
print("total us",tot/1000)
