import csv, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: agg.setdefault(r[ki][:70],[]).append(float(r[vi].replace(',','')))
    except: pass
tot=sum(sum(v) for k,v in agg.items() if 'FillFunctor' not in k)
print(f"{'kernel':72s} {'n':>4s} {'mean us':>9s} {'share of event':>14s}")
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    sh = '   (L2 flush, outside events)' if 'FillFunctor' in k else f"{100*sum(v)/tot:13.1f}%"
    print(f"{k:72s} {len(v):4d} {sum(v)/len(v)/1e3:9.2f} {sh}")
