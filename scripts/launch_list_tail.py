import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10 and r[0].isdigit()]
names=[(r[4].split("(")[0].replace("void <unnamed>::","").replace("<unnamed>::",""), float(r[-1])/1e3) for r in rows]
idx=[i for i,(n,_) in enumerate(names) if n.startswith("rs_histogram_all")]
i0=idx[-1]
for n,t in names[i0:i0+11]: print(f"{t:9.1f} us  {n[:100]}")
