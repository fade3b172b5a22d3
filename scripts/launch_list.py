import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
h=rows[0]
ik=h.index("Kernel Name"); iv=h.index("Metric Value")
for r in rows[1:]:
    if "k2_" in r[ik] or "k1_" in r[ik]: print(r[ik][:60], r[iv])
