import csv, sys
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    h = rows[0]; ik = h.index("Kernel Name"); iv = h.index("Metric Value"); ig = h.index("Grid Size"); iu = h.index("Metric Unit")
    print(f)
    for r in rows[1:][-int(12):] if False else rows[1:]:
        print("  ", r[ik][:70].ljust(70), r[ig].ljust(14), r[iv], r[iu])
