"""Summarise an ncu --metrics gpu__time_duration.sum --csv launch list: per kernel name count / total us."""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[14]) / 1e3
tot = sum(a[1] for a in agg.values())
for k, (c, t) in agg.items():
    print(f"{k:40s} x{c:3d} {t:9.1f} us")
print(f"{'total':40s} x{len(rows):3d} {tot:9.1f} us")
