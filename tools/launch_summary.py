"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/launch_summary.py file.csv [last_n]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
if len(sys.argv) > 2:
    rows = rows[-int(sys.argv[2]):]
agg = collections.OrderedDict()
tot = 0.0
for r in rows:
    v = float(r[-1].replace(',', '')); u = r[-2]
    ms = v / 1e6 if u == 'ns' else (v / 1e3 if u == 'us' else (v * 1e3 if u == 's' else v))
    k = r[4].split('(')[0][:50]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms; tot += ms
for k, (n, ms) in agg.items():
    print(f"{k:52s} x{n:4d} {ms:9.3f} ms  {100 * ms / tot:5.1f} %")
print(f"{'total':52s} x{len(rows):4d} {tot:9.3f} ms")
