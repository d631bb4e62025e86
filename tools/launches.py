"""Print kernel name, grid and duration from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]
ki, vi, gi = H.index('Kernel Name'), H.index('Metric Value'), H.index('Grid Size')
for r in rows[hdr + 1:]:
    if len(r) > vi:
        print(f"{r[ki][:48]:48s} {r[gi]:>16s} {float(r[vi].replace(',', '')) / 1e3:10.2f} us")
