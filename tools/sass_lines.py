"""Join an `ncu --page source --print-source sass --csv` export with `nvdisasm -g` line info: executed warp instructions
and stall samples per source line.  usage: sass_lines.py <sass.csv> <nvdisasm -g text of the same function> [top]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hdr]
ia, ie, isamp, ithr = H.index("Address"), H.index("Instructions Executed"), H.index("# Samples"), H.index("Thread Instructions Executed")
stall_cols = {h: i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h}
prof = []
for r in rows[hdr + 1:]:
    if len(r) <= ie: continue
    prof.append(r)
base = int(prof[0][ia], 16)
line_of = {}
cur = None
chain = False
for l in open(sys.argv[2]):
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        # with `nvdisasm -gi` an instruction is preceded by its whole inline chain, innermost first: keep the outermost
        if m.group(3):
            cur = (m.group(3).split("/")[-1], int(m.group(4)))
        elif not chain:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
        chain = True
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
        chain = False
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
tot = [0, 0]
for r in prof:
    off = int(r[ia], 16) - base
    k = line_of.get(off, ("?", 0))
    e, s, t = int(r[ie] or 0), int(r[isamp] or 0), int(r[ithr] or 0)
    a = agg[k]
    a[0] += e; a[1] += s; a[2] += t
    for h, i in stall_cols.items():
        v = int(r[i] or 0)
        if v: a[3][h[6:]] += v
    tot[0] += e; tot[1] += s
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
print(f"total warp inst {tot[0]}, samples {tot[1]}")
for k, a in sorted(agg.items(), key=(lambda kv: kv[0]) if top < 0 else (lambda kv: -kv[1][1]))[:abs(top)]:
    st = " ".join(f"{n}:{v}" for n, v in a[3].most_common(3))
    print(f"{k[0]}:{k[1]:<5d} inst {a[0]:>9d} ({100*a[0]/tot[0]:5.1f}%) lanes {a[2]/max(a[0],1):5.1f}  samples {a[1]:>6d} ({100*a[1]/tot[1]:5.1f}%)  {st}")
