"""Print the most-stalled SASS instructions of an `ncu --page source --csv` dump (developer tool)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
iS, iN, iE = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    try:
        data.append((int(r[iN] or 0), r))
    except ValueError:
        pass
tot = sum(n for n, _ in data)
print("total samples", tot, "instructions", len(data))
keep = set(id(t[1]) for t in sorted(data, key=lambda t: -t[0])[:top_n])
for n, r in data:
    if id(r) in keep:
        st = {hdr[i][6:]: int(r[i] or 0) for i in stall_cols if int(r[i] or 0) > 0}
        print(f"{n:6d} {r[iE]:>10} {r[iS].strip()[:64]:64s} {st}")
