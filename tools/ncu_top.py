"""Aggregate an `ncu --page source --csv` dump: top SASS instructions by stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc = hdr.index('Source'); ist = hdr.index('Warp Stall Sampling (All Samples)'); iex = hdr.index('Instructions Executed')
data = []
for r in rows[2:]:
    try:
        data.append((int(r[ist] or 0), r[isrc][:120], int(r[iex] or 0)))
    except Exception:
        pass
tot = sum(d[0] for d in data) or 1
print('total samples', tot)
for d in sorted(data, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{d[0]:6d} {100*d[0]/tot:5.1f}%  ex={d[2]:7d}  {d[1]}")
