"""Summarise an ncu --page source --csv dump: top SASS instructions by stall samples + stall-reason totals."""
import csv, sys, collections
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr) and (r[idx["# Samples"]] or "0").isdigit()]
tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.Counter()
for r in data:
    for h in reasons:
        agg[h] += int(r[idx[h]] or 0)
print("stall reasons:", ", ".join("%s=%.1f%%" % (k, 100.0 * v / max(tot, 1)) for k, v in agg.most_common(8)))
data.sort(key=lambda r: -int(r[idx["# Samples"]] or 0))
for r in data[:top]:
    n = int(r[idx["# Samples"]] or 0)
    main = max(reasons, key=lambda h: int(r[idx[h]] or 0))
    print("%5.1f%%  %-18s %s" % (100.0 * n / max(tot, 1), main, r[idx["Source"]][:110]))
