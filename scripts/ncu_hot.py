"""Summarise `ncu --page source --csv` output: top SASS instructions by stall samples, per kernel section.
usage: ncu_hot.py file.csv [section_index] [top_n]"""
import csv, sys
lines = open(sys.argv[1]).read().split("\n")
starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
sec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
beg = starts[sec]
end = starts[sec + 1] if sec + 1 < len(starts) else len(lines)
rows = list(csv.reader(lines[beg:end]))
print(rows[0][1][:100])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot, agg, data = 0, {s: 0 for s in stalls}, []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    n = int(r[ix["# Samples"]]); tot += n
    for s in stalls: agg[s] += int(r[ix[s]] or 0)
    data.append((n, r))
print("total samples", tot, sorted(((v, k) for k, v in agg.items()), reverse=True)[:7])
data.sort(key=lambda t: -t[0])
for n, r in data[:topn]:
    top = sorted(((int(r[ix[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{n:7d} {100*n/max(tot,1):5.1f}% exec={r[ix['Instructions Executed']]:>9} {r[ix['Source']].strip()[:60]:60s} {top}")
