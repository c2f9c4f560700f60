"""Summarise an ncu --page source --csv dump: top stalled SASS lines and overall stall reasons."""
import csv, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]; data = rows[hi + 1:]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = 0; big = []
for k, r in enumerate(data):
    if len(r) < len(hdr): continue
    try: n = int(r[col['# Samples']])
    except ValueError: continue
    tot += n
    st = {s: int(r[col[s]] or 0) for s in stalls if int(r[col[s]] or 0) > 0}
    big.append((n, k, r[col['Address']], r[col['Source']][:72], int(r[col['Instructions Executed']] or 0), st))
print("total samples", tot)
for n, k, a, s, ie, st in sorted(big, key=lambda x: (-x[0], x[1]))[:topn]:
    top = sorted(st.items(), key=lambda x: -x[1])[:3]
    print(f"{n:6d} {n/tot*100:5.1f}% #{k:4d} {s:72s} ie={ie:8d} {top}")
allst = {}
for n, k, a, s, ie, st in big:
    for kk, v in st.items(): allst[kk] = allst.get(kk, 0) + v
print(sorted(allst.items(), key=lambda x: -x[1]))
