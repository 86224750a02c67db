"""Top warp-stall lines of one kernel from `ncu -i rep --page source --csv` output (stdin or file)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
st = {h: i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h}
tot = sum(int(r[isamp]) for r in data)
print('total samples', tot, 'instructions', len(data))
agg = {h: sum(int(r[i]) for r in data) for h, i in st.items()}
print('by reason:', {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for k, r in sorted(enumerate(data), key=lambda kr: -int(kr[1][isamp]))[:n]:
    why = max(st, key=lambda h: int(r[st[h]]))
    print(f"{k:5d} {r[isamp]:>6} {100 * int(r[isamp]) / tot:5.1f}%  ex={r[iex]:>8} {why:14s} {r[isrc].strip()[:100]}")
