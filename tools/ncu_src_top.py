"""Top stall sites from `ncu -i rep --page source --csv --kernel-name regex:X` output (multi-launch aware)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 14
blocks, cur = [], None
for r in rows:
    if 'Address' in r and 'Source' in r:
        cur = {"hdr": r, "data": []}
        blocks.append(cur)
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for bi, b in enumerate(blocks[:1]):
    hdr, data = b["hdr"], b["data"]
    si, so = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Source')
    stall_cols = [j for j, h in enumerate(hdr) if h.startswith('stall_')]

    def f(x):
        try:
            return float(x)
        except ValueError:
            return 0.0
    tot = sum(f(r[si]) for r in data) or 1.0
    print("launch %d: %d samples" % (bi, tot))
    for r in sorted(data, key=lambda r: -f(r[si]))[:N]:
        st = sorted([(f(r[j]), hdr[j][6:]) for j in stall_cols], reverse=True)[:2]
        print('%6.1f%%  %-100s %s' % (100 * f(r[si]) / tot, r[so][:100], ["%s:%d" % (n, v) for v, n in st]))
