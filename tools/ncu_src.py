"""Summarises an `ncu --page source --csv` dump: instructions per point-day-warp by segment and by opcode.
   python tools/ncu_src.py <csv> <points> <days>"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
N, D = int(sys.argv[2]), int(sys.argv[3])
hdr = rows[1]; data = rows[2:]
ia = hdr.index("Address"); isrc = hdr.index("Source"); iex = hdr.index("Instructions Executed"); ismp = hdr.index("# Samples")
base = int(data[0][ia], 16)
nw = N * D / 32
out = [(int(r[ia], 16) - base, int(r[iex]) / nw, int(r[ismp]), r[isrc].strip()) for r in data]
print("kernel:", rows[0][1][:100])
print("instructions per point-day (warp level): %.1f   stall samples %d" % (sum(o[1] for o in out), sum(o[2] for o in out)))
lo = 0
for i in range(1, len(out) + 1):
    if i == len(out) or abs(out[i][1] - out[lo][1]) > 0.12 * max(out[lo][1], 0.02):
        s = sum(o[1] for o in out[lo:i]); sm = sum(o[2] for o in out[lo:i])
        if s > 2:
            print("  0x%04x-0x%04x: %4d instrs, each %.2f, sum=%.1f, samples=%d" % (out[lo][0], out[i - 1][0], i - lo, out[lo][1], s, sm))
        lo = i
c = collections.Counter()
for o in out:
    op = o[3].split()[0]
    if op.startswith('@'):
        op = o[3].split()[1]
    c[op.split('.')[0]] += o[1]
print("by opcode: " + ", ".join("%s=%.1f" % (k, v) for k, v in c.most_common(32)))
cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = {hdr[i]: 0 for i in cols}
for r in data:
    for i in cols:
        try:
            tot[hdr[i]] += int(r[i])
        except ValueError:
            pass
s = sum(tot.values())
print("stalls: " + ", ".join("%s=%.1f%%" % (k[6:], 100 * v / s) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
