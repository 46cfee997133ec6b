"""Top stall sites of a kernel from an .ncu-rep captured with --import-source on:  python tools/ncu_hot_sass.py file.ncu-rep [n]
Prints the SASS instructions with the most warp-stall samples and their dominant stall reason."""
import csv, subprocess, sys
path = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source=sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: s = int(r[ix["# Samples"]])
    except ValueError: continue
    st = {h: int(r[ix[h]] or 0) for h in stall_cols}
    data.append((s, r[ix["Address"]], r[ix["Source"]], st))
tot = sum(d[0] for d in data)
print("total samples", tot)
agg = {}
for s, a, src, st in data:
    for k, v in st.items(): agg[k] = agg.get(k, 0) + v
print("by reason:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for i, (s, a, src, st) in enumerate(data):
    data[i] = (s, a, src, st, i)
for s, a, src, st, i in sorted(data, reverse=True)[:n]:
    top = max(st, key=st.get)
    print("%6d %5.1f%%  #%-5d %-60s %s" % (s, 100.0 * s / tot, i, src[:60], top))
if len(sys.argv) > 3:                      # context: python tools/ncu_hot_sass.py file n idx [idx ...]
    byi = {d[4]: d for d in data}
    for c in sys.argv[3:]:
        c = int(c)
        print("---- around #%d" % c)
        for i in range(c - 8, c + 4):
            if i in byi: print("%6d  #%-5d %s" % (byi[i][0], i, byi[i][2][:110]))
