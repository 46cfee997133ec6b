"""Warp-stall samples per CUDA source line of an .ncu-rep captured with --import-source on:  python tools/ncu_hot_lines.py file.ncu-rep [n]"""
import csv, subprocess, sys
path = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source=cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# find header row containing "# Samples"
h = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[h]; ix = {k: i for i, k in enumerate(hdr)}
agg = {}
si = hdr.index("# Samples")
fname = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]
    if len(r) < len(hdr) or r[0] == "Line No": continue
    try: s = int(r[si])
    except ValueError: continue
    key = (fname, r[0])
    a = agg.setdefault(key, [0, r[1]])
    a[0] += s
data = [(v[0], "%s:%s" % k, v[1]) for k, v in agg.items()]
tot = sum(d[0] for d in data)
print("total samples", tot)
for s, ln, src in sorted(data, reverse=True)[:n]:
    print("%6d %5.1f%%  L%-5s %s" % (s, 100.0 * s / tot, ln, src.strip()[:130]))
