# Per-CUDA-line aggregation of an ncu report's source page: stall samples and executed instructions.  usage: ncu_lines.py report.ncu-rep [top]
import sys, csv, collections, subprocess, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
for i, r in enumerate(rows):
    if r and r[0] == "Line No": hdr = r; start = i + 1; break
ci = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
per = collections.Counter(); samp = collections.Counter(); src = {}; cur = None
for r in rows[start:]:
    if r and r[0].strip().isdigit():
        cur = int(r[0]); src[cur] = ",".join(r[1:])[:120]; continue
    if len(r) > ii and r[2].startswith("0x"):
        try: per[cur] += int(r[ii]); samp[cur] += int(r[ci])
        except ValueError: pass
ts, ti = sum(samp.values()), sum(per.values())
print(f"total samples {ts}, warp instructions {ti}")
for l, n in samp.most_common(top):
    print(f"{l:5d} {100*n/ts:5.1f}% samples {100*per[l]/ti:5.1f}% inst  {src[l]}")
if len(sys.argv) > 3:   # ranges "a-b,c-d"
    for rg in sys.argv[3].split(","):
        a, b = map(int, rg.split("-"))
        print(f"lines {a}-{b}: {100*sum(v for k, v in samp.items() if a <= k <= b)/ts:5.1f}% samples {100*sum(v for k, v in per.items() if a <= k <= b)/ti:5.1f}% inst")
