"""Per CUDA source line aggregation of an ncu report's warp-stall samples and executed instructions:
  python tools/ncu_lines.py X.ncu-rep [top_n] [file-substring]
Uses `ncu -i X --page source --csv --print-source cuda,sass` (needs -lineinfo at compile time; works without a GPU)."""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
filt = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = defaultdict(lambda: [0, 0, 0, ""])  # samples, non-barrier samples, instructions
cur_file = ""
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        # two columns are called "Source": the CUDA line text (index 1) and the SASS text (index 3)
        continue
    if hdr is None or len(r) < 8 or not r[0].strip().isdigit():
        continue
    line = int(r[0])
    key = (cur_file, line)
    def num(name):
        i = hdr.get(name)
        try:
            return int(r[i]) if i is not None and r[i] else 0
        except ValueError:
            return 0
    s = num("# Samples")
    bar = num("stall_barrier")
    a = agg[key]
    a[0] += s
    a[1] += s - bar
    a[2] += num("Instructions Executed")
    if r[1].strip():
        a[3] = r[1].strip()[:110]
tot = sum(a[0] for a in agg.values())
totnb = sum(a[1] for a in agg.values())
print(f"samples {tot}, without CTA-barrier stalls {totnb}")
items = [(k, v) for k, v in agg.items() if filt in k[0]]
for (f, l), (s, nb, ins, txt) in sorted(items, key=lambda kv: -kv[1][1])[:top]:
    print(f"{f}:{l:<5d} nb={nb:5d} ({100.0 * nb / max(totnb, 1):4.1f}%) all={s:5d} inst={ins:8d}  {txt}")
