"""Instructions executed per CUDA source line from `ncu --page source --print-source cuda,sass --csv`.
   python tools/ncu_source_lines.py src_cuda.csv [top N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur_file = ""
out = []
tot = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] in ("Function Name", "Line No"):
        continue
    if r[0].isdigit() and len(r) >= 8:
        try:
            n = int(float(r[7]))
            s = int(float(r[6] or 0))
        except ValueError:
            continue
        out.append((n, s, cur_file, int(r[0]), r[1].strip()))
        tot += n
print("total", tot)
for n, s, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{n:11d} {100.0 * n / tot:5.1f}%  samp {s:6d}  {f}:{ln}  {src[:100]}")
