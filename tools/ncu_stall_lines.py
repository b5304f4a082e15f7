"""per source line: samples of one stall reason (ncu --page source csv).  usage: ncu_stall_lines.py rep.ncu-rep stall_no_inst [top]"""
import csv, subprocess, sys
rep, col = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[2]
ic = hdr.index(col); iInst = hdr.index("Instructions Executed")
agg = {}; src = {}; tot = 0
for r in rows[3:]:
    if len(r) <= ic or r[2] != '-':
        continue
    try:
        ln = int(r[0]); v = int(r[ic]); inst = int(r[iInst])
    except ValueError:
        continue
    a = agg.setdefault(ln, [0, 0]); a[0] += v; a[1] += inst; src[ln] = r[1]; tot += v
print(col, "total", tot)
for ln, (v, inst) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{ln:5d} {100 * v / max(tot, 1):5.1f}%  inst {inst:10d}  {src[ln][:100]}")
