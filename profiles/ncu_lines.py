"""Summarise an ncu report by CUDA source line: python profiles/ncu_lines.py <rep> <kernel-regex> [top]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
L = []
for r in rows:
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[2] != "-":  # per-SASS rows; keep the per-line aggregate rows (Address == "-")
        continue
    try:
        n = float(r[hdr.index("Instructions Executed")])
        s = float(r[hdr.index("# Samples")])
        ti = float(r[hdr.index("Avg. Threads Executed")])
    except Exception:
        continue
    L.append((n, s, ti, r[0], r[1].strip()[:110]))
tot = sum(x[0] for x in L) or 1
tots = sum(x[1] for x in L) or 1
print(f"total warp-instructions {tot:.3e}, samples {tots:.0f}")
for n, s, ti, ln, src in sorted(L, key=lambda x: -x[1])[:top]:
    print(f"{n / tot * 100:5.1f}% inst {s / tots * 100:5.1f}% samp thr={ti:4.0f} L{ln:>4s}  {src}")
