"""Per-source-line stall samples of a kernel from an .ncu-rep captured with --import-source on (-lineinfo build).
Usage: python tools/ncu_source_hot.py rep.ncu-rep [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
items, total, fname, hdr = [], 0, "", None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        # source text with embedded quotes breaks the CSV fields at the front of a row: index the columns from the END
        ci = {h: k - len(hdr) for k, h in enumerate(hdr) if h not in ("Source",)}
        stall_cols = [(h, k - len(hdr)) for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":      # SASS rows have an empty line number
        continue
    try:
        n, ex = int(r[ci["# Samples"]] or 0), int(r[ci["Instructions Executed"]] or 0)
    except (ValueError, KeyError):
        continue
    st = sorted(((int(r[k] or 0), h[6:]) for h, k in stall_cols), reverse=True)[:3]
    total += n
    items.append((n, ex, fname, r[0], r[1].strip()[:90], st))
print(f"total samples {total}")
for n, ex, f, l, s, st in sorted(items, reverse=True)[:top]:
    print(f"{n:7d} {100 * n / max(total, 1):5.1f}%  ex {ex:11d}  {f}:{l}  {s}   [{', '.join(f'{b} {a}' for a, b in st if a)}]")
