"""Shared-memory wavefront budget of a kernel from an `ncu --page source --csv` export: wavefronts per SASS opcode
class and the heaviest instructions, normalised per `rows` (e.g. 2^28/32 = 8388608 rows of 32 keys per pass).
usage: python tools/ncu_smem.py file.csv [rows=8388608] [top=40]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows_n = float(sys.argv[2]) if len(sys.argv) > 2 else 8388608.0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    body = rows[2:]
    for k, r in enumerate(body):          # several launches in one export: keep the first
        if r and r[0] == "Kernel Name":
            body = body[:k]
            break
    data = [r for r in body if len(r) >= len(hdr) - 2]
    wf, ideal, tag = ci["L1 Wavefronts Shared"], ci["L1 Wavefronts Shared Ideal"], ci["L1 Tag Requests Global"]
    by_op = defaultdict(lambda: [0.0, 0.0, 0.0, 0])
    items = []
    for i, r in enumerate(data):
        w, wi, tg = float(r[wf] or 0), float(r[ideal] or 0), float(r[tag] or 0)
        if w == 0 and tg == 0:
            continue
        m = re.search(r"(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ci["Source"]].strip())
        op = m.group(1) if m else "?"
        cls = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith("ATOMS") and "." in op else "")
        b = by_op[cls]
        b[0] += w; b[1] += wi; b[2] += tg; b[3] += 1
        items.append((w + tg, i, op, w, wi, tg, r[ci["Instructions Executed"]], r[ci["Source"]].strip()[:60]))
    print(f"per {rows_n:.0f} rows: class, shared wavefronts (ideal), global tag requests, #instructions")
    tot = [0.0, 0.0, 0.0]
    for cls, b in sorted(by_op.items(), key=lambda kv: -(kv[1][0] + kv[1][2])):
        print(f"  {cls:14s} {b[0] / rows_n:7.3f} ({b[1] / rows_n:6.3f})   global {b[2] / rows_n:6.3f}   n={b[3]}")
        for k in range(3):
            tot[k] += b[k]
    print(f"  {'TOTAL':14s} {tot[0] / rows_n:7.3f} ({tot[1] / rows_n:6.3f})   global {tot[2] / rows_n:6.3f}")
    print("heaviest instructions (shared wavefronts + global tag requests per row):")
    for t, i, op, w, wi, tg, ex, src in sorted(items, reverse=True)[:top]:
        print(f"  #{i:4d} {w / rows_n:6.3f} ({wi / rows_n:5.3f}) g{tg / rows_n:6.3f} exec={ex:>9s}  {src}")


main()
