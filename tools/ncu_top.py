"""Summarise an `ncu --page source --csv` export: stall reasons overall and the hottest SASS lines.
usage: python tools/ncu_top.py file.csv [top=25]"""
import csv
import sys

def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    body = rows[2:]
    for k, r in enumerate(body):          # several launches in one export: keep the first
        if r and r[0] == "Kernel Name":
            body = body[:k]
            break
    data = [r for r in body if len(r) >= len(hdr) - 2]
    ci = {h: i for i, h in enumerate(hdr)}
    samp = ci["# Samples"]
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[samp] or 0) for r in data)
    print("total samples", total, "instructions", len(data))
    agg = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in stall_cols}
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
        print(f"  {k:28s} {v:8d} {100.0 * v / max(total, 1):5.1f}%")
    print("hottest instructions:")
    order = sorted(range(len(data)), key=lambda i: -int(data[i][samp] or 0))[:top]
    for i in sorted(order):
        r = data[i]
        reasons = sorted(((int(r[c] or 0), hdr[c]) for c in stall_cols), reverse=True)[:2]
        rs = ", ".join(f"{n}:{c.replace('stall_', '')}" for c, n in [(x[1], x[0]) for x in reasons] if n)
        print(f"  #{i:4d} {int(r[samp]):7d} {100.0 * int(r[samp]) / max(total, 1):5.1f}%  {r[ci['Source']].strip()[:70]:70s} exec={r[ci['Instructions Executed']]:>9s} [{rs}]")

main()
