"""Per-source-line summary of an ncu report's source page (stall samples, instructions):
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X.csv ; python tests/ncu_lines.py X.csv [top]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = []
h = None
for r in rows:
    if r and r[0] == "Line No":
        h = r
        continue
    if h is None or len(r) < 10 or not r[0] or not r[0].isdigit():
        continue
    g = lambda k: int(r[h.index(k)]) if r[h.index(k)].lstrip("-").isdigit() else 0
    st = {k[6:]: g(k) for k in h if k.startswith("stall_") and "Not Issued" not in k}
    out.append((g("# Samples"), int(r[0]), g("Instructions Executed"), r[1].strip()[:100], st))
tot = sum(o[0] for o in out)
print("total samples", tot)
for s, ln, ex, src, st in sorted(out, key=lambda o: -o[0])[:top]:
    best = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{ln:4d} {s:5d} {100 * s / max(tot, 1):5.1f}% ex {ex:7d} {dict((k, v) for k, v in best if v)} | {src}")
