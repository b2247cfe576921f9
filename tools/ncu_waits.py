"""Attribute warp-stall samples of the first kernel of an .ncu-rep (source page) to mbarrier wait loops and to the rest.
usage: python tools/ncu_waits.py <rep> [top N instructions]"""
import csv, subprocess, sys, re
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
his = [i for i, r in enumerate(rows) if "# Samples" in r]
hi = his[0]
end = [i for i, r in enumerate(rows) if i > hi and r and r[0] == "Kernel Name"]
end = end[0] if end else len(rows)
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
print(rows[hi - 1][:2])
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
print("total samples", int(tot), "instructions", len(data))
seen = set()
waits = {}
for i, r in enumerate(data):
    s = r[ix["Source"]]
    if "TRYWAIT" in s:
        m = re.search(r"\[(.*?)\]", s)
        key = m.group(1) if m else s
        loop = sum(f(data[j], "# Samples") for j in range(i, min(i + 8, len(data))) if j not in seen)
        for j in range(i, min(i + 8, len(data))): seen.add(j)
        waits[key] = waits.get(key, 0) + loop
for k, v in sorted(waits.items(), key=lambda kv: -kv[1]):
    print(f"  wait {k:32s} {int(v):9d} {100*v/tot:5.1f} %")
rest = [(f(r, "# Samples"), r[ix["Source"]].strip()) for i, r in enumerate(data) if i not in seen]
print("  non-wait samples", int(sum(x for x, _ in rest)), f"{100*sum(x for x,_ in rest)/tot:.1f} %")
# by opcode
ops = {}
for v, s in rest:
    op = s.split()[0] if not s.startswith("@") else s.split()[1]
    op = op.split(".")[0]
    ops[op] = ops.get(op, 0) + v
for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:18]:
    print(f"    op {k:12s} {int(v):9d} {100*v/tot:5.1f} %")
rest.sort(reverse=True)
for v, s in rest[:topn]:
    print(f"    {int(v):8d} {100*v/tot:5.2f} %  {s[:100]}")
