"""Attribute warp-stall samples of a kernel to mbarrier wait loops (by barrier smem offset) and to the rest."""
import csv, subprocess, sys, re
rep, kid = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if "# Samples" in r][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
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
rest.sort(reverse=True)
print("  non-wait samples", int(sum(x for x, _ in rest)), f"{100*sum(x for x,_ in rest)/tot:.1f} %")
for v, s in rest[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    print(f"    {int(v):8d} {100*v/tot:5.2f} %  {s[:90]}")
