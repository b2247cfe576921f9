"""Top stall-sample instructions of the (single) kernel in an .ncu-rep, with mbarrier wait loops grouped."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if "# Samples" in r][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
S = lambda r: float(r[ix["# Samples"]] or 0)
tot = sum(S(r) for r in data)
print("total samples", int(tot), "instructions", len(data))
# group: consecutive window around TRYWAIT
i = 0; groups = []
while i < len(data):
    s = data[i][ix["Source"]]
    if "TRYWAIT" in s:
        j0 = max(0, i - 1); j1 = min(len(data), i + 6)
        groups.append((sum(S(data[j]) for j in range(j0, j1)), i, s.strip()))
        i = j1
    else:
        i += 1
print("-- mbarrier wait loops (samples, sass index)")
for v, i, s in sorted(groups, reverse=True)[:12]:
    print(f"  {int(v):8d} {100*v/tot:5.1f}%  #{i}  {s[:100]}")
print("-- top instructions")
for v, i, s in sorted(((S(r), i, r[ix['Source']].strip()) for i, r in enumerate(data)), reverse=True)[:top]:
    print(f"  {int(v):8d} {100*v/tot:5.1f}%  #{i}  {s[:110]}")
