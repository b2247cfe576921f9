"""Launch list of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` run -> per-kernel
summary (launches, time, share, DRAM bytes) and the traffic record that bench.py puts into roofline.traffic.
usage: python tools/ncu_launch_summary.py <launches.csv> <summary.txt> [<traffic.json>]"""
import csv
import json
import sys
from collections import OrderedDict

src, out_txt = sys.argv[1], sys.argv[2]
out_json = sys.argv[3] if len(sys.argv) > 3 else None
rows = []
with open(src, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    rows.append(r)
per_launch = OrderedDict()   # launch id -> {kernel, metric: value}
units = {}
for r in rows:
    d = per_launch.setdefault(r["ID"], {"kernel": r["Kernel Name"]})
    try:
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    units[r["Metric Name"]] = r["Metric Unit"]
kern = OrderedDict()
for d in per_launch.values():
    k = kern.setdefault(d["kernel"], {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0, "each": []})
    k["n"] += 1
    k["ns"] += d.get("gpu__time_duration.sum", 0.0)
    k["rd"] += d.get("dram__bytes_read.sum", 0.0)
    k["wr"] += d.get("dram__bytes_write.sum", 0.0)
    k["each"].append(d)
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units.get("gpu__time_duration.sum", "ns"), 1e-6)
total = sum(k["ns"] for k in kern.values()) or 1.0
with open(out_txt, "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python bench.py "
            "--steps 2 --warmup 1 --no-cpu-baseline --no-train --no-config4 --no-full-dict (tc2, device recursion, fused compositor)\n")
    f.write("units: " + json.dumps(units, sort_keys=True) + "\n")
    for name, k in sorted(kern.items(), key=lambda kv: -kv[1]["ns"]):
        f.write(f"{k['n']:5d} launches {k['ns'] * scale:11.3f} ms {100 * k['ns'] / total:7.2f} %   dram rd {k['rd']:15.1f}  wr {k['wr']:15.1f}   {name[:90]}\n")
if out_json:
    def pick(pred):
        return [(n, k) for n, k in kern.items() if pred(n)]
    fused = pick(lambda n: "k_field_tc<2, 0, 1" in n)
    coarse = pick(lambda n: "k_field_tc<2, 0, 0" in n)
    n_rays = 640000
    rec = {"field_impl": "tc2"}
    if fused:
        name, k = fused[0]
        b = (k["rd"] + k["wr"]) / k["n"]
        rec.update({"kernel": "k_field_tc<2,0,1,0> (fused fine pass: field + compositor)", "launches": k["n"],
                    "rays_per_launch": n_rays, "points_per_launch": n_rays * 192,
                    "dram_bytes_read_per_launch_avg": k["rd"] / k["n"], "dram_bytes_write_per_launch_avg": k["wr"] / k["n"],
                    "dram_bytes_per_launch_avg": b, "dram_bytes_per_ray": b / n_rays,
                    "algorithmic_io_bytes_per_ray": "rays 32 + z_fine 768 + per-ray dir term 512 in, 58 out (compact per-ray outputs)"})
    if coarse:
        name, k = coarse[0]
        rec["coarse_kernel"] = {"kernel": "k_field_tc<2,0,0,0> (sigma-only coarse pass)",
                                "dram_bytes_per_launch_avg": (k["rd"] + k["wr"]) / k["n"]}
    rec["source"] = ("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python bench.py "
                     "--steps 2 --warmup 1 --no-cpu-baseline --no-train --no-config4 --no-full-dict (" + out_txt + ")")
    with open(out_json, "w") as f:
        json.dump(rec, f, indent=1)
print(open(out_txt).read()[:1800])
