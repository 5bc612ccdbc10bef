"""Rebuilds the tracked round summaries from a bench / ncu run kept under gpurun_out/ (profiles/refresh_round.sh):
    python profiles/make_round_summaries.py PREFIX [ROUND]     (e.g. gpurun_out/r02 r02)
expects PREFIX_bench_default.json, PREFIX_bench_reference_arm.json, PREFIX_launches.csv, PREFIX_prof.ncu-rep"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

pre = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r02"
here = os.path.dirname(os.path.abspath(__file__))
def last_json_line(path):
    return [l for l in open(path).read().strip().splitlines() if l.startswith("{")][-1]


open(os.path.join(here, rnd + "_bench_default.json"), "w").write(last_json_line(pre + "_bench_default.json") + "\n")
open(os.path.join(here, rnd + "_bench_reference_arm.json"), "w").write(last_json_line(pre + "_bench_reference_arm.json") + "\n")
shutil.copy(pre + "_launches.csv", os.path.join(here, rnd + "_launches_bench_default.csv"))
raw = pre + "_prof_raw.csv"
with open(raw, "w") as f:
    subprocess.run(["ncu", "-i", pre + "_prof.ncu-rep", "--page", "raw", "--csv"], stdout=f, stderr=subprocess.DEVNULL)
txt = subprocess.run([sys.executable, os.path.join(here, "ncu_summary.py"), raw], capture_output=True, text=True).stdout
open(os.path.join(here, rnd + "_ncu_full_summary.txt"), "w").write(txt)

rows = [r for r in csv.reader(open(pre + "_launches.csv")) if len(r) > 10]
hdr = [r for r in rows if r[0] == "ID"][0]
rows = [r for r in rows if r[0].isdigit()]
ig, ik, iv = hdr.index("Grid Size"), hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows:
    if "d2pc" in r[ik]:
        agg.setdefault((r[ik][:78], r[ig]), []).append(float(r[iv]) / 1e3)
out = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 400, command: python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras",
       "(batch 128 frames of 1080p per step: grids with 128 frames; the e2e pipeline runs chunks of 8 frames: the small grids)",
       "cold-cache, serialised launch times: compare shares, not absolutes", ""]
for (k, g), v in agg.items():
    out.append(f"{k:80s} grid {g:18s} launches {len(v):3d}  mean {sum(v) / len(v):9.1f} us")
step = [(k, g, sum(v) / len(v)) for (k, g), v in agg.items()
        if (g in ("(128, 1, 1)", "(254, 128, 1)", "(2, 128, 1)", "(259200, 1, 1)", "(1, 1, 1)") or
            ("scan_native" in k and g.endswith(", 128, 1)"))) and "FillFunctor" not in k]
tot = sum(x[2] for x in step)
out += ["", "one batch-128 step = sample + scan + select + status + emit = %.1f us; shares:" % tot]
out += [f"  {k.split('(')[0][:40]:42s} {t:8.1f} us  {100 * t / tot:5.1f} %" for k, g, t in step]
open(os.path.join(here, rnd + "_launches_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[-7:]))

blk = txt[txt.index("emit_fast_kernel"):]
rd = float(re.search(r"dram__bytes_read.sum\s+([\d.]+) Gbyte", blk).group(1))
wr = float(re.search(r"dram__bytes_write.sum\s+([\d.]+) Gbyte", blk).group(1))
json.dump({"kernel": "emit_fast_kernel<1,0,0,6> (TMA bulk stores)",
           "source": "ncu --set full --clock-control none --import-source on, python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras (batch 128 x 1080p); summary in profiles/%s_ncu_full_summary.txt" % rnd,
           "frames_per_launch": 128, "dram_bytes_read": int(rd * 1e9), "dram_bytes_write": int(wr * 1e9),
           "dram_bytes_per_frame": int((rd + wr) * 1e9 / 128), "algorithmic_bytes_per_frame": 64281600,
           "note": "measured DRAM traffic / algorithmic bytes = %.3f: no re-reads" % ((rd + wr) * 1e9 / 128 / 64281600)},
          open(os.path.join(here, rnd + "_emit_traffic.json"), "w"), indent=1)
d = json.loads(last_json_line(pre + "_bench_default.json"))
print(d["value"], d["ms_per_step"], d["roofline"], d["e2e"]["value"])
print({k: (v.get("ms_per_step"), v.get("frac_of_measured_peak")) for k, v in d.get("extras", {}).items() if isinstance(v, dict) and "ms_per_step" in v})
