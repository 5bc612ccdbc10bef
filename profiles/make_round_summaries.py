"""Rebuilds the tracked round summaries from a bench / ncu run kept under gpurun_out/:
    python profiles/make_round_summaries.py PREFIX      (e.g. gpurun_out/s37)
expects PREFIX_bench.log, PREFIX_bench_extras.log, PREFIX_launches.csv, PREFIX_prof.ncu-rep"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

pre = sys.argv[1]
here = os.path.dirname(os.path.abspath(__file__))
shutil.copy(pre + "_bench.log", os.path.join(here, "r01_bench_default.json"))
shutil.copy(pre + "_bench_extras.log", os.path.join(here, "r01_bench_extras.json"))
shutil.copy(pre + "_launches.csv", os.path.join(here, "r01_launches_bench_default.csv"))
raw = pre + "_prof_raw.csv"
with open(raw, "w") as f:
    subprocess.run(["ncu", "-i", pre + "_prof.ncu-rep", "--page", "raw", "--csv"], stdout=f, stderr=subprocess.DEVNULL)
txt = subprocess.run([sys.executable, os.path.join(here, "ncu_summary.py"), raw], capture_output=True, text=True).stdout
open(os.path.join(here, "r01_ncu_full_summary.txt"), "w").write(txt)

rows = [r for r in csv.reader(open(pre + "_launches.csv")) if len(r) > 10]
hdr = [r for r in rows if r[0] == "ID"][0]
rows = [r for r in rows if r[0].isdigit()]
ig, ik, iv = hdr.index("Grid Size"), hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows:
    if "d2pc" in r[ik]:
        agg.setdefault((r[ik][:78], r[ig]), []).append(float(r[iv]) / 1e3)
out = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 400, command: python bench.py --steps 2 --warmup 3 --no-cpu-baseline",
       "(batch 128 frames of 1080p per step: grids with 128 frames; the e2e pipeline runs chunks of 8 frames: the small grids)",
       "cold-cache, serialised launch times: compare shares, not absolutes", ""]
for (k, g), v in agg.items():
    out.append(f"{k:80s} grid {g:18s} launches {len(v):3d}  mean {sum(v) / len(v):9.1f} us")
step = [(k, g, sum(v) / len(v)) for (k, g), v in agg.items()
        if g in ("(128, 1, 1)", "(254, 128, 1)", "(2, 128, 1)", "(259200, 1, 1)", "(1, 1, 1)")]
tot = sum(x[2] for x in step)
out += ["", "one batch-128 step = sample + scan + select + status + emit = %.1f us; shares:" % tot]
out += [f"  {k.split('(')[0][:40]:42s} {t:8.1f} us  {100 * t / tot:5.1f} %" for k, g, t in step]
open(os.path.join(here, "r01_launches_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[-7:]))

blk = txt[txt.index("emit_fast_kernel"):]
rd = float(re.search(r"dram__bytes_read.sum\s+([\d.]+) Gbyte", blk).group(1))
wr = float(re.search(r"dram__bytes_write.sum\s+([\d.]+) Gbyte", blk).group(1))
json.dump({"kernel": "emit_fast_kernel<1,0,0,6> (TMA bulk stores)",
           "source": "ncu --set full --clock-control none --import-source on, python bench.py --steps 2 --warmup 3 --no-cpu-baseline (batch 128 x 1080p); summary in profiles/r01_ncu_full_summary.txt",
           "frames_per_launch": 128, "dram_bytes_read": int(rd * 1e9), "dram_bytes_write": int(wr * 1e9),
           "dram_bytes_per_frame": int((rd + wr) * 1e9 / 128), "algorithmic_bytes_per_frame": 64281600,
           "note": "measured DRAM traffic / algorithmic bytes = %.3f: no re-reads" % ((rd + wr) * 1e9 / 128 / 64281600)},
          open(os.path.join(here, "r01_emit_traffic.json"), "w"), indent=1)
d = json.loads(open(pre + "_bench.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"], d["e2e"]["value"])
e = json.loads(open(pre + "_bench_extras.log").read().strip().splitlines()[-1])["extras"]
print({k: (v.get("ms_per_step"), v.get("mpoints_out_per_s")) for k, v in e.items() if isinstance(v, dict) and "ms_per_step" in v})
