"""Condense `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K` (SASS view): per instruction the
executed count (per warp of the launch) and its share of stall samples, to find where a kernel's issue slots
and latencies go.  usage: python profiles/ncu_srcpage.py file.csv [min_share_pct]"""
import csv
import sys


def main(path, min_share=0.0, top=None):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ia, isrc, ismp, iex = (hdr.index(x) for x in ("Address", "Source", "# Samples", "Instructions Executed"))
    data = [r for r in rows[2:] if len(r) > iex and r[ia] != "Address"]
    tot_s = sum(float(r[ismp] or 0) for r in data) or 1.0
    tot_i = sum(float(r[iex] or 0) for r in data)
    warps = max(float(r[iex] or 0) for r in data[:3])  # the prologue runs once per warp
    print("total warp instructions %.0f ; per warp %.1f ; samples %.0f" % (tot_i, tot_i / warps, tot_s))
    cum = 0.0
    for n, r in enumerate(data):
        ex = float(r[iex] or 0)
        sh = 100.0 * float(r[ismp] or 0) / tot_s
        cum += ex
        if sh >= min_share:
            print("%4d %-70s exec/warp %6.2f  cum %7.1f  stall%% %5.2f" % (n, r[isrc].strip()[:70], ex / warps, cum / warps, sh))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.0)
