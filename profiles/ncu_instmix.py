"""Dynamic SASS instruction mix of one kernel from `ncu --page source --csv` output."""
import collections
import csv
import sys


def main(path, per_thread_items=4):
    rows = list(csv.reader(open(path)))
    # find header rows (one per profiled launch); use the first launch only
    hidx = [i for i, r in enumerate(rows) if 'Instructions Executed' in r]
    hdr = rows[hidx[0]]
    end = hidx[1] - 1 if len(hidx) > 1 else len(rows)
    data = [r for r in rows[hidx[0] + 1:end] if len(r) == len(hdr)]
    ia, ie, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    warps = int(data[0][ie])
    by, samp, tot = collections.Counter(), collections.Counter(), 0
    for r in data:
        n = int(r[ie])
        src = r[ia].split()
        op = (src[1] if src[0].startswith('@') else src[0]).split('.')[0]
        by[op] += n
        tot += n
        samp[op] += int(r[isamp] or 0)
    print(f"warps {warps}  warp-instr/warp {tot / warps:.1f}  per item {tot / warps / per_thread_items:.1f}")
    for op, n in by.most_common(34):
        print(f"{op:10s} {n / warps:8.1f} per thread   stall samples {samp[op]}")
    print("hottest lines:")
    for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:14]:
        print(f"  {r[isamp]:>6s} {r[ie]:>9s}  {r[ia][:100]}")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 4)
