"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (us per step and share)."""
import collections
import csv
import sys


def main(path, steps):
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hdr_i]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr_i + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0][:72]
        v = float(r[mv].replace(",", ""))
        if r[mu] == "ns":
            v /= 1000
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"total {tot / steps:.1f} us/step over {steps} steps ({sum(v[0] for v in agg.values()) / steps:.0f} launches/step)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / steps:9.1f} us/step {100 * v[1] / tot:5.1f}%  x{v[0] / steps:4.1f}  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
