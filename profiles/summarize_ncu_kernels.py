"""Per-kernel table from an `ncu --metrics ... --csv` capture of bench.py (long format: one row per launch x metric):
launches/step, average duration, DRAM bytes and GB/s against the measured HBM peak, tensor-pipe activity.

    python profiles/summarize_ncu_kernels.py gpurun_out/r1_ncu_step.csv STEPS > profiles/r01_ncu_per_kernel.txt
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(path, steps):
    peak = 6650.0
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    c_id, c_k, c_n, c_u, c_v = (hdr.index(s) for s in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    c_g = hdr.index("Grid Size")
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= c_v:
            continue
        d = launches.setdefault(r[c_id], {"name": r[c_k].split("(")[0].replace("void ", "")[:44], "grid": r[c_g]})
        try:
            v = float(r[c_v].replace(",", ""))
        except ValueError:
            continue
        unit = r[c_u]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[r[c_n]] = v * scale
    agg = collections.OrderedDict()
    for d in launches.values():
        a = agg.setdefault(d["name"], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "tensor": 0.0, "tc": 0.0, "l2hit": 0.0, "max_us": 0.0})
        t = d.get("gpu__time_duration.sum", 0.0)
        a["n"] += 1
        a["us"] += t
        a["max_us"] = max(a["max_us"], t)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
        a["tensor"] += t * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
        a["tc"] += t * d.get("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
        a["l2hit"] += t * d.get("lts__t_sector_hit_rate.pct", 0.0)
    tot = sum(a["us"] for a in agg.values())
    print(f"# {os.path.basename(path)}: {len(launches)} launches over {steps} steps, {tot / steps:.1f} us of kernel time per step "
          f"(ncu: serialised, cold caches -> compare shares); HBM peak {peak:.1f} GB/s (measured)")
    print(f"{'kernel':44s} {'x/step':>6s} {'us/step':>8s} {'share':>6s} {'max us':>7s} {'DRAM MB/launch':>14s} {'GB/s':>7s} {'of peak':>7s} "
          f"{'tensor%':>7s} {'tc%':>6s} {'L2hit%':>6s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        by = a["rd"] + a["wr"]
        gbs = by / a["us"] / 1e3 if a["us"] > 0 else 0.0
        print(f"{k:44s} {a['n'] / steps:6.1f} {a['us'] / steps:8.1f} {100 * a['us'] / tot:5.1f}% {a['max_us']:7.1f} {by / a['n'] / 1e6:14.2f} "
              f"{gbs:7.0f} {gbs / peak:7.3f} {a['tensor'] / max(a['us'], 1e-9):7.1f} {a['tc'] / max(a['us'], 1e-9):6.1f} "
              f"{a['l2hit'] / max(a['us'], 1e-9):6.1f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
