"""Per-kernel-family summary of an `ncu --csv --metrics ...` launch log (long format: one row per launch and metric).

    python scratch/ncu_family_summary.py gpurun_out/flow_launches.csv profiles/r02_ncu_flow_step_v28_summary.csv
"""
import csv, re, sys
from collections import defaultdict


def main(src, dst):
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, im, iv, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = defaultdict(dict)
    for r in rows:
        if r is hdr or not r[ii].isdigit():
            continue
        per[(int(r[ii]), r[ik])][r[im]] = float(r[iv].replace(",", "") or 0)
    fam = defaultdict(lambda: defaultdict(float))
    for (_, name), m in per.items():
        short = re.sub(r"\(.*", "", name).replace("void ", "").strip()
        short = re.sub(r"^at::native::.*?(\w+_kernel|\w+Functor|\w+Copy\w*).*", r"torch::\1", short)[:70]
        f = fam[short]
        f["launches"] += 1
        f["time_us"] += m.get("gpu__time_duration.sum", 0) / 1e3
        f["dram_mb"] += (m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)) / 1e6
        for k in ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                  "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active"):
            f[k] += m.get(k, 0) * m.get("gpu__time_duration.sum", 0)
    total = sum(f["time_us"] for f in fam.values())
    with open(dst, "w", newline="") as out:
        w = csv.writer(out)
        w.writerow(["kernel family", "launches", "time_us", "share_of_step", "dram_MB", "avg_GBps", "dram_pct (time-weighted)",
                    "sm_pct", "fma_pipe_pct", "warps_active_pct"])
        for name, f in sorted(fam.items(), key=lambda kv: -kv[1]["time_us"]):
            t = f["time_us"] * 1e3 or 1
            w.writerow([name, int(f["launches"]), round(f["time_us"], 1), round(f["time_us"] / total, 4), round(f["dram_mb"], 1),
                        round(f["dram_mb"] / max(f["time_us"], 1e-9) * 1e3, 0),
                        round(f["dram__throughput.avg.pct_of_peak_sustained_elapsed"] / t, 1),
                        round(f["sm__throughput.avg.pct_of_peak_sustained_elapsed"] / t, 1),
                        round(f["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"] / t, 1),
                        round(f["sm__warps_active.avg.pct_of_peak_sustained_active"] / t, 1)])
    print(dst, len(per), "launches", round(total / 1e3, 2), "ms")


if __name__ == "__main__":
    main(*sys.argv[1:3])
