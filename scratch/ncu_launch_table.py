"""Per-launch table from an `ncu --csv --metrics ...` log (long format: one row per launch and metric -> one row per launch).

    python scratch/ncu_launch_table.py gpurun_out/flow_block.csv profiles/r02_ncu_flow_block_v30.csv
"""
import csv, re, sys


def main(src, dst):
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, im, iv, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per, order, metrics = {}, [], []
    for r in rows:
        if r is hdr or not r[ii].isdigit():
            continue
        key = int(r[ii])
        if key not in per:
            per[key] = {"name": re.sub(r"\(.*", "", r[ik]).replace("void ", "").strip()[:80]}
            order.append(key)
        per[key][r[im]] = r[iv].replace(",", "")
        if r[im] not in metrics:
            metrics.append(r[im])
    with open(dst, "w", newline="") as out:
        w = csv.writer(out)
        w.writerow(["id", "kernel"] + metrics + ["dram_MB", "dram_GBps"])
        for k in order:
            m = per[k]
            t = float(m.get("gpu__time_duration.sum", 0) or 0)
            mb = (float(m.get("dram__bytes_read.sum", 0) or 0) + float(m.get("dram__bytes_write.sum", 0) or 0)) / 1e6
            w.writerow([k, m["name"]] + [m.get(x, "") for x in metrics] + [round(mb, 2), round(mb / t * 1e6, 0) if t else ""])
    print(dst, len(order), "launches")


if __name__ == "__main__":
    main(*sys.argv[1:3])
